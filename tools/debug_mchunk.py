import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from bpp_b200 import engine, synth
from oracle import felsenstein as F
from helpers import char_map

tips, R = 29, 2
w = synth.make_workload("mchunk", n_loci=6, tips=tips, sites=530, states=4, rate_cats=R, model="GTR", scaling=False, seed=1000 + tips, rates=[0.3, 1.7])
eng = engine.Engine(0)
loci, trees = engine.load_workload(eng, w)
batch = engine.Batch(eng, loci)
step = trees.full_pass_step()
lnl, tot = batch.full_pass(step)
cm = char_map(4)
i = 5
o = F.locus_from_workload(w, i, cm)
ref = o.full_pass()
print("lnl", lnl[i], ref)
mc, mi, mb, oc, ops, rc, rs = step
ooff = np.concatenate([[0], np.cumsum(oc)]).astype(int)
lo = ops[ooff[i]:ooff[i + 1]]
for k, op in enumerate(lo):
    n = int(op["parent_clv_index"])
    got = loci[i].get_clv(n).reshape(-1, R, 4)
    exp = np.asarray(o.clv[n]).reshape(-1, R, 4)
    err = np.abs(got - exp) / np.maximum(np.abs(exp), 1e-300)
    bad = np.argwhere(err.max(axis=2) > 1e-9)
    print("op %2d parent %2d left %2d right %2d  maxerr %.2e  bad cells %d  first bad %s" % (
        k, n, op["left_clv_index"], op["right_clv_index"], err.max(), len(bad), bad[:3].tolist()))

# which tip pair reproduces what the device computed for node 47 = tips 20 x 11 ?
n, a, b = 47, 20, 11
got = loci[i].get_clv(n).reshape(-1, R, 4)
Pa, Pb = o.pmat[o.pmatrix_index[a]], o.pmat[o.pmatrix_index[b]]      # [R,4,4]
def xvec(P, clv):
    return np.einsum("rij,prj->pri", P, np.asarray(clv).reshape(-1, R, 4))
best = []
for ta in range(tips):
    for tb in range(tips):
        e = xvec(Pa, o.clv[ta]) * xvec(Pb, o.clv[tb])
        err = np.abs(e - got).max()
        best.append((err, ta, tb))
best.sort()
print("closest tip pairs for node 47 (expected 20, 11):", best[:5])
print("tip chars 20:", bytes(w.tip_chars[i, 20][:24]), " 11:", bytes(w.tip_chars[i, 11][:24]))
print("got[0..3]", got[:3, 0], "\nexp", np.asarray(o.clv[n]).reshape(-1, R, 4)[:3, 0])

# which P-matrix (index, category) was applied to tip 20 / tip 11 for node 47?
ca = np.asarray(o.clv[20]).reshape(-1, R, 4); cb = np.asarray(o.clv[11]).reshape(-1, R, 4)
allP = [loci[i].get_pmatrix(k).reshape(R, 4, 4) for k in range(2 * (2 * tips - 2))]
for cat in range(R):
    res = []
    for ka in range(len(allP)):
        for ra in range(R):
            xa = np.einsum("ij,pj->pi", allP[ka][ra], ca[:, cat])
            for kb in (11,):
                xb = np.einsum("ij,pj->pi", allP[kb][cat], cb[:, cat])
                res.append((np.abs(xa * xb - got[:, cat]).max(), ka, ra))
    res.sort()
    print("cat", cat, "best P for tip 20 (expected index 20, cat %d):" % cat, res[:3])
    res = []
    for kb in range(len(allP)):
        for rb in range(R):
            xb = np.einsum("ij,pj->pi", allP[kb][rb], cb[:, cat])
            xa = np.einsum("ij,pj->pi", allP[20][cat], ca[:, cat])
            res.append((np.abs(xa * xb - got[:, cat]).max(), kb, rb))
    res.sort()
    print("cat", cat, "best P for tip 11 (expected index 11):", res[:3])
