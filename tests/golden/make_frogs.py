"""Config 1 of BASELINE.json (frogs, A00, JC69, 5 diploid loci) as a golden fixture.

Build container only.  Runs the UNMODIFIED reference on its own shipped example
(/root/reference/examples/frogs) through oracle/_ref/libbppref_hook.so, whose method.o is compiled
with the first-lnL call of init() (method.c:4297) redirected to oracle/ref_hook.c.  The hook dumps, per
locus, what BPP's own PHYLIP parsing, site-pattern compression and diploid phase resolution handed to
the likelihood path, plus the reference's answer.  Known answer:
log-L0 = -7320.932289 (sum over the 5 loci).

    python tests/golden/make_frogs.py      ->  tests/golden/frogs_A00.npz
"""
import ctypes as C
import os
import shutil
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EX = "/root/reference/examples/frogs"
WORK = os.path.join(ROOT, "oracle", "_ref", "frogs_work")
HOOK = os.path.join(ROOT, "oracle", "_ref", "libbppref_hook.so")


def run_reference():
    shutil.rmtree(WORK, ignore_errors=True)
    os.makedirs(WORK)
    for f in ("frogs.txt", "frogs.Imap.txt"):
        shutil.copy(os.path.join(EX, f), WORK)
    ctl = open(os.path.join(EX, "A00.bpp.ctl")).read().splitlines()
    out = []
    for line in ctl:
        if line.strip().startswith("finetune"):
            line = "finetune = 1"                    # v4.8.7 rejects the shipped pre-4.8.1 syntax
        if line.strip().startswith("seed"):
            line = "seed = 12345"
        out.append(line)
    open(os.path.join(WORK, "A00.bpp.ctl"), "w").write("\n".join(out) + "\n")
    code = ("import ctypes as C, sys; L = C.CDLL(%r); "
            "argv = (C.c_char_p * 4)(b'bpp', b'--cfile', b'A00.bpp.ctl', None); L.bpp_main(3, argv)" % HOOK)
    env = dict(os.environ, BPP_HOOK_DUMP=os.path.join(WORK, "dump.bin"))
    r = subprocess.run([sys.executable, "-c", code], cwd=WORK, env=env, capture_output=True, text=True)
    if not os.path.exists(os.path.join(WORK, "dump.bin")):
        raise RuntimeError("reference run failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return open(os.path.join(WORK, "dump.bin"), "rb").read()


class Reader:
    def __init__(self, b):
        self.b, self.o = b, 0

    def i(self):
        v = struct.unpack_from("<q", self.b, self.o)[0]
        self.o += 8
        return v

    def d(self):
        v = struct.unpack_from("<d", self.b, self.o)[0]
        self.o += 8
        return v

    def arr(self, dtype, n):
        a = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.o).copy()
        self.o += a.nbytes
        return a


def main():
    r = Reader(run_reference())
    n = r.i()
    data = {"n_loci": np.array(n)}
    total = 0.0
    for k in range(n):
        tips, sites, states, cats, model, dtype, diploid, unphased, scale_buffers = [r.i() for _ in range(9)]
        logl, bfbeta = r.d(), r.d()
        freqs = r.arr("<f8", states)
        rates = r.arr("<f8", cats)
        rw = r.arr("<f8", cats)
        clv = r.arr("<f8", tips * sites * states).reshape(tips, sites, states)
        masks = (clv.astype(np.uint32) << np.arange(states, dtype=np.uint32)).sum(axis=2).astype(np.uint8)
        assert np.array_equal(((masks[:, :, None] >> np.arange(states)) & 1).astype(np.float64), clv)
        p = "l%d_" % k
        if diploid:
            data[p + "weights"] = r.arr("<u4", unphased)
            data[p + "resolution_count"] = r.arr("<u8", unphased)
            maplen = r.i()
            data[p + "mapping"] = r.arr("<u8", maplen)
            data[p + "likelihood_vector"] = r.arr("<f8", sites)
        else:
            data[p + "weights"] = r.arr("<u4", sites)
        nn = r.i()
        nodes = np.zeros((nn, 7), dtype=np.int64)
        tl = np.zeros((nn, 2))
        for j in range(nn):
            nodes[j] = [r.i() for _ in range(7)]
            tl[j] = [r.d(), r.d()]
        nodes[nodes == 0xFFFFFFFF] = -1       # absent child/parent (the hook prints an unsigned -1)
        root_clv = r.arr("<f8", sites * states * cats)
        data[p + "dims"] = np.array([tips, sites, states, cats, model, dtype, diploid, unphased, scale_buffers])
        data[p + "logl"] = np.array(logl)
        data[p + "freqs"], data[p + "rates"], data[p + "rate_weights"] = freqs, rates, rw
        data[p + "tip_masks"] = masks
        data[p + "nodes"] = nodes       # pre-order: node_index, left, right, parent, clv_index, scaler_index, pmatrix_index
        data[p + "time_length"] = tl    # node time, branch length (node->length as locus_update_matrices left it)
        data[p + "root_clv"] = root_clv
        total += logl
        print("locus %d: tips %d sites %d diploid %d unphased %d  lnL %.6f" % (k, tips, sites, diploid, unphased, logl))
    data["logl_sum"] = np.array(total)
    print("log-L0 = %.6f" % total)
    assert abs(total - (-7320.932289)) < 5e-6, total
    np.savez_compressed(os.path.join(HERE, "frogs_A00.npz"), **data)


if __name__ == "__main__":
    main()
