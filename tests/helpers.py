"""Shared test helpers (CPU side)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bpp_b200 import synth  # noqa: E402
from oracle import felsenstein as F  # noqa: E402
from oracle import refbind  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def char_map(states):
    return synth.iupac_nt_map() if states == 4 else synth.aa_map()


def ref_set_from_workload(w, arch=refbind.ARCH_AVX2):
    """Load a synth.Workload into the compiled reference (oracle/_ref)."""
    model = refbind.AA_MODEL_LG if w.model == "LG" else refbind.DNA_MODELS[w.model]
    rs = refbind.RefSet(w.n_loci, w.states, w.rate_cats, w.scaling, model=model, arch=arch)
    for i in range(w.n_loci):
        rs.create(i, w.tips, w.sites)
        for t in range(w.tips):
            rs.set_tip_states(i, t, w.tip_chars[i, t].tobytes())
        rs.set_weights(i, w.weights[i])
        if w.model == "JC69":
            # locus_set_frequencies_and_rates (locus.c:899) sets pi = 1/4 for JC69
            rs.set_model(i, freqs=w.freqs[i], rates=w.rates)
        else:
            rs.set_model(i, freqs=w.freqs[i], subst=w.subst[i], rates=w.rates)
        rs.set_tree(i, w.left[i], w.right[i], w.times[i], float(w.rate_mui[i]))
    return rs


def lg_tables():
    d = np.load(os.path.join(GOLDEN, "lg_model.npz"))
    return d["rates"], d["freqs"]


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def load_case(name):
    """A golden case as (Workload, dict of reference outputs)."""
    d = np.load(os.path.join(GOLDEN, "case_%s.npz" % name))
    n, T, P, S, R, sc = [int(x) for x in d["meta"]]
    w = synth.Workload(name, n, T, P, S, R, str(d["model"]), bool(sc), d["left"], d["right"], d["times"],
                       d["rate_mui"], d["tip_chars"], d["weights"], d["freqs"], d["subst"], d["rates"])
    return w, d


GOLDEN_CASES = ["jc69_r1", "gtr_g4_scale", "gtr_g4", "lg_g4", "jc69_deep_scale", "gtr_g4_deep_scale",
                "lg_g4_deep_scale", "k80_g4", "f81_r1_scale", "hky_g4", "t92_g4", "tn93_g4_scale", "f84_r2"]


def frogs_fixture():
    return np.load(os.path.join(GOLDEN, "frogs_A00.npz"))


def frogs_oracle_locus(d, k):
    """Locus k of the frogs A00 fixture (real data as the reference's init() saw it) as an OracleLocus."""
    p = "l%d_" % k
    tips, sites, states, cats = [int(x) for x in d[p + "dims"][:4]]
    o = F.OracleLocus(tips, sites, states, cats, scaling=False, model="JC69")
    nodes = d[p + "nodes"]
    nn = 2 * tips - 1
    left, right = np.zeros(tips - 1, dtype=np.int64), np.zeros(tips - 1, dtype=np.int64)
    times = np.zeros(nn)
    for row, tl in zip(nodes, d[p + "time_length"]):
        idx = int(row[0])
        times[idx] = tl[0]
        if row[1] >= 0:
            left[idx - tips], right[idx - tips] = row[1], row[2]
    o.set_tree(left, right, times, 1.0)
    for t in range(tips):
        o.set_tip_masks(t, d[p + "tip_masks"][t].astype(np.uint32))
    o.set_model(freqs=d[p + "freqs"], rates=d[p + "rates"])
    return o
