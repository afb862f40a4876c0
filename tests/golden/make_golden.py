"""Generate the golden fixtures of tests/golden/ from the REFERENCE ITSELF.

Runs only in the build container: needs oracle/_ref/libbppref.so, i.e. the unmodified bpp v4.8.7
sources compiled by oracle/Makefile.  The reference ships no golden vectors for this path
(SURVEY.md F10), so outputs of the reference run here are the pins (prompt section 3).

    python tests/golden/make_golden.py

Each case_*.npz holds the seeded inputs (a bpp_b200.synth.Workload) and what the reference's
locus_update_matrices -> locus_update_partials -> locus_root_loglikelihood produced with
--arch avx2 semantics (attributes = PLL_ATTRIB_ARCH_AVX2): per-locus lnL, and for the first
`keep` loci every P-matrix, every inner CLV, every scaler and the eigen-decomposition.
A second pass does one mixing-style step (all times * c, flip every index, recompute), so the
double-buffer index scheme is pinned too.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import ref_set_from_workload, GOLDEN  # noqa: E402
from bpp_b200 import synth  # noqa: E402
from oracle import refbind  # noqa: E402

CASES = {
    # name: (workload kwargs, keep)
    "jc69_r1":      (dict(n_loci=6, tips=8, sites=67, states=4, rate_cats=1, model="JC69", seed=11), 3),
    "gtr_g4_scale": (dict(n_loci=5, tips=16, sites=45, states=4, rate_cats=4, model="GTR", scaling=True, seed=12), 2),
    "gtr_g4":       (dict(n_loci=4, tips=11, sites=33, states=4, rate_cats=4, model="GTR", seed=13), 2),
    "lg_g4":        (dict(n_loci=3, tips=8, sites=21, states=20, rate_cats=4, model="LG", seed=14), 1),
    # deep tree with long branches: per-site rescaling really fires (scalers > 0)
    "jc69_deep_scale": (dict(n_loci=2, tips=300, sites=12, states=4, rate_cats=1, model="JC69", scaling=True,
                             seed=15, dt_lo=0.05, dt_hi=0.4), 1),
    "gtr_g4_deep_scale": (dict(n_loci=2, tips=260, sites=9, states=4, rate_cats=4, model="GTR", scaling=True,
                               seed=16, dt_lo=0.05, dt_hi=0.4), 1),
    # closed-form DNA models (locus.c:1981-2324)
    "k80_g4":        (dict(n_loci=3, tips=7, sites=29, states=4, rate_cats=4, model="K80", seed=21), 1),
    "f81_r1_scale":  (dict(n_loci=3, tips=9, sites=31, states=4, rate_cats=1, model="F81", scaling=True, seed=22), 1),
    "hky_g4":        (dict(n_loci=3, tips=8, sites=27, states=4, rate_cats=4, model="HKY", seed=23), 1),
    "t92_g4":        (dict(n_loci=3, tips=6, sites=25, states=4, rate_cats=4, model="T92", seed=24), 1),
    "tn93_g4_scale": (dict(n_loci=3, tips=10, sites=23, states=4, rate_cats=4, model="TN93", scaling=True, seed=25), 1),
    "f84_r2":        (dict(n_loci=3, tips=5, sites=33, states=4, rate_cats=2, model="F84", seed=26, rates=[0.3, 1.7]), 1),
    "lg_g4_deep_scale": (dict(n_loci=1, tips=90, sites=6, states=20, rate_cats=4, model="LG", scaling=True,
                              seed=17, dt_lo=0.05, dt_hi=0.4), 1),
}

WORKLOAD_FIELDS = ("left", "right", "times", "rate_mui", "tip_chars", "weights", "freqs", "subst", "rates")


def workload_to_dict(w):
    d = {k: getattr(w, k) for k in WORKLOAD_FIELDS}
    d["meta"] = np.array([w.n_loci, w.tips, w.sites, w.states, w.rate_cats, int(w.scaling)], dtype=np.int64)
    d["model"] = np.array(w.model)
    return d


def dump_locus(rs, w, i):
    T = w.tips
    out = {}
    if T <= 32:     # deep-tree cases pin CLVs / scalers / lnL only (keeps the fixtures small)
        out["pmat"] = np.stack([rs.pmatrix(i, rs.node_get(i, n, 2)) for n in range(2 * T - 2)])
    out["clv"] = np.stack([rs.clv(i, rs.node_get(i, n, 0)) for n in range(T, 2 * T - 1)])
    if w.scaling:
        out["scaler"] = np.stack([rs.scaler(i, rs.node_get(i, n, 1)) for n in range(T, 2 * T - 1)])
    if w.model in ("GTR", "LG"):
        ev, iev, lam = rs.eigen(i)
        out["eigenvecs"], out["inv_eigenvecs"], out["eigenvals"] = ev, iev, lam
    return out


def main():
    rates_lg, freqs_lg = refbind.aa_lg()
    if len(sys.argv) == 1:
        np.savez(os.path.join(GOLDEN, "lg_model.npz"), rates=rates_lg, freqs=freqs_lg)
        np.savez(os.path.join(GOLDEN, "gamma_rates.npz"),
                 **{"a%g_k%d" % (a, k): refbind.gamma_rates(a, k)
                    for a in (0.05, 0.2, 0.5, 1.0, 2.5, 10.0) for k in (2, 3, 4, 5, 8)})
        nt, aa = refbind.char_maps()
        np.savez(os.path.join(GOLDEN, "char_maps.npz"), nt=nt, aa=aa)
    only = set(sys.argv[1:])          # optional: regenerate only the named cases
    for name, (kw, keep) in CASES.items():
        if only and name not in only:
            continue
        w = synth.make_workload(name, lg=(rates_lg, freqs_lg), **kw)
        rs = ref_set_from_workload(w)
        data = workload_to_dict(w)
        data["lnl"] = np.array([rs.full_pass(i) for i in range(w.n_loci)])
        for i in range(keep):
            for k, v in dump_locus(rs, w, i).items():
                data["l%d_%s" % (i, k)] = v
            lnl, persite = rs.root_loglikelihood(i, persite=True)
            data["l%d_persite" % i] = persite
        # one mixing-style step: times *= c, flip pmatrix/clv/scaler indices of every node, recompute
        c = 1.07
        T = w.tips
        lnl2 = []
        for i in range(w.n_loci):
            rs.set_times(i, w.times[i] * c)
            for n in range(2 * T - 2):
                rs.node_set(i, n, 2, (2 * T - 2 + rs.node_get(i, n, 2)) % (2 * (2 * T - 2)))
            for n in range(T, 2 * T - 1):
                rs.node_set(i, n, 0, T + (rs.node_get(i, n, 0) - 1) % (2 * T - 2))
                if w.scaling:
                    rs.node_set(i, n, 1, (T + rs.node_get(i, n, 1) - 1) % (2 * T - 2))
            lnl2.append(rs.full_pass(i))
        data["lnl_mix"] = np.array(lnl2)
        data["mix_c"] = np.array(c)
        if w.scaling:
            data["max_scaler"] = np.array(max(int(data["l%d_scaler" % i].max()) for i in range(keep)))
        np.savez_compressed(os.path.join(GOLDEN, "case_%s.npz" % name), **data)
        print(name, "lnl", data["lnl"][:3], "mix", data["lnl_mix"][:2],
              "max scaler", int(data["max_scaler"]) if w.scaling else "-")
        rs.close()


if __name__ == "__main__":
    main()
