"""CPU: pin the numpy oracle (oracle/felsenstein.py) against outputs of the reference itself.

Fixtures under tests/golden/ were produced by tests/golden/make_golden.py from the unmodified
bpp v4.8.7 sources (oracle/_ref).  Tolerances: CLVs bit-identical when the oracle is fed the
reference's own P-matrices (4 states, AVX association order); lnL relative error <= 1e-12
end to end (libm exp/expm1/log of numpy vs glibc differ by <= 1 ulp).
"""
import numpy as np
import pytest

from helpers import F, GOLDEN_CASES, char_map, load_case, refbind, rel_err, ref_set_from_workload, synth

LNL_RTOL = 1e-12


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_lnl_matches_reference_fixture(name):
    w, d = load_case(name)
    for i in range(w.n_loci):
        o = F.locus_from_workload(w, i, char_map(w.states))
        lnl = o.full_pass()
        assert abs(lnl - d["lnl"][i]) <= LNL_RTOL * abs(d["lnl"][i]), (name, i, lnl, d["lnl"][i])


@pytest.mark.parametrize("name", ["jc69_r1", "gtr_g4_scale", "gtr_g4", "jc69_deep_scale", "gtr_g4_deep_scale"])
def test_oracle_clv_bit_exact_given_reference_pmatrices(name):
    """4 states: with the reference's own P-matrices the oracle's CLVs, scalers and site lnL are
    bit-identical to --arch avx/avx2 (core_partials_avx.c:368-531)."""
    w, d = load_case(name)
    if "l0_pmat" not in d:
        pytest.skip("deep case stores no P-matrices")
    T = w.tips
    o = F.locus_from_workload(w, 0, char_map(4))
    for n in range(2 * T - 2):
        o.pmat[n] = d["l0_pmat"][n].reshape(w.rate_cats, 4, 4)
    o.update_partials(o.post_order())
    for k, n in enumerate(range(T, 2 * T - 1)):
        assert np.array_equal(o.clv[n].ravel(), d["l0_clv"][k]), (name, n)
        if w.scaling:
            assert np.array_equal(o.scale[n - T], d["l0_scaler"][k])
    lnl, persite = o.root_loglikelihood(persite=True)
    assert np.allclose(persite, d["l0_persite"], rtol=1e-15, atol=0)
    assert abs(lnl - d["lnl"][0]) <= 1e-15 * abs(lnl)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_clv_and_scalers_close(name):
    w, d = load_case(name)
    T = w.tips
    o = F.locus_from_workload(w, 0, char_map(w.states))
    o.full_pass()
    tol = 1e-9 if w.states == 20 else 1e-11       # small P entries of the 20-state eigen form cancel
    for k, n in enumerate(range(T, 2 * T - 1)):
        assert rel_err(o.clv[n].ravel(), d["l0_clv"][k]) < tol, (name, n)
        if w.scaling:
            assert np.array_equal(o.scale[n - T], d["l0_scaler"][k])
    if w.scaling and "deep" in name:
        assert int(d["max_scaler"]) >= 1            # the rescale branch really fired in the fixture


@pytest.mark.parametrize("name", ["gtr_g4", "lg_g4"])
def test_oracle_pmatrix_and_eigen(name):
    w, d = load_case(name)
    o = F.locus_from_workload(w, 0, char_map(w.states))
    ev, iev, lam = F.update_eigen(w.subst[0], w.freqs[0])
    S = w.states
    # eigenvalues are unique up to ordering; P(t) is invariant to the basis
    assert np.allclose(np.sort(lam), np.sort(d["l0_eigenvals"]), rtol=1e-10, atol=1e-13)
    edges = [n for n in range(2 * w.tips - 1) if o.parent[n] >= 0]
    o.update_matrices(edges)
    for n in edges:
        ref = d["l0_pmat"][n]
        assert np.max(np.abs(o.pmat[n].ravel() - ref)) < 1e-14, (name, n)
        rows = o.pmat[n].reshape(w.rate_cats, S, S).sum(axis=2)
        assert np.allclose(rows, 1.0, atol=1e-12)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_mixing_step_index_flips(name):
    """times *= c, flip every pmatrix/clv/scaler index (locus.c:24-26), recompute: pins the
    double-buffer index scheme against the reference."""
    w, d = load_case(name)
    c = float(d["mix_c"])
    T = w.tips
    for i in range(min(w.n_loci, 2)):
        o = F.locus_from_workload(w, i, char_map(w.states))
        o.full_pass()
        o.times = o.times * c
        for n in range(2 * T - 2):
            o.flip_pmatrix(n)
        for n in range(T, 2 * T - 1):
            o.flip_clv(n)
        lnl = o.full_pass()
        assert abs(lnl - d["lnl_mix"][i]) <= LNL_RTOL * abs(lnl)


def test_tip_clv_layout_and_illegal_code():
    m = synth.iupac_nt_map()
    clv = F.tip_clv(m[np.frombuffer(b"ACGTNRY-", dtype=np.uint8)], 4, 3)
    assert clv.shape == (8, 3, 4)
    assert clv[0, 0].tolist() == [1, 0, 0, 0] and clv[3, 2].tolist() == [0, 0, 0, 1]
    assert clv[4, 1].tolist() == [1, 1, 1, 1] and clv[5, 0].tolist() == [1, 0, 1, 0]
    with pytest.raises(ValueError):
        F.tip_clv(m[np.frombuffer(b"AC!T", dtype=np.uint8)], 4, 1)


def test_char_maps_match_reference_tables():
    d = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "char_maps.npz"))
    assert np.array_equal(synth.iupac_nt_map(), d["nt"])
    assert np.array_equal(synth.aa_map(), d["aa"])


def test_zero_length_branch_is_identity():
    assert np.array_equal(F.pmatrix_jc69(0.0, np.ones(1))[0], np.eye(4))
    ev, iev, lam = F.update_eigen(np.ones(6), np.full(4, 0.25))
    assert np.array_equal(F.pmatrix_eigen(ev, iev, lam, 1e-101, np.ones(2))[1], np.eye(4))


def test_scaling_threshold_is_strict_and_all_entries():
    """core_partials.c:720: strict <; a site is rescaled only if ALL S*R entries are below."""
    thr = F.SCALE_THRESHOLD
    eye = np.eye(4)[None]
    l = np.zeros((3, 1, 4))
    r = np.ones((3, 1, 4))
    l[0, 0] = [thr, thr / 2, thr / 2, thr / 2]        # one entry == threshold -> not rescaled
    l[1, 0] = [thr / 2] * 4                           # all below -> rescaled
    l[2, 0] = [0, 0, 0, 0]                            # exact zeros count as below
    clv, sc = F.update_partial_ii(l, r, eye, eye, None, np.array([1, 2, 3], dtype=np.uint32), scaling=True)
    assert sc.tolist() == [1, 3, 4]
    assert clv[1, 0, 0] == 0.5 and clv[0, 0, 0] == thr


@pytest.mark.skipif(not refbind.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("cfg", [dict(tips=5, sites=31, states=4, rate_cats=1, model="JC69"),
                                 dict(tips=9, sites=17, states=4, rate_cats=4, model="GTR", scaling=True),
                                 dict(tips=6, sites=9, states=20, rate_cats=4, model="LG")])
def test_oracle_vs_live_reference_fresh_inputs(cfg):
    from helpers import lg_tables
    w = synth.make_workload("live", n_loci=3, seed=991, lg=lg_tables(), **cfg)
    rs = ref_set_from_workload(w)
    for i in range(w.n_loci):
        ref = rs.full_pass(i)
        lnl = F.locus_from_workload(w, i, char_map(w.states)).full_pass()
        assert abs(lnl - ref) <= LNL_RTOL * abs(ref)
    rs.close()


@pytest.mark.skipif(not refbind.available(), reason="oracle/_ref not built")
def test_reference_archs_agree():
    """the reference's own CPU/SSE/AVX/AVX2 kernels agree to ~1e-15 in lnL."""
    w = synth.make_workload("archs", n_loci=2, tips=8, sites=40, states=4, rate_cats=4, model="GTR", seed=5)
    vals = []
    for arch in (refbind.ARCH_CPU, refbind.ARCH_SSE, refbind.ARCH_AVX, refbind.ARCH_AVX2):
        rs = ref_set_from_workload(w, arch=arch)
        vals.append([rs.full_pass(i) for i in range(w.n_loci)])
        rs.close()
    vals = np.array(vals)
    assert np.max(np.abs(vals - vals[0]) / np.abs(vals[0])) < 1e-13


def test_oracle_on_frogs_real_data():
    """BASELINE.json config 1: frogs A00, 5 diploid JC69 loci, real data through the reference's own
    parsing / compression / phase resolution.  Known answer log-L0 = -7320.932289."""
    from helpers import frogs_fixture, frogs_oracle_locus
    d = frogs_fixture()
    total = 0.0
    for k in range(int(d["n_loci"])):
        p = "l%d_" % k
        o = frogs_oracle_locus(d, k)
        edges = [n for n in range(2 * o.tips - 1) if o.parent[n] >= 0]
        o.update_matrices(edges)
        o.update_partials(o.post_order())
        lh = o.root_likelihood_vector()
        # per-site likelihoods down to 1e-54 after up to 59 nodes: libm exp of numpy vs glibc shows
        assert np.allclose(lh, d[p + "likelihood_vector"], rtol=1e-10, atol=0)
        lnl = F.diploid_loglikelihood(lh, d[p + "resolution_count"], d[p + "mapping"], d[p + "weights"])
        assert abs(lnl - float(d[p + "logl"])) <= 1e-12 * abs(lnl)
        total += lnl
    assert abs(total - (-7320.932289)) < 5e-6


@pytest.mark.parametrize("model", ["K80", "F81", "HKY", "TN93", "F84"])
def test_closed_form_pmatrix_properties(model):
    """Size-independent properties of the closed-form matrices (locus.c:1981-2324) as restated in the oracle: rows
    sum to 1, P(0) = I, Chapman-Kolmogorov P(s)P(t) = P(s+t), and detailed balance pi_i P_ij = pi_j P_ji."""
    rng = np.random.default_rng(3)
    f = rng.uniform(0.5, 1.5, 4)
    f /= f.sum()
    if model == "K80":
        f = np.full(4, 0.25)
    q = rng.uniform(0.5, 4.0, 6)
    rates = np.array([0.3, 1.0, 2.2])
    P0 = F.pmatrix_closed(model, 0.0, rates, f, q)
    assert np.allclose(P0, np.eye(4)[None], atol=1e-15)
    s, t = 0.037, 0.21
    Ps, Pt, Pst = (F.pmatrix_closed(model, x, rates, f, q) for x in (s, t, s + t))
    for r in range(len(rates)):
        assert np.allclose(Pt[r].sum(axis=1), 1.0, atol=1e-14)
        assert np.allclose(Ps[r] @ Pt[r], Pst[r], atol=1e-14)
        assert np.allclose(f[:, None] * Pt[r], (f[:, None] * Pt[r]).T, atol=1e-15)
        assert (Pt[r] > 0).all()
