"""GPU: parity of the CUDA path (through the C-ABI) with the oracle and the reference fixtures.

Bars (BASELINE.json north_star): per-locus lnL relative error <= 1e-10 against the reference;
scalers (integers) exactly equal; 4-state CLVs BIT-IDENTICAL to the reference's AVX kernels
when both sides use the same P-matrices (math="exact").
"""
import numpy as np
import pytest

from helpers import F, GOLDEN_CASES, char_map, lg_tables, load_case, rel_err, synth

pytestmark = pytest.mark.gpu

LNL_RTOL = 1e-10


@pytest.fixture(scope="module")
def eng():
    from bpp_b200 import engine
    e = engine.Engine(0, math="exact")
    yield e
    e.close()


def _load(eng, w):
    from bpp_b200 import engine
    loci, trees = engine.load_workload(eng, w)
    return loci, trees, engine.Batch(eng, loci)


def _free(loci, batch):
    batch.destroy()
    for l in loci:
        l.destroy()


@pytest.mark.parametrize("math", ["exact", "fma"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_full_pass_lnl_matches_reference_fixture(eng, name, math):
    w, d = load_case(name)
    eng.set_math(math)
    loci, trees, batch = _load(eng, w)
    lnl, total = batch.full_pass(trees.full_pass_step())
    eng.set_math("exact")
    assert rel_err(lnl, d["lnl"]) <= LNL_RTOL, (name, lnl, d["lnl"])
    assert abs(total - d["lnl"].sum()) <= LNL_RTOL * abs(total)
    _free(loci, batch)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_clv_scaler_pmatrix_match_reference_fixture(eng, name):
    w, d = load_case(name)
    T = w.tips
    loci, trees, batch = _load(eng, w)
    batch.full_pass(trees.full_pass_step())
    l = loci[0]
    tol = 1e-9 if w.states == 20 else 1e-11
    for k, n in enumerate(range(T, 2 * T - 1)):
        assert rel_err(l.get_clv(n), d["l0_clv"][k]) < tol, (name, n)
        if w.scaling:
            assert np.array_equal(l.get_scaler(n - T), d["l0_scaler"][k]), (name, n)
    if "l0_pmat" in d:
        for n in range(2 * T - 2):
            assert np.max(np.abs(l.get_pmatrix(n) - d["l0_pmat"][n])) < 1e-14
    _free(loci, batch)


@pytest.mark.parametrize("name", ["gtr_g4_scale", "gtr_g4_deep_scale", "f81_r1_scale", "tn93_g4_scale"])
def test_specialised_scaled_launch_reproduces_clvs_and_scalers(eng, name):
    """Runs on a cached plan of scaled one-chunk loci launch the specialised kernel (SCALED_ONLY), which evaluates a
    tile speculatively without scaler bookkeeping and repeats it only where a site really has to be rescaled
    (the deep-tree fixture: scalers up to 2).  Every inner CLV and every scaler must equal the reference's, exactly
    as after the first run on the kernel that carries all paths."""
    w, d = load_case(name)
    T = w.tips
    loci, trees, batch = _load(eng, w)
    batch.stage(trees.full_pass_step())
    batch.run()
    lnl0, _ = batch.collect()
    l = loci[0]
    first = [(l.get_clv(n).copy(), l.get_scaler(n - T).copy()) for n in range(T, 2 * T - 1)]
    # scribble over the scalers so that the later runs have to write every one of them
    for _ in range(3):
        batch.run()
        lnl, _ = batch.collect()
    if T <= 17:
        assert batch.kernel_name.endswith(",scaled_only>"), batch.kernel_name
    assert np.array_equal(lnl, lnl0)
    for k, n in enumerate(range(T, 2 * T - 1)):
        assert np.array_equal(l.get_clv(n), first[k][0]), (name, n)
        assert np.array_equal(l.get_scaler(n - T), first[k][1]), (name, n)
        assert np.array_equal(l.get_scaler(n - T), d["l0_scaler"][k]), (name, n)
    _free(loci, batch)


@pytest.mark.parametrize("rate_cats,tips,dt", [(4, 16, 1e-12), (1, 12, 1e-12), (2, 16, 1e-12), (8, 8, 1e-15)])
def test_specialised_scaled_launch_with_real_rescaling(eng, rate_cats, tips, dt):
    """One-chunk trees whose sites DO get rescaled (branches of 1e-12: every mismatch costs ten orders of magnitude):
    the specialised scaled launch must notice it in its speculative pass and repeat those tiles with the rescaling
    in place.  Scalers and lnL against the oracle, CLVs and scalers against the first run (general kernel)."""
    rates = None if rate_cats in (1, 4) else list(np.linspace(0.3, 1.7, rate_cats))
    w = synth.make_workload("resc", n_loci=5, tips=tips, sites=333, states=4, rate_cats=rate_cats, model="GTR",
                            scaling=True, seed=77, dt_lo=dt, dt_hi=10 * dt, rates=rates)
    T = w.tips
    loci, trees, batch = _load(eng, w)
    batch.stage(trees.full_pass_step())
    batch.run()
    lnl0, _ = batch.collect()
    first = [[(l.get_clv(n).copy(), l.get_scaler(n - T).copy()) for n in range(T, 2 * T - 1)] for l in loci]
    for _ in range(3):
        batch.run()
        lnl, _ = batch.collect()
    assert batch.kernel_name.endswith(",scaled_only>"), batch.kernel_name
    assert np.array_equal(lnl, lnl0)
    cm = char_map(4)
    fired = 0
    for i, l in enumerate(loci):
        o = F.locus_from_workload(w, i, cm)
        ref = o.full_pass()
        assert abs(lnl[i] - ref) <= LNL_RTOL * abs(ref), (i, lnl[i], ref)
        for k, n in enumerate(range(T, 2 * T - 1)):
            sc = l.get_scaler(n - T)
            assert np.array_equal(l.get_clv(n), first[i][k][0]), (i, n)
            assert np.array_equal(sc, first[i][k][1]), (i, n)
            assert np.array_equal(sc, o.scale[o.scaler_index[n]]), (i, n)
            fired += int(sc.max())
    assert fired > 0, "the workload was meant to trigger per-site rescaling"
    _free(loci, batch)


@pytest.mark.parametrize("rate_cats,states,dt", [(5, 4, 0.01), (5, 4, 1e-12), (3, 4, 0.01), (7, 4, 0.01), (3, 20, 0.01)])
def test_category_counts_that_are_not_powers_of_two(eng, rate_cats, states, dt):
    """BPP users run 5 rate categories as readily as 4.  The device holds 4 or 8 (the extra ones are copies of category
    0 with weight 0: +0.0 in every site likelihood, and no influence on whether ALL categories of a site are below the
    scaling threshold), the accessors speak the caller's count.  lnL, every inner CLV, every scaler (dt = 1e-12: with
    real rescaling) and the P-matrices against the oracle; set_pmatrix / get_pmatrix round trip."""
    model = "GTR" if states == 4 else "LG"
    rates = list(np.linspace(0.3, 1.9, rate_cats))
    w = synth.make_workload("odd", n_loci=3, tips=11 if states == 4 else 6, sites=150, states=states, rate_cats=rate_cats,
                            model=model, scaling=True, seed=31, dt_lo=dt, dt_hi=10 * dt, rates=rates, lg=lg_tables())
    T, R, S = w.tips, rate_cats, states
    loci, trees, batch = _load(eng, w)
    lnl, _ = batch.full_pass(trees.full_pass_step())
    assert batch.kernel_name.startswith("tree_kernel_s4<%d" % (4 if R < 4 else 8)) or states == 20, batch.kernel_name
    cm = char_map(states)
    tol = 1e-9 if states == 20 else 1e-11
    fired = 0
    for i, l in enumerate(loci):
        o = F.locus_from_workload(w, i, cm)
        ref = o.full_pass()
        assert abs(lnl[i] - ref) <= LNL_RTOL * abs(ref), (i, lnl[i], ref)
        for n in range(T, 2 * T - 1):
            got = l.get_clv(n)
            assert got.size == w.sites * R * S
            assert rel_err(got.reshape(w.sites, R, S), o.clv[o.clv_index[n]]) < tol, (i, n)
            sc = l.get_scaler(n - T)
            assert np.array_equal(sc, o.scale[o.scaler_index[n]]), (i, n)
            fired += int(sc.max())
        for n in range(2 * T - 2):
            pm = l.get_pmatrix(n)
            assert pm.size == R * S * S
            assert np.max(np.abs(pm.reshape(R, S, S) - o.pmat[o.pmatrix_index[n]])) < 1e-13
        probe = np.arange(R * S * S, dtype=np.float64) / (R * S * S)
        l.set_pmatrix(0, probe)
        assert np.array_equal(l.get_pmatrix(0), probe)
    if dt < 1e-6:
        assert fired > 0, "the workload was meant to trigger per-site rescaling"
    # a packed tip handed out with the caller's category count
    tip = loci[0].get_clv(0)
    assert tip.size == w.sites * R * S
    _free(loci, batch)


@pytest.mark.parametrize("name", ["jc69_r1", "gtr_g4_scale", "gtr_g4"])
def test_clv_bit_exact_with_reference_pmatrices(eng, name):
    """Upload the reference's own P-matrices: every inner CLV, scaler and the per-site lnL must be
    bit-identical to --arch avx/avx2 (core_partials_avx.c:368-531, core_likelihood_avx.c:98-157)."""
    w, d = load_case(name)
    T = w.tips
    loci, trees, batch = _load(eng, w)
    l = loci[0]
    for n in range(2 * T - 2):
        l.set_pmatrix(n, d["l0_pmat"][n])
    mc, mi, mb, oc, ops, rc, rs = trees.full_pass_step()
    ops0 = ops[:T - 1]
    l.update_partials(ops0)
    for k, n in enumerate(range(T, 2 * T - 1)):
        assert np.array_equal(l.get_clv(n), d["l0_clv"][k]), (name, n)
        if w.scaling:
            assert np.array_equal(l.get_scaler(n - T), d["l0_scaler"][k])
    lnl, persite = l.root_loglikelihood(int(rc[0]), int(rs[0]), persite=True)
    assert np.allclose(persite, d["l0_persite"], rtol=4e-16, atol=0)       # CUDA log vs glibc log: <= 1 ulp
    assert abs(lnl - d["lnl"][0]) <= 1e-13 * abs(lnl)
    _free(loci, batch)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mixing_step_with_index_flips(eng, name):
    """prop_mixing_update_gtrees (prop_mixing.c:52-220): scale ages, flip every index to the spare
    buffers, recompute; then flip back (rejection) and read the untouched old lnL again."""
    w, d = load_case(name)
    loci, trees, batch = _load(eng, w)
    lnl0, _ = batch.full_pass(trees.full_pass_step())
    old_root = trees.clv_index[:, trees.root].copy()
    old_sc = trees.scaler_index[:, trees.root].copy()
    trees.times = trees.times * float(d["mix_c"])
    trees.flip_pmatrix()
    trees.flip_clv()
    lnl1, _ = batch.full_pass(trees.full_pass_step())
    assert rel_err(lnl1, d["lnl_mix"]) <= LNL_RTOL
    # rejection: the old buffers still hold the old state
    again = batch.root_loglikelihood(old_root, old_sc)
    assert np.array_equal(again, lnl0)
    _free(loci, batch)


@pytest.mark.parametrize("name", ["jc69_r1", "gtr_g4_scale", "gtr_g4_deep_scale", "lg_g4"])
def test_device_side_flip_and_cached_plan(eng, name):
    """The same mixing move with the lists resident on the device: bppgpu_batch_flip_indices applies the SWAP_*
    flips of every node (locus.c:24-26) to the staged step, bppgpu_batch_set_branch_lengths sends the only thing
    that changed, and from the third step on the planned program of either index parity is reused (refresh of the
    P-matrices only).  Every variant must give bit-identical numbers to staging the host-flipped arrays."""
    w, d = load_case(name)
    loci, trees, batch = _load(eng, w)
    step0 = trees.full_pass_step()
    batch.stage(step0)
    batch.run()
    lnl0, _ = batch.collect()
    batch.run()                                   # same stage again: cached plan, matrices refreshed
    again, _ = batch.collect()
    assert np.array_equal(again, lnl0)
    # proposal 1: scale the ages, flip on the device, send branch lengths
    trees.times = trees.times * float(d["mix_c"])
    trees.flip_pmatrix()
    trees.flip_clv()
    step1 = trees.full_pass_step()
    batch.flip_indices()
    batch.set_branch_lengths(step1[2])
    batch.run()
    lnl1, tot1 = batch.collect()
    assert rel_err(lnl1, d["lnl_mix"]) <= LNL_RTOL
    # rejected: flip back; proposal 2 with other ages lands in the same spare buffers, plan of parity 1 is cached now
    batch.flip_indices()
    c2 = 0.5 * (1.0 + float(d["mix_c"]))
    trees.times = trees.times * c2
    step2 = trees.full_pass_step()
    batch.flip_indices()
    bl2 = engine_pinned(step2[2])
    batch.set_branch_lengths(bl2.array)
    batch.run()
    batch.wait_inputs()
    bl2.array[:] = -1.0                           # the device has read them: the host may scribble
    lnl2, tot2 = batch.collect()
    bl2.free()
    # accepted this time; proposal 3 flips to parity 0 (cached) with new ages
    trees.times = trees.times * 1.01
    trees.flip_pmatrix()
    trees.flip_clv()
    step3 = trees.full_pass_step()
    batch.flip_indices()
    batch.set_branch_lengths(step3[2])
    batch.run()
    lnl3, _ = batch.collect()
    # the same three states through fresh stages of host-flipped arrays
    for step, got in ((step1, lnl1), (step2, lnl2), (step3, lnl3)):
        want, _ = batch.full_pass(step)
        assert np.array_equal(got, want)
    assert abs(tot2 - lnl2.sum()) <= 1e-9 * abs(tot2)
    _free(loci, batch)


@pytest.mark.parametrize("cfg", [dict(tips=8, rate_cats=1, model="JC69", scaling=False),
                                 dict(tips=16, rate_cats=4, model="GTR", scaling=True),
                                 dict(tips=48, rate_cats=4, model="GTR", scaling=False),
                                 dict(tips=100, rate_cats=4, model="GTR", scaling=True),
                                 dict(tips=128, rate_cats=1, model="JC69", scaling=False)])
def test_age_moves_across_all_loci_match_the_reference(eng, cfg):
    """The schedule for the most frequent proposal (gene-tree age move, gtree.c:4585, 5437-5467): the SAME move in
    every locus as one batch -- 2-3 P-matrices and the root path's partials per locus, ragged, children outside
    the path read from HBM.  Three rounds of moves on top of each other against the compiled reference, which does
    them locus by locus; indices are flipped exactly as the reference flips them."""
    from bpp_b200 import engine
    from helpers import ref_set_from_workload
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built")
    w = synth.make_workload("age", n_loci=40, sites=333, states=4, seed=77, **cfg)
    loci, trees, batch = _load(eng, w)
    rs = ref_set_from_workload(w)
    lnl, _ = batch.full_pass(trees.full_pass_step())
    _, ref = rs.full_pass_all(0, w.n_loci, 1, 1)
    assert rel_err(lnl, ref) <= LNL_RTOL
    rng = np.random.default_rng(5)
    for _ in range(3):
        nodes, ages = engine.propose_ages(trees, rng)
        step = engine.age_move_step(trees, nodes, ages)
        lnl, total = batch.full_pass(step)
        st = batch.plan_stats()              # root paths stay on the fast path whatever the size of the tree (<= 128 tips)
        assert st["walker"] == 0 and st["fast"] == w.n_loci, st
        _, ref = rs.age_move_all(0, w.n_loci, nodes, ages, 2)
        assert rel_err(lnl, ref) <= LNL_RTOL
        assert abs(total - lnl.sum()) <= 1e-9 * abs(total)
    rs.close()
    _free(loci, batch)


def test_batch_calls_reject_out_of_range_indices(eng):
    """Indices address a shared arena, so one that is out of range would overwrite another locus' buffers without any
    fault; the batch entry points check every index of a step whose shape changed."""
    from bpp_b200 import engine
    w = synth.make_workload("rng", n_loci=6, tips=5, sites=40, states=4, rate_cats=1, model="JC69", seed=9)
    loci, trees, batch = _load(eng, w)
    good = trees.full_pass_step()
    lnl, _ = batch.full_pass(good)
    for field, value in (("parent_clv_index", 5 + 8), ("left_clv_index", 99), ("right_pmatrix_index", 16), ("parent_scaler_index", 0)):
        mc, mi, mb, oc, ops, rc, rs = [np.array(a, copy=True) for a in good]
        ops[7][field] = value
        oc[0], oc[1] = oc[0] - 1, oc[1] + 1                  # another shape: the check runs
        with pytest.raises(engine.BppGpuError):
            batch.full_pass((mc, mi, mb, oc, ops, rc, rs))
    mc, mi, mb, oc, ops, rc, rs = [np.array(a, copy=True) for a in good]
    mi[3] = 16                                               # prob_matrices = 2 * (2T - 2) = 16
    mc[0], mc[1] = mc[0] - 1, mc[1] + 1
    with pytest.raises(engine.BppGpuError):
        batch.full_pass((mc, mi, mb, oc, ops, rc, rs))
    again, _ = batch.full_pass(good)                         # the batch is still usable
    assert np.array_equal(again, lnl)
    _free(loci, batch)


def engine_pinned(a):
    from bpp_b200 import engine
    return engine.PinnedArray(np.ascontiguousarray(a, dtype=np.float64))


@pytest.mark.parametrize("cfg", [
    dict(tips=8, sites=1000, states=4, rate_cats=1, model="JC69"),                  # config-2 shape
    dict(tips=16, sites=1000, states=4, rate_cats=4, model="GTR", scaling=True),     # config-3 shape
    dict(tips=8, sites=500, states=20, rate_cats=4, model="LG"),                     # config-4 shape
    dict(tips=2, sites=1, states=4, rate_cats=1, model="JC69"),                      # smallest legal locus
    dict(tips=3, sites=257, states=4, rate_cats=2, model="GTR", rates=[0.4, 1.6]),   # ragged last tile
    dict(tips=7, sites=300, states=4, rate_cats=3, model="GTR", rates=[0.2, 0.9, 1.9]),   # R not a power of 2: 4 on the device
    dict(tips=9, sites=300, states=4, rate_cats=5, model="GTR", scaling=True, rates=[0.1, 0.4, 0.8, 1.3, 2.4]),   # 5 -> 8
    dict(tips=20, sites=130, states=4, rate_cats=6, model="HKY", rates=list(np.linspace(0.2, 2.0, 6))),           # 6 -> 8, chunks
    dict(tips=6, sites=90, states=20, rate_cats=3, model="LG", rates=[0.3, 1.0, 1.7]),                            # 20 states, 3 -> 4
    dict(tips=5, sites=64, states=4, rate_cats=9, model="GTR", rates=list(np.linspace(0.2, 2.0, 9))),             # generic kernel
    dict(tips=33, sites=129, states=4, rate_cats=8, model="GTR", scaling=True, rates=list(np.linspace(0.1, 3, 8))),
    # big trees: serially planned with Sethi-Ullman ordering, several chunks, up to 16 packed tip words per cell
    dict(tips=60, sites=301, states=4, rate_cats=1, model="JC69"),                   # frogs-sized loci
    dict(tips=100, sites=140, states=4, rate_cats=4, model="GTR", scaling=True),
    dict(tips=129, sites=77, states=4, rate_cats=2, model="GTR", rates=[0.4, 1.6]),  # the last tip word has one tip
    dict(tips=200, sites=40, states=4, rate_cats=1, model="JC69", scaling=True),     # beyond every fast-path limit
])
def test_against_oracle_seeded(eng, cfg):
    w = synth.make_workload("seeded", n_loci=12, seed=4242, lg=lg_tables(), **cfg)
    loci, trees, batch = _load(eng, w)
    lnl, _ = batch.full_pass(trees.full_pass_step())
    if w.states == 4 and w.rate_cats in (1, 2, 4, 8) and 16 < w.tips <= 128:
        st = batch.plan_stats()            # big trees up to 128 tips: every locus on the fast path
        assert st["walker"] == 0 and st["fast"] == w.n_loci, st
    cm = char_map(w.states)
    step = 1 if w.states == 4 else 4
    for i in range(0, w.n_loci, step):
        ref = F.locus_from_workload(w, i, cm).full_pass()
        assert abs(lnl[i] - ref) <= LNL_RTOL * abs(ref), (i, lnl[i], ref)
    _free(loci, batch)


def test_ragged_batch_of_different_loci(eng):
    """Loci of one batch may differ in tips and sites (real data do, SURVEY 7 hard part 6)."""
    from bpp_b200 import engine
    shapes = [(4, 10), (9, 513), (2, 3), (17, 256), (5, 255), (8, 1)]
    loci, steps, refs = [], [], []
    cm = char_map(4)
    for k, (T, P) in enumerate(shapes):
        w = synth.make_workload("rag%d" % k, n_loci=1, tips=T, sites=P, states=4, rate_cats=4, model="GTR",
                                scaling=True, seed=100 + k)
        ls, tr = engine.load_workload(eng, w)
        loci += ls
        steps.append(tr.full_pass_step())
        refs.append(F.locus_from_workload(w, 0, cm).full_pass())
    batch = engine.Batch(eng, loci)
    step = tuple(np.concatenate([s[j] for s in steps]) for j in range(7))
    lnl, total = batch.full_pass(step)
    assert rel_err(lnl, refs) <= LNL_RTOL
    assert abs(total - float(np.sum(refs))) <= LNL_RTOL * abs(total)
    _free(loci, batch)


def test_partial_update_root_path(eng):
    """gene-tree age move (gtree.c:5437-5467): change one node age, update 2-3 P-matrices and the
    CLVs on the path node -> root only; unchanged siblings are read from HBM."""
    from bpp_b200 import engine
    w = synth.make_workload("path", n_loci=5, tips=12, sites=77, states=4, rate_cats=4, model="GTR",
                            scaling=True, seed=77)
    loci, trees, batch = _load(eng, w)
    batch.full_pass(trees.full_pass_step())
    cm = char_map(4)
    T = w.tips
    for i, l in enumerate(loci):
        o = F.locus_from_workload(w, i, cm)
        o.full_pass()
        node = T + 1 + (i % (T - 3))
        kids = [o.left[node - T], o.right[node - T]]
        lo = max(o.times[kids[0]], o.times[kids[1]])
        hi = o.times[o.parent[node]]
        newt = lo + 0.37 * (hi - lo)
        o.times[node] = newt
        trees.times[i, node] = newt
        touched = kids + [node]
        path = []
        n = node
        while n >= 0:
            path.append(n)
            n = o.parent[n]
        # oracle side
        for n in touched:
            o.flip_pmatrix(n)
        o.update_matrices(touched)
        for n in path:
            o.flip_clv(n)
        o.update_partials(path)
        ref = o.root_loglikelihood()
        # device side: same index arithmetic on the host, indices passed down
        e2 = 2 * T - 2
        for n in touched:
            trees.pmatrix_index[i, n] = (e2 + trees.pmatrix_index[i, n]) % (2 * e2)
        bl = [(trees.times[i, trees.parent[i, n]] - trees.times[i, n]) * trees.rate_mui[i] for n in touched]
        l.update_matrices([trees.pmatrix_index[i, n] for n in touched], bl)
        for n in path:
            trees.clv_index[i, n] = T + (trees.clv_index[i, n] - 1) % (2 * T - 2)
            trees.scaler_index[i, n] = (T + trees.scaler_index[i, n] - 1) % (2 * T - 2)
        ops = np.zeros(len(path), dtype=engine.OP_DTYPE)
        for k, n in enumerate(path):
            a, b = trees.left[i, n - T], trees.right[i, n - T]
            ops[k] = (trees.clv_index[i, n], trees.clv_index[i, a], trees.clv_index[i, b],
                      trees.pmatrix_index[i, a], trees.pmatrix_index[i, b],
                      trees.scaler_index[i, n], trees.scaler_index[i, a], trees.scaler_index[i, b])
        l.update_partials(ops)
        got = l.root_loglikelihood(int(trees.clv_index[i, trees.root]), int(trees.scaler_index[i, trees.root]))
        assert abs(got - ref) <= LNL_RTOL * abs(ref), (i, got, ref)
    _free(loci, batch)


def test_dense_tip_clv_and_likelihood_vector_and_diploid(eng):
    """pll_set_tip_clv with non-0/1 values (locus.c:596), pll_core_root_likelihood_vector
    (core_likelihood.c:214) and the diploid phase-mean branch (locus.c:2586-2615)."""
    from bpp_b200 import engine
    rng = np.random.default_rng(3)
    T, P, R = 6, 41, 4
    w = synth.make_workload("dense", n_loci=1, tips=T, sites=P, states=4, rate_cats=R, model="GTR", seed=31)
    loci, trees, batch = _load(eng, w)
    l = loci[0]
    o = F.locus_from_workload(w, 0, char_map(4))
    vals = rng.uniform(0.05, 1.0, size=(P, 4))
    l.set_tip_clv(2, vals)
    o.set_tip_values(2, vals)
    onehot = np.eye(4)[rng.integers(0, 4, size=P)]
    l.set_tip_clv(4, onehot)                      # 0/1 values are re-packed
    o.set_tip_values(4, onehot)
    lnl, _ = batch.full_pass(trees.full_pass_step())
    ref = o.full_pass()
    assert abs(lnl[0] - ref) <= LNL_RTOL * abs(ref)
    assert np.array_equal(l.get_clv(2).reshape(P, R, 4)[:, 1, :], vals)
    lh = l.root_likelihood_vector(int(trees.clv_index[0, trees.root]))
    assert rel_err(lh, o.root_likelihood_vector()) <= 1e-13
    # diploid: unphased sites map to 1, 2 or 4 phase resolutions
    counts, mapping = [], []
    k = 0
    while k < P:
        c = min([1, 2, 4][len(counts) % 3], P - k)
        counts.append(c)
        mapping += list(range(k, k + c))
        k += c
    uw = rng.integers(1, 4, size=len(counts)).astype(np.uint32)
    l.set_diploid(counts, mapping, uw)
    got = l.root_loglikelihood_diploid(int(trees.clv_index[0, trees.root]))
    want = F.diploid_loglikelihood(o.root_likelihood_vector(), counts, mapping, uw)
    assert abs(got - want) <= LNL_RTOL * abs(want)
    _free(loci, batch)


def test_zero_branch_and_minus_infinity(eng):
    """bl < 1e-100 -> identity P (core_pmatrix.c:738-743); incompatible tips across zero-length
    branches give term == 0 -> lnL = -inf like the reference (method.c:4302)."""
    from bpp_b200 import engine
    cm = char_map(4)
    l = engine.Locus.create_like_bpp(eng, 2, 3, 4, 1, False)
    l.set_tip_states(0, cm, b"AAC")
    l.set_tip_states(1, cm, b"AGC")
    l.set_frequencies(np.full(4, 0.25))
    l.update_matrices([0, 1], [0.0, 1e-101])
    assert np.array_equal(l.get_pmatrix(0), np.eye(4).ravel())
    assert np.array_equal(l.get_pmatrix(1), np.eye(4).ravel())
    ops = np.zeros(1, dtype=engine.OP_DTYPE)
    ops[0] = (2, 0, 1, 0, 1, -1, -1, -1)
    l.update_partials(ops)
    v, persite = l.root_loglikelihood(2, -1, persite=True)
    assert v == -np.inf and np.isfinite(persite[0]) and persite[1] == -np.inf
    l.destroy()


def test_illegal_state_code_is_fatal(eng):
    from bpp_b200 import engine
    l = engine.Locus.create_like_bpp(eng, 2, 4, 4, 1, False)
    with pytest.raises(engine.BppGpuError, match="Illegal state code"):
        l.set_tip_states(0, char_map(4), b"AC!T")
    l.destroy()


@pytest.mark.parametrize("scaling,tips,kind", [(False, 8, "all_paths"), (True, 8, "scaled_only"), (True, 24, "all_paths")])
def test_staged_run_is_idempotent_and_deterministic(eng, scaling, tips, kind):
    """stage once, run several times: identical bits (fixed-order reductions, no float atomics).  The first runs
    launch the kernel that carries every path; once the class of the cached plan is known (all loci scaled one-chunk
    lists) the later runs launch the specialised instantiation, which must reproduce the first run's bits; unscaled
    batches and batches of multi-chunk trees (24 tips) stay on the general kernel."""
    w = synth.make_workload("idem", n_loci=40, tips=tips, sites=300, states=4, rate_cats=4, model="GTR", seed=9,
                            scaling=scaling)
    loci, trees, batch = _load(eng, w)
    batch.stage(trees.full_pass_step())
    batch.run()
    a, ta = batch.collect()
    assert batch.kernel_name.endswith(",all_paths>")
    for _ in range(3):
        batch.run()
        b, tb = batch.collect()
        assert np.array_equal(a, b) and ta == tb
    assert batch.kernel_name.endswith(",%s>" % kind), batch.kernel_name
    # runs queued back to back without a collect in between (the class of the plan may still be on its way when the
    # next run is issued: "not ready" must not surface as an error), on a freshly staged copy of the same step
    batch.stage(trees.full_pass_step())
    for _ in range(6):
        batch.run()
    b, tb = batch.collect()
    assert np.array_equal(a, b) and ta == tb
    cm = char_map(4)
    for i in range(0, w.n_loci, 8):
        ref = F.locus_from_workload(w, i, cm).full_pass()
        assert abs(b[i] - ref) <= LNL_RTOL * abs(ref), (i, b[i], ref)
    _free(loci, batch)


def test_linearity_in_pattern_weights_full_size_property(eng):
    """Size-independent property at a BASELINE-sized locus count per launch: lnL is linear in the
    pattern weights, so doubling every weight doubles every per-locus lnL exactly (power of two)."""
    w = synth.make_workload("lin", n_loci=300, tips=8, sites=1000, states=4, rate_cats=1, model="JC69", seed=21)
    loci, trees, batch = _load(eng, w)
    a, _ = batch.full_pass(trees.full_pass_step())
    for i, l in enumerate(loci):
        l.set_pattern_weights(w.weights[i] * 2)
    b, _ = batch.full_pass(trees.full_pass_step())
    assert np.array_equal(b, 2 * a)
    _free(loci, batch)


def test_frogs_real_data_diploid(eng):
    """BASELINE.json config 1 (frogs A00): 5 real, ragged, diploid loci (42-60 sequences, 22-102
    patterns) through update_matrices -> update_partials -> likelihood vector -> device-side phase
    mean.  log-L0 = -7320.932289 as printed by the reference."""
    from bpp_b200 import engine
    from helpers import frogs_fixture
    d = frogs_fixture()
    total = 0.0
    for k in range(int(d["n_loci"])):
        p = "l%d_" % k
        T, P, S, R = [int(x) for x in d[p + "dims"][:4]]
        l = engine.Locus.create_like_bpp(eng, T, P, S, R, False, engine.DNA_MODEL_JC69)
        for t in range(T):
            l.set_tip_clv(t, ((d[p + "tip_masks"][t][:, None] >> np.arange(4)) & 1).astype(np.float64))
        l.set_frequencies(d[p + "freqs"])
        l.set_category_rates(d[p + "rates"])
        l.set_diploid(d[p + "resolution_count"], d[p + "mapping"], d[p + "weights"])
        nodes, tl = d[p + "nodes"], d[p + "time_length"]
        by_idx = {int(r[0]): (r, t) for r, t in zip(nodes, tl)}
        edges = [i for i in by_idx if by_idx[i][0][3] >= 0]
        l.update_matrices([int(by_idx[i][0][6]) for i in edges], [by_idx[i][1][1] for i in edges])
        # post-order = reversed pre-order of the inner nodes (children before parents)
        inner = [int(r[0]) for r in nodes if r[1] >= 0][::-1]
        ops = np.zeros(len(inner), dtype=engine.OP_DTYPE)
        for j, i in enumerate(inner):
            r = by_idx[i][0]
            a, b = by_idx[int(r[1])][0], by_idx[int(r[2])][0]
            ops[j] = (r[4], a[4], b[4], a[6], b[6], -1, -1, -1)
        l.update_partials(ops)
        root = int(nodes[0][4])
        assert rel_err(l.get_clv(root), d[p + "root_clv"]) < 1e-10
        lh = l.root_likelihood_vector(root)
        assert np.allclose(lh, d[p + "likelihood_vector"], rtol=1e-10, atol=0)
        lnl = l.root_loglikelihood_diploid(root)
        assert abs(lnl - float(d[p + "logl"])) <= LNL_RTOL * abs(lnl)
        total += lnl
        l.destroy()
    assert abs(total - (-7320.932289)) < 5e-6


def test_frogs_diploid_loci_in_one_batch(eng):
    """The same 5 diploid loci as ONE batch (the callers' `for each locus` loop as a single launch): the batch's
    root evaluation must apply the phase-resolution mean (locus.c:2586-2615) to every locus that carries a diploid
    mapping -- diploid_batch_kernel between the tree kernel and finish_kernel."""
    from bpp_b200 import engine
    from helpers import frogs_fixture
    d = frogs_fixture()
    loci, mc, mi, mb, oc, opl, rc, want = [], [], [], [], [], [], [], []
    for k in range(int(d["n_loci"])):
        p = "l%d_" % k
        T, P, S, R = [int(x) for x in d[p + "dims"][:4]]
        l = engine.Locus.create_like_bpp(eng, T, P, S, R, False, engine.DNA_MODEL_JC69)
        for t in range(T):
            l.set_tip_clv(t, ((d[p + "tip_masks"][t][:, None] >> np.arange(4)) & 1).astype(np.float64))
        l.set_frequencies(d[p + "freqs"])
        l.set_category_rates(d[p + "rates"])
        l.set_diploid(d[p + "resolution_count"], d[p + "mapping"], d[p + "weights"])
        nodes, tl = d[p + "nodes"], d[p + "time_length"]
        by_idx = {int(r[0]): (r, t) for r, t in zip(nodes, tl)}
        edges = [i for i in by_idx if by_idx[i][0][3] >= 0]
        mc.append(len(edges))
        mi += [int(by_idx[i][0][6]) for i in edges]
        mb += [by_idx[i][1][1] for i in edges]
        inner = [int(r[0]) for r in nodes if r[1] >= 0][::-1]
        ops = np.zeros(len(inner), dtype=engine.OP_DTYPE)
        for j, i in enumerate(inner):
            r = by_idx[i][0]
            a, b = by_idx[int(r[1])][0], by_idx[int(r[2])][0]
            ops[j] = (r[4], a[4], b[4], a[6], b[6], -1, -1, -1)
        oc.append(len(inner))
        opl.append(ops)
        rc.append(int(nodes[0][4]))
        want.append(float(d[p + "logl"]))
        loci.append(l)
    batch = engine.Batch(eng, loci)
    step = (np.array(mc), np.array(mi), np.array(mb), np.array(oc), np.concatenate(opl), np.array(rc),
            np.full(len(loci), -1))
    lnl, total = batch.full_pass(step)
    assert rel_err(lnl, np.array(want)) <= LNL_RTOL
    assert abs(total - (-7320.932289)) < 5e-6
    # real data (42-60 sequences per locus) stays on the chunk-by-chunk fast path, no locus on the cell-at-a-time walker
    st = batch.plan_stats()
    assert st["fast"] == len(loci) and st["walker"] == 0, st
    # and the split form: root-only evaluation of resident CLVs
    assert np.array_equal(batch.root_loglikelihood(np.array(rc)), lnl)
    _free(loci, batch)


def test_pinned_step_inputs_give_identical_results(eng):
    """Step arrays kept in pinned memory (bppgpu_host_alloc) go to the device without the staging
    copy; the result must be bit-identical to the pageable path."""
    from bpp_b200 import engine
    w = synth.make_workload("pin", n_loci=30, tips=9, sites=301, states=4, rate_cats=4, model="GTR", scaling=True, seed=5)
    loci, trees, batch = _load(eng, w)
    step = trees.full_pass_step()
    a, ta = batch.full_pass(step)
    pstep, holders = engine.pin_step(step)
    b, tb = batch.full_pass(pstep)
    assert np.array_equal(a, b) and ta == tb
    for h in holders:
        h.free()
    _free(loci, batch)


@pytest.mark.parametrize("waves", [2, 3, 8])
def test_pipelined_waves_give_identical_results(eng, waves):
    """A full pass uploaded and launched in waves of loci (copy stream + two compute streams) must give the
    same bits as the single-launch path, for ragged loci, with scaling, and when the same staged inputs are
    run again (then without waves), and when the step changes between stages (offset tables re-uploaded)."""
    from bpp_b200 import engine
    shapes = [(4, 10), (9, 513), (2, 3), (17, 256), (5, 255), (8, 1), (8, 1000), (3, 700), (12, 129), (6, 64), (16, 2048)]
    loci, steps, all_trees = [], [], []
    for k, (T, P) in enumerate(shapes):
        w = synth.make_workload("wav%d" % k, n_loci=3, tips=T, sites=P, states=4, rate_cats=4, model="GTR",
                                scaling=True, seed=300 + k)
        ls, tr = engine.load_workload(eng, w)
        loci += ls
        all_trees.append(tr)
        steps.append(tr.full_pass_step())
    batch = engine.Batch(eng, loci)
    step = tuple(np.concatenate([s[j] for s in steps]) for j in range(7))
    batch.set_waves(1)
    a, ta = batch.full_pass(step)
    batch.set_waves(waves)
    b, tb = batch.full_pass(step)
    assert np.array_equal(a, b) and ta == tb
    pstep, holders = engine.pin_step(step)
    prep = batch.prepare(pstep)
    batch.stage(prep)
    batch.run()
    c, tc = batch.collect()
    batch.run()                      # inputs now resident: single launch
    d, td = batch.collect()
    assert np.array_equal(a, c) and np.array_equal(a, d) and ta == tc == td
    # a different step shape through the same batch: only a root-path update for some loci
    short = list(step)
    mc, oc = step[0].copy(), step[3].copy()
    moff = np.concatenate([[0], np.cumsum(mc)]).astype(np.int64)
    keep_m = np.ones(int(mc.sum()), bool)
    for i in range(0, len(loci), 2):          # drop the matrices of every other locus (ops still recompute all)
        keep_m[moff[i]:moff[i + 1]] = False
        mc[i] = 0
    short[0], short[1], short[2] = mc, step[1][keep_m], step[2][keep_m]
    e1, te1 = batch.full_pass(tuple(short))
    batch.set_waves(1)
    e2, te2 = batch.full_pass(tuple(short))
    assert np.array_equal(e1, e2) and np.array_equal(e1, a) and te1 == te2
    for h in holders:
        h.free()
    _free(loci, batch)


def test_concurrent_host_threads_on_disjoint_batches(eng):
    """threads.c:87-200 calls the locus seam from up to opt_threads pthreads on disjoint loci without locks;
    the ABI promises the same.  Four host threads, each with its own loci and batch on the same engine, run
    full passes concurrently; every result must equal the single-threaded one bit for bit."""
    import threading
    from bpp_b200 import engine
    ws = [synth.make_workload("thr%d" % k, n_loci=40, tips=6 + k, sites=300 + 17 * k, states=4, rate_cats=4,
                              model="GTR", scaling=bool(k & 1), seed=700 + k) for k in range(4)]
    serial = []
    for w in ws:
        loci, trees, batch = _load(eng, w)
        serial.append(batch.full_pass(trees.full_pass_step())[0])
        _free(loci, batch)
    out, errs = [None] * 4, []

    def work(k):
        try:
            loci, trees = engine.load_workload(eng, ws[k])
            batch = engine.Batch(eng, loci)
            step = trees.full_pass_step()
            r = None
            for _ in range(20):
                r = batch.full_pass(step)[0]
            out[k] = r
            batch.destroy()
            for l in loci:
                l.destroy()
        except Exception as ex:          # noqa: BLE001
            errs.append(ex)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for k in range(4):
        assert np.array_equal(out[k], serial[k]), k


@pytest.mark.parametrize("cfg", [dict(tips=9, sites=120, states=4, rate_cats=4, model="GTR"),
                                 dict(tips=6, sites=40, states=20, rate_cats=4, model="LG"),
                                 dict(tips=7, sites=90, states=4, rate_cats=4, model="TN93")])
def test_batched_model_update_and_device_eigen(eng, cfg, monkeypatch):
    """Model changes of many loci (propose_alpha, propose_qrates / propose_freqs) travel in one blob and
    pll_update_eigen runs on the device (model_update_kernel); a batch below the threshold takes the per-locus
    path with the host-side decomposition.  Both must agree with each other to rounding and with the oracle to
    the lnL bar, also after a second round of changes (new rates, new Q) and for bppgpu_get_eigen."""
    w = synth.make_workload("mdl", n_loci=10, seed=4711, lg=lg_tables(), **cfg)
    cm = char_map(w.states)
    new_rates = np.array([0.1, 0.5, 1.1, 2.3])
    res = {}
    for mode, thresh in (("host", "1000000"), ("device", "1")):
        monkeypatch.setenv("BPPGPU_MODEL_BATCH_MIN", thresh)
        loci, trees, batch = _load(eng, w)
        step = trees.full_pass_step()
        a, _ = batch.full_pass(step)
        eig = loci[2].get_eigen() if w.model in ("GTR", "LG") else None
        pm = loci[1].get_pmatrix(int(step[1][0]))
        # second round: every locus gets new category rates, every other one new exchangeabilities / frequencies
        for i, l in enumerate(loci):
            l.set_category_rates(new_rates)
            if i % 2 == 0 and w.model != "LG":
                l.set_subst_params(w.subst[i][::-1].copy())
                l.set_frequencies(w.freqs[(i + 1) % w.n_loci])
        b, _ = batch.full_pass(step)
        # third round through the OTHER path: a decomposition that lives on the device only must survive a
        # per-locus model upload (and a host-side one a batched upload)
        monkeypatch.setenv("BPPGPU_MODEL_BATCH_MIN", "1" if mode == "host" else "1000000")
        for l in loci[:3]:
            l.set_category_rates(new_rates[::-1].copy() * 0.9)
        c, _ = batch.full_pass(step)
        res[mode] = (a, b, eig, pm, c)
        _free(loci, batch)
    monkeypatch.delenv("BPPGPU_MODEL_BATCH_MIN", raising=False)
    assert rel_err(res["device"][4], res["host"][4]) < 1e-12
    assert np.all(res["device"][4][:3] != res["device"][1][:3]) and np.array_equal(res["device"][4][3:], res["device"][1][3:])
    assert rel_err(res["device"][0], res["host"][0]) < 1e-12 and rel_err(res["device"][1], res["host"][1]) < 1e-12
    assert np.allclose(res["device"][3], res["host"][3], rtol=1e-12, atol=1e-15)
    if res["device"][2] is not None:
        ev, iev, lam = res["device"][2]
        S = w.states
        assert np.allclose(iev.reshape(S, S) @ ev.reshape(S, S), np.eye(S), atol=1e-12)
        assert np.allclose(np.sort(lam), np.sort(res["host"][2][2]), rtol=1e-10, atol=1e-13)
    for i in range(0, w.n_loci, 3):
        o = F.locus_from_workload(w, i, cm)
        ref = o.full_pass()
        assert abs(res["device"][0][i] - ref) <= LNL_RTOL * abs(ref)
        o.set_model(rates=new_rates)
        if i % 2 == 0 and w.model != "LG":
            o.set_model(subst=w.subst[i][::-1].copy(), freqs=w.freqs[(i + 1) % w.n_loci])
        ref2 = o.full_pass()
        assert abs(res["device"][1][i] - ref2) <= LNL_RTOL * abs(ref2)


def test_batches_of_different_shapes_alternate(eng):
    """Two live batches whose tree kernels need different amounts of shared memory, used in turn: the
    function-level shared-memory limit must not be lowered by one batch under the other."""
    wa = synth.make_workload("alt_a", n_loci=20, tips=16, sites=600, states=4, rate_cats=4, model="GTR", seed=81)
    wb = synth.make_workload("alt_b", n_loci=20, tips=4, sites=600, states=4, rate_cats=4, model="GTR", seed=82)
    la, ta, ba = _load(eng, wa)
    lb, tb, bb = _load(eng, wb)
    sa, sb = ta.full_pass_step(), tb.full_pass_step()
    ra0, rb0 = ba.full_pass(sa)[0], bb.full_pass(sb)[0]
    for _ in range(3):
        assert np.array_equal(ba.full_pass(sa)[0], ra0)
        assert np.array_equal(bb.full_pass(sb)[0], rb0)
    _free(la, ba)
    _free(lb, bb)


@pytest.mark.parametrize("cfg", [dict(tips=8, sites=300, states=4, rate_cats=4, model="HKY", scaling=True),
                                 dict(tips=5, sites=64, states=20, rate_cats=4, model="LG")])
def test_batch_split_calls_equal_full_pass(eng, cfg):
    """The reference's three calls in batch form -- update_matrices, update_partials, root_loglikelihood -- give
    what the fused full pass gives; loci with an empty op / matrix list in a batch call are left untouched."""
    w = synth.make_workload("split", n_loci=9, seed=77, lg=lg_tables(), **cfg)
    loci, trees, batch = _load(eng, w)
    mc, mi, mb, oc, ops, rc, rs = trees.full_pass_step()
    full, tot = batch.full_pass((mc, mi, mb, oc, ops, rc, rs))
    clv_full = loci[5].get_clv(int(rc[5]))
    loci2, trees2, batch2 = _load(eng, w)
    batch2.update_matrices(mc, mi, mb)
    batch2.update_partials(oc, ops)
    split = batch2.root_loglikelihood(rc, rs)
    assert np.array_equal(full, split)
    assert np.array_equal(clv_full, loci2[5].get_clv(int(rc[5])))
    # second round on the even loci only (new branch lengths), the odd ones keep their values
    moff = np.concatenate([[0], np.cumsum(mc)]).astype(np.int64)
    ooff = np.concatenate([[0], np.cumsum(oc)]).astype(np.int64)
    km, ko = np.zeros(len(mi), bool), np.zeros(len(ops), bool)
    mc2, oc2 = mc.copy(), oc.copy()
    for i in range(w.n_loci):
        if i % 2 == 0:
            km[moff[i]:moff[i + 1]] = True
            ko[ooff[i]:ooff[i + 1]] = True
        else:
            mc2[i] = 0
            oc2[i] = 0
    batch2.update_matrices(mc2, mi[km], mb[km] * 1.3)
    batch2.update_partials(oc2, ops[ko])
    second = batch2.root_loglikelihood(rc, rs)
    assert np.array_equal(second[1::2], full[1::2])
    assert np.all(second[0::2] != full[0::2])
    cm = char_map(w.states)
    o = F.locus_from_workload(w, 2, cm)
    o.rate_mui = float(w.rate_mui[2]) * 1.3
    ref = o.full_pass()
    assert abs(second[2] - ref) <= LNL_RTOL * abs(ref)
    _free(loci, batch)
    _free(loci2, batch2)


@pytest.mark.parametrize("tips,rate_cats,scaling", [(17, 4, False), (24, 4, True), (32, 1, False), (32, 4, True),
                                                    (29, 2, False), (20, 8, True)])
def test_multi_chunk_fast_path(eng, tips, rate_cats, scaling):
    """Trees of 17..32 tips need more than one staged chunk (16 ops, or the lookup-table capacity) and tip words
    beyond the first two; they run the fast path chunk by chunk, X and its scaler carried in registers.  lnL
    against the oracle, every inner CLV against the oracle's (same arithmetic, P-matrices equal to rounding),
    scalers exactly; then a full pass in batch form after a root-path style change of all branch lengths."""
    rates = {1: None, 4: None, 2: [0.3, 1.7], 8: list(np.linspace(0.1, 3, 8))}[rate_cats]
    kw = dict(dt_lo=0.02, dt_hi=0.2) if scaling else {}
    w = synth.make_workload("mchunk", n_loci=6, tips=tips, sites=530, states=4, rate_cats=rate_cats, model="GTR",
                            scaling=scaling, seed=1000 + tips, rates=rates, **kw)
    loci, trees, batch = _load(eng, w)
    lnl, tot = batch.full_pass(trees.full_pass_step())
    cm = char_map(4)
    for i in (0, 3, 5):
        o = F.locus_from_workload(w, i, cm)
        ref = o.full_pass()
        assert abs(lnl[i] - ref) <= LNL_RTOL * abs(ref), (i, lnl[i], ref)
        for n in range(tips, 2 * tips - 1):
            assert rel_err(loci[i].get_clv(n), np.asarray(o.clv[o.clv_index[n]]).ravel()) < 1e-11, (i, n)
            if scaling:
                assert np.array_equal(loci[i].get_scaler(n - tips), o.scale[o.scaler_index[n]]), (i, n)
    # the same staged inputs again, and a second batch built the same way: bitwise reproducible
    again, tot2 = batch.full_pass(trees.full_pass_step())
    assert np.array_equal(lnl, again) and tot == tot2
    _free(loci, batch)


@pytest.mark.parametrize("rate_cats,scaling", [(4, True), (1, False)])
def test_partial_update_long_root_path_multi_chunk(eng, rate_cats, scaling):
    """A node-age move deep in a 30-tip tree: the root path can be longer than one chunk of 16 ops, its unchanged
    siblings are HBM-resident operands with staged matrices in every chunk (full instantiation of the fast path,
    chunk by chunk).  Batch form over all loci, against the oracle."""
    from bpp_b200 import engine
    T = 30
    w = synth.make_workload("lpath", n_loci=8, tips=T, sites=300, states=4, rate_cats=rate_cats, model="GTR",
                            scaling=scaling, seed=555)
    # caterpillar trees (random coalescent trees of 30 tips have root paths of ~10 ops only): inner node k joins
    # inner node k-1 and tip k+1, so the path from the first cherry to the root has all 29 ops
    rng = np.random.default_rng(9)
    for i in range(w.n_loci):
        for k in range(T - 1):
            w.left[i, k] = T + k - 1 if k > 0 else 0
            w.right[i, k] = k + 1
        w.times[i, :T] = 0.0
        w.times[i, T:] = np.cumsum(rng.uniform(0.002, 0.02, size=T - 1))
    loci, trees, batch = _load(eng, w)
    batch.full_pass(trees.full_pass_step())
    cm = char_map(4)
    mcs, mis, mbs, ocs, opss, refs = [], [], [], [], [], []
    longest = 0
    for i in range(w.n_loci):
        o = F.locus_from_workload(w, i, cm)
        o.full_pass()
        # the inner node with the longest path to the root
        depth = {}
        for n in range(T, 2 * T - 1):
            d, m = 0, n
            while o.parent[m] >= 0:
                m = o.parent[m]
                d += 1
            depth[n] = d
        node = max(depth, key=lambda n: depth[n])
        kids = [o.left[node - T], o.right[node - T]]
        lo = max(o.times[kids[0]], o.times[kids[1]])
        hi = o.times[o.parent[node]]
        newt = lo + 0.41 * (hi - lo)
        o.times[node] = newt
        trees.times[i, node] = newt
        touched = kids + [node]
        path = []
        n = node
        while n >= 0:
            path.append(n)
            n = o.parent[n]
        longest = max(longest, len(path))
        for n in touched:
            o.flip_pmatrix(n)
        o.update_matrices(touched)
        for n in path:
            o.flip_clv(n)
        o.update_partials(path)
        refs.append(o.root_loglikelihood())
        e2 = 2 * T - 2
        for n in touched:
            trees.pmatrix_index[i, n] = (e2 + trees.pmatrix_index[i, n]) % (2 * e2)
        mcs.append(len(touched))
        mis += [trees.pmatrix_index[i, n] for n in touched]
        mbs += [(trees.times[i, trees.parent[i, n]] - trees.times[i, n]) * trees.rate_mui[i] for n in touched]
        for n in path:
            trees.clv_index[i, n] = T + (trees.clv_index[i, n] - 1) % (2 * T - 2)
            if scaling:
                trees.scaler_index[i, n] = (T + trees.scaler_index[i, n] - 1) % (2 * T - 2)
        ops = np.zeros(len(path), dtype=engine.OP_DTYPE)
        for k, n in enumerate(path):
            a, b = trees.left[i, n - T], trees.right[i, n - T]
            ops[k] = (trees.clv_index[i, n], trees.clv_index[i, a], trees.clv_index[i, b],
                      trees.pmatrix_index[i, a], trees.pmatrix_index[i, b],
                      trees.scaler_index[i, n], trees.scaler_index[i, a], trees.scaler_index[i, b])
        ocs.append(len(path))
        opss.append(ops)
    rc = np.array([trees.clv_index[i, trees.root] for i in range(w.n_loci)], dtype=np.uint32)
    rs = np.array([trees.scaler_index[i, trees.root] for i in range(w.n_loci)], dtype=np.int32)
    got, _ = batch.full_pass((np.array(mcs, dtype=np.uint32), np.array(mis, dtype=np.uint32), np.array(mbs),
                              np.array(ocs, dtype=np.uint32), np.concatenate(opss), rc, rs))
    assert longest == T - 1
    assert rel_err(got, refs) <= LNL_RTOL, (got, refs, longest)
    _free(loci, batch)


@pytest.mark.parametrize("repeat", [2, 4, 6])
def test_op_list_longer_than_the_tree(eng, repeat):
    """locus_update_partials accepts any node list; one that visits every inner node several times has more tip
    children than lookup-table slots and more ops than a chunk, so it takes many chunks (the blocks are sized
    from the launch's slot capacity).  Recomputing a node is idempotent: same lnL as the plain pass."""
    w = synth.make_workload("long", n_loci=7, tips=8, sites=200, states=4, rate_cats=1, model="JC69", seed=66)
    loci, trees, batch = _load(eng, w)
    mc, mi, mb, oc, ops, rc, rs = trees.full_pass_step()
    plain, _ = batch.full_pass((mc, mi, mb, oc, ops, rc, rs))
    ooff = np.concatenate([[0], np.cumsum(oc)]).astype(np.int64)
    rep = np.concatenate([np.tile(ops[ooff[i]:ooff[i + 1]], repeat) for i in range(w.n_loci)])
    again, _ = batch.full_pass((mc, mi, mb, oc * repeat, rep, rc, rs))
    assert np.array_equal(plain, again)
    _free(loci, batch)
