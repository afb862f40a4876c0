"""The C host-side mirror of the reference's locus seam (bpp_b200/host/), driven through ctypes.

CPU: discrete-Gamma rates (bppgpu_compute_gamma_cats, gamma.c:221) against values printed by the
reference itself (tests/golden/gamma_rates.npz).  GPU: the reference's call sequence
locus_create -> pll_set_* -> locus_update_matrices -> locus_update_partials ->
locus_root_loglikelihood, and the batched mixing-move pass, against the reference fixtures.
"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import GOLDEN, char_map, load_case, rel_err
from bpp_b200 import build


class GNode(C.Structure):
    pass


GNode._fields_ = [("left", C.POINTER(GNode)), ("right", C.POINTER(GNode)), ("parent", C.POINTER(GNode)),
                  ("length", C.c_double), ("time", C.c_double), ("node_index", C.c_uint), ("clv_index", C.c_uint),
                  ("scaler_index", C.c_int), ("pmatrix_index", C.c_uint), ("pop", C.c_int)]


class STree(C.Structure):
    _fields_ = [("node_count", C.c_uint), ("parent", C.POINTER(C.c_int)), ("tau", C.POINTER(C.c_double)),
                ("brate", C.POINTER(C.c_double))]


class GTree(C.Structure):
    _fields_ = [("tip_count", C.c_uint), ("inner_count", C.c_uint), ("edge_count", C.c_uint),
                ("nodes", C.POINTER(C.POINTER(GNode))), ("root", C.POINTER(GNode)), ("rate_mui", C.c_double),
                ("logl", C.c_double), ("stree", C.POINTER(STree)), ("rate_scale", C.c_double)]


@pytest.fixture(scope="module")
def host():
    build.build_cuda()
    path = build.build_host()
    H = C.CDLL(path)
    H.bppgpu_compute_gamma_cats.argtypes = [C.c_double, C.c_double, C.c_uint, C.POINTER(C.c_double)]
    H.gtree_create_gpu.restype = C.POINTER(GTree)
    H.gtree_create_gpu.argtypes = [C.c_uint, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_double, C.c_int]
    H.gtree_destroy_gpu.argtypes = [C.POINTER(GTree)]
    H.gtree_all_partials_gpu.argtypes = [C.POINTER(GNode), C.POINTER(C.POINTER(GNode)), C.POINTER(C.c_uint)]
    H.locus_create_gpu.restype = C.c_void_p
    H.locus_create_gpu.argtypes = [C.c_void_p] + [C.c_uint] * 11
    H.locus_destroy_gpu.argtypes = [C.c_void_p]
    H.pll_set_tip_states_gpu.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_uint), C.c_char_p]
    H.pll_set_pattern_weights_gpu.argtypes = [C.c_void_p, C.POINTER(C.c_uint)]
    H.pll_set_frequencies_gpu.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_double)]
    H.pll_set_subst_params_gpu.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_double)]
    H.pll_set_category_rates_gpu.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    H.locus_update_matrices_gpu.argtypes = [C.c_void_p, C.POINTER(GTree), C.POINTER(C.POINTER(GNode)), C.c_uint]
    H.locus_update_partials_gpu.argtypes = [C.c_void_p, C.POINTER(C.POINTER(GNode)), C.c_uint]
    H.locus_root_loglikelihood_gpu.restype = C.c_double
    H.locus_root_loglikelihood_gpu.argtypes = [C.c_void_p, C.POINTER(GNode), C.POINTER(C.c_double)]
    H.locus_batch_create_gpu.restype = C.c_void_p
    H.locus_batch_create_gpu.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_uint]
    H.locus_batch_destroy_gpu.argtypes = [C.c_void_p]
    H.locus_batch_full_pass_gpu.restype = C.c_double
    H.locus_batch_full_pass_gpu.argtypes = [C.c_void_p, C.POINTER(C.POINTER(GTree)), C.POINTER(C.c_double)]
    return H


def test_gamma_rates_match_reference(host):
    d = np.load(os.path.join(GOLDEN, "gamma_rates.npz"))
    for key in d.files:
        a, k = key[1:].split("_k")
        alpha, cats = float(a), int(k)
        out = (C.c_double * cats)()
        assert host.bppgpu_compute_gamma_cats(alpha, alpha, cats, out) == 1
        got = np.array(out[:])
        assert np.allclose(got, d[key], rtol=1e-13, atol=0), (key, got, d[key])
        assert abs(got.mean() - 1.0) < 1e-6          # mean-one rates


def test_gamma_rates_used_by_synth(host):
    from bpp_b200 import synth
    out = (C.c_double * 4)()
    host.bppgpu_compute_gamma_cats(0.5, 0.5, 4, out)
    assert np.array_equal(np.array(out[:]), synth.GAMMA4_ALPHA_0_5)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_relaxed_clock_branch_lengths(host):
    """update_branchlength_relaxed_clock (locus.c:1146-1190): a branch that crosses species-tree nodes is the sum of
    (time span in a population) x (that population's rate); the strict clock is (parent time - time) x locus rate."""
    host.gtree_branch_length_gpu.restype = C.c_double
    host.gtree_branch_length_gpu.argtypes = [C.POINTER(GTree), C.POINTER(GNode)]
    host.gtree_set_relaxed_clock_gpu.argtypes = [C.POINTER(GTree), C.POINTER(STree), C.POINTER(C.c_int), C.c_double]
    # species tree ((A,B)AB,C)ABC: nodes A=0 B=1 C=2 AB=3 ABC=4
    parent = np.array([3, 3, 4, 4, -1], dtype=np.int32)
    tau = np.array([0.0, 0.0, 0.0, 0.010, 0.025])
    brate = np.array([0.8, 1.3, 0.9, 1.7, 0.6])
    st = STree(5, parent.ctypes.data_as(C.POINTER(C.c_int)), _dp(tau), _dp(brate))
    # gene tree of 4 tips (a1, a2 in A; b in B; c in C): ((a1,a2),b),c) with coalescences in A, AB and ABC
    left = np.array([0, 4, 5], dtype=np.int32)
    right = np.array([1, 2, 3], dtype=np.int32)
    times = np.array([0, 0, 0, 0, 0.004, 0.018, 0.040])
    pops = np.array([0, 0, 1, 2, 0, 3, 4], dtype=np.int32)
    t = host.gtree_create_gpu(4, left.ctypes.data_as(C.POINTER(C.c_int)), right.ctypes.data_as(C.POINTER(C.c_int)),
                              _dp(times), 0.7, 0)
    gpar = {0: 4, 1: 4, 4: 5, 2: 5, 5: 6, 3: 6}

    def expect(k, scale):
        tm, pop, end, length = times[k], pops[k], pops[gpar[k]], 0.0
        while pop != end:
            nxt = parent[pop]
            length += (tau[nxt] - tm) * brate[pop] * scale
            tm, pop = tau[nxt], nxt
        return length + (times[gpar[k]] - tm) * brate[end] * scale

    for k in gpar:                                   # strict clock
        assert host.gtree_branch_length_gpu(t, t.contents.nodes[k]) == (times[gpar[k]] - times[k]) * 0.7
    for scale in (1.0, 2.5):                         # relaxed clocks; BPP_CLOCK_SIMPLE multiplies by the locus rate
        host.gtree_set_relaxed_clock_gpu(t, C.byref(st), pops.ctypes.data_as(C.POINTER(C.c_int)), scale)
        for k in gpar:
            got = host.gtree_branch_length_gpu(t, t.contents.nodes[k])
            assert abs(got - expect(k, scale)) <= 1e-15, (k, got, expect(k, scale))
    # b (in B at time 0) joins at 0.018 in AB: 0.010 in B, 0.008 in AB
    assert abs(host.gtree_branch_length_gpu(t, t.contents.nodes[2]) - 2.5 * (0.010 * 1.3 + 0.008 * 1.7)) < 1e-15
    host.gtree_set_relaxed_clock_gpu(t, None, None, 1.0)
    assert host.gtree_branch_length_gpu(t, t.contents.nodes[2]) == (0.018 - 0.0) * 0.7
    host.gtree_destroy_gpu(t)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["jc69_r1", "gtr_g4_scale", "lg_g4"])
def test_reference_call_sequence_through_c_host(host, name):
    """method.c:4137-4297: create, set tips / weights / model, all matrices, all partials, root lnL;
    then a mixing-style step through the batch call with the SWAP_* index flips (prop_mixing.c:52-220)."""
    from bpp_b200 import engine
    w, d = load_case(name)
    T, P, S, R = w.tips, w.sites, w.states, w.rate_cats
    eng = engine.Engine(0)
    cm = np.ascontiguousarray(char_map(S), dtype=np.uint32)
    model = {"JC69": 0, "GTR": 7, "LG": 1}[w.model]
    loci, trees = [], []
    for i in range(w.n_loci):
        l = host.locus_create_gpu(eng.h, 0 if S == 4 else 1, model, T, 2 * (T - 1), S, P, 1, 2 * (2 * T - 2), R,
                                  2 * (T - 1) if w.scaling else 0, 1 << 6)
        assert l
        for t in range(T):
            assert host.pll_set_tip_states_gpu(l, t, cm.ctypes.data_as(C.POINTER(C.c_uint)), w.tip_chars[i, t].tobytes()) == 1
        wt = np.ascontiguousarray(w.weights[i], dtype=np.uint32)
        host.pll_set_pattern_weights_gpu(l, wt.ctypes.data_as(C.POINTER(C.c_uint)))
        host.pll_set_frequencies_gpu(l, 0, _dp(np.ascontiguousarray(w.freqs[i])))
        if w.model != "JC69":
            host.pll_set_subst_params_gpu(l, 0, _dp(np.ascontiguousarray(w.subst[i])))
        host.pll_set_category_rates_gpu(l, _dp(np.ascontiguousarray(w.rates)))
        left = np.ascontiguousarray(w.left[i], dtype=np.int32)
        right = np.ascontiguousarray(w.right[i], dtype=np.int32)
        times = np.ascontiguousarray(w.times[i], dtype=np.float64)
        gt = host.gtree_create_gpu(T, left.ctypes.data_as(C.POINTER(C.c_int)), right.ctypes.data_as(C.POINTER(C.c_int)),
                                   _dp(times), float(w.rate_mui[i]), int(w.scaling))
        loci.append(l)
        trees.append(gt)
    # per-locus synchronous seam
    nn = 2 * T - 1
    trav = (C.POINTER(GNode) * nn)()
    for i in range(w.n_loci):
        gt = trees[i].contents
        k = 0
        for j in range(nn):
            if gt.nodes[j].contents.parent:
                trav[k] = gt.nodes[j]
                k += 1
        host.locus_update_matrices_gpu(loci[i], trees[i], trav, k)
        cnt = C.c_uint(0)
        host.gtree_all_partials_gpu(gt.root, trav, C.byref(cnt))
        assert cnt.value == T - 1
        host.locus_update_partials_gpu(loci[i], trav, cnt.value)
        lnl = host.locus_root_loglikelihood_gpu(loci[i], gt.root, None)
        assert abs(lnl - d["lnl"][i]) <= 1e-10 * abs(lnl), (name, i)
    # batched mixing step
    arr = (C.c_void_p * w.n_loci)(*loci)
    batch = host.locus_batch_create_gpu(eng.h, arr, w.n_loci)
    c = float(d["mix_c"])
    e2 = 2 * T - 2
    for i in range(w.n_loci):
        gt = trees[i].contents
        for j in range(nn):
            node = gt.nodes[j].contents
            node.time *= c
            if node.parent:
                node.pmatrix_index = (e2 + node.pmatrix_index) % (2 * e2)          # SWAP_PMAT_INDEX
            if j >= T:
                node.clv_index = T + (node.clv_index - 1) % (2 * T - 2)            # SWAP_CLV_INDEX
                if w.scaling:
                    node.scaler_index = (T + node.scaler_index - 1) % (2 * T - 2)  # SWAP_SCALER_INDEX
    tarr = (C.POINTER(GTree) * w.n_loci)(*trees)
    out = np.zeros(w.n_loci)
    total = host.locus_batch_full_pass_gpu(batch, tarr, _dp(out))
    assert rel_err(out, d["lnl_mix"]) <= 1e-10
    assert abs(total - d["lnl_mix"].sum()) <= 1e-10 * abs(total)
    host.locus_batch_destroy_gpu(batch)
    for l, t in zip(loci, trees):
        host.locus_destroy_gpu(l)
        host.gtree_destroy_gpu(t)
    eng.close()
