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
                  ("scaler_index", C.c_int), ("pmatrix_index", C.c_uint)]


class GTree(C.Structure):
    _fields_ = [("tip_count", C.c_uint), ("inner_count", C.c_uint), ("edge_count", C.c_uint),
                ("nodes", C.POINTER(C.POINTER(GNode))), ("root", C.POINTER(GNode)), ("rate_mui", C.c_double),
                ("logl", C.c_double)]


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
