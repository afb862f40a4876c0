"""ctypes binding of oracle/_ref/libbppref.so -- TEST INFRASTRUCTURE ONLY.

The library is the unmodified reference (bpp v4.8.7) compiled by oracle/Makefile
plus oracle/ref_shim.c.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this module; the product
package (bpp_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libbppref.so")

# attribute bits, bpp.h:364-374
ARCH_CPU, ARCH_SSE, ARCH_AVX, ARCH_AVX2 = 0, 1, 2, 4
# bpp.h:208-222
DATA_DNA, DATA_AA = 0, 1
MODEL_JC69, MODEL_GTR = 0, 7
# bpp.h:215-222
DNA_MODELS = {"JC69": 0, "K80": 1, "F81": 2, "HKY": 3, "T92": 4, "TN93": 5, "F84": 6, "GTR": 7}
AA_MODEL_LG = 1  # bpp.h BPP_AA_MODEL_LG; the shim overrides freqs/rates anyway

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        up = C.POINTER(C.c_uint)
        ip = C.POINTER(C.c_int)
        vp = C.c_void_p
        L.ref_set_create.restype = vp
        L.ref_set_create.argtypes = [C.c_int] * 7
        L.ref_set_destroy.argtypes = [vp]
        L.ref_locus_create.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.ref_set_tip_states.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
        L.ref_set_tip_clv.argtypes = [vp, C.c_int, C.c_int, dp]
        L.ref_set_weights.argtypes = [vp, C.c_int, up]
        L.ref_set_model.argtypes = [vp, C.c_int, dp, dp, dp]
        L.ref_gamma_rates.argtypes = [C.c_double, C.c_int, dp]
        L.ref_set_tree.argtypes = [vp, C.c_int, ip, ip, dp, C.c_double]
        L.ref_set_times.argtypes = [vp, C.c_int, dp]
        L.ref_node_get.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.ref_node_set.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_update_matrices.argtypes = [vp, C.c_int, C.c_int, ip]
        L.ref_update_partials.argtypes = [vp, C.c_int, C.c_int, ip]
        L.ref_root_loglikelihood.restype = C.c_double
        L.ref_root_loglikelihood.argtypes = [vp, C.c_int, dp]
        L.ref_set_diploid.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_ulong),
                                      C.POINTER(C.c_ulong), C.c_int]
        L.ref_full_pass.restype = C.c_double
        L.ref_full_pass.argtypes = [vp, C.c_int]
        L.ref_full_pass_all.restype = C.c_double
        L.ref_full_pass_all.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.ref_age_move.restype = C.c_double
        L.ref_age_move.argtypes = [vp, C.c_int, C.c_int, C.c_double]
        L.ref_age_move_all.restype = C.c_double
        L.ref_age_move_all.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, dp, dp]
        for name in ("ref_clv", "ref_pmatrix"):
            getattr(L, name).restype = dp
            getattr(L, name).argtypes = [vp, C.c_int, C.c_int]
        L.ref_scaler.restype = up
        L.ref_scaler.argtypes = [vp, C.c_int, C.c_int]
        for name in ("ref_eigenvecs", "ref_inv_eigenvecs", "ref_eigenvals", "ref_rates",
                     "ref_freqs", "ref_likelihood_vector"):
            getattr(L, name).restype = dp
            getattr(L, name).argtypes = [vp, C.c_int]
        L.ref_aa_rates_lg.restype = dp
        L.ref_aa_freqs_lg.restype = dp
        L.ref_map_nt.restype = up
        L.ref_map_aa.restype = up
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def gamma_rates(alpha, cats):
    out = np.zeros(cats)
    lib().ref_gamma_rates(alpha, cats, _d(out))
    return out


def aa_lg():
    L = lib()
    rates = np.ctypeslib.as_array(L.ref_aa_rates_lg(), (190,)).copy()
    freqs = np.ctypeslib.as_array(L.ref_aa_freqs_lg(), (20,)).copy()
    return rates, freqs


def char_maps():
    L = lib()
    return (np.ctypeslib.as_array(L.ref_map_nt(), (256,)).copy(),
            np.ctypeslib.as_array(L.ref_map_aa(), (256,)).copy())


class RefSet:
    """N loci driven through the reference's locus seam (one global rate_cats /
    scaling setting per set because the reference reads them from globals)."""

    def __init__(self, n_loci, states, rate_cats, scaling, model=None, arch=ARCH_AVX2):
        self.L = lib()
        self.n, self.S, self.R, self.scaling = n_loci, states, rate_cats, int(bool(scaling))
        dtype = DATA_DNA if states == 4 else DATA_AA
        if model is None:
            model = MODEL_JC69 if states == 4 else AA_MODEL_LG
        self.model = model
        self.h = self.L.ref_set_create(n_loci, dtype, model, states, rate_cats, self.scaling, arch)
        self.tips = [0] * n_loci
        self.sites = [0] * n_loci

    def close(self):
        if self.h:
            self.L.ref_set_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def create(self, i, tips, sites):
        self.tips[i], self.sites[i] = tips, sites
        self.L.ref_locus_create(self.h, i, tips, sites)

    def set_tip_states(self, i, tip, seq_bytes):
        assert len(seq_bytes) == self.sites[i]
        return self.L.ref_set_tip_states(self.h, i, tip, bytes(seq_bytes))

    def set_tip_clv(self, i, tip, clv):
        clv = np.ascontiguousarray(clv, dtype=np.float64)
        assert clv.size == self.sites[i] * self.S
        return self.L.ref_set_tip_clv(self.h, i, tip, _d(clv))

    def set_weights(self, i, w):
        w = np.ascontiguousarray(w, dtype=np.uint32)
        self.L.ref_set_weights(self.h, i, w.ctypes.data_as(C.POINTER(C.c_uint)))

    def set_model(self, i, freqs=None, subst=None, rates=None):
        f = None if freqs is None else np.ascontiguousarray(freqs, dtype=np.float64)
        s = None if subst is None else np.ascontiguousarray(subst, dtype=np.float64)
        r = None if rates is None else np.ascontiguousarray(rates, dtype=np.float64)
        self.L.ref_set_model(self.h, i, None if f is None else _d(f),
                             None if s is None else _d(s), None if r is None else _d(r))

    def set_tree(self, i, left, right, times, rate_mui=1.0):
        left = np.ascontiguousarray(left, dtype=np.int32)
        right = np.ascontiguousarray(right, dtype=np.int32)
        times = np.ascontiguousarray(times, dtype=np.float64)
        self.L.ref_set_tree(self.h, i, _i(left), _i(right), _d(times), rate_mui)

    def set_times(self, i, times):
        times = np.ascontiguousarray(times, dtype=np.float64)
        self.L.ref_set_times(self.h, i, _d(times))

    def node_get(self, i, node, field):
        return self.L.ref_node_get(self.h, i, node, field)

    def node_set(self, i, node, field, v):
        self.L.ref_node_set(self.h, i, node, field, v)

    def update_matrices(self, i, node_ids):
        a = np.ascontiguousarray(node_ids, dtype=np.int32)
        self.L.ref_update_matrices(self.h, i, len(a), _i(a))

    def update_partials(self, i, node_ids):
        a = np.ascontiguousarray(node_ids, dtype=np.int32)
        self.L.ref_update_partials(self.h, i, len(a), _i(a))

    def root_loglikelihood(self, i, persite=False):
        if persite:
            out = np.zeros(self.sites[i])
            v = self.L.ref_root_loglikelihood(self.h, i, _d(out))
            return v, out
        return self.L.ref_root_loglikelihood(self.h, i, None)

    def set_diploid(self, i, resolution_count, mapping):
        rc = np.ascontiguousarray(resolution_count, dtype=np.uint64)
        mp = np.ascontiguousarray(mapping, dtype=np.uint64)
        ul = C.POINTER(C.c_ulong)
        self.L.ref_set_diploid(self.h, i, len(rc), rc.ctypes.data_as(ul), mp.ctypes.data_as(ul), len(mp))

    def full_pass(self, i):
        return self.L.ref_full_pass(self.h, i)

    def full_pass_all(self, first, count, nthreads=1, passes=1):
        out = np.zeros(self.n)
        secs = self.L.ref_full_pass_all(self.h, first, count, nthreads, passes, _d(out))
        return secs, out

    def age_move_all(self, first, count, node_ids, new_ages, nthreads=1):
        """One gene-tree age move per locus (gtree.c:5437-5467, likelihood part); returns (seconds, lnl[n])."""
        nodes = np.ascontiguousarray(node_ids, dtype=np.int32)
        ages = np.ascontiguousarray(new_ages, dtype=np.float64)
        out = np.zeros(self.n)
        secs = self.L.ref_age_move_all(self.h, first, count, nthreads, _i(nodes), _d(ages), _d(out))
        return secs, out

    def clv(self, i, clv_index):
        n = self.sites[i] * self.R * self.S
        return np.ctypeslib.as_array(self.L.ref_clv(self.h, i, clv_index), (n,)).copy()

    def pmatrix(self, i, idx):
        n = self.R * self.S * self.S
        return np.ctypeslib.as_array(self.L.ref_pmatrix(self.h, i, idx), (n,)).copy()

    def scaler(self, i, idx):
        return np.ctypeslib.as_array(self.L.ref_scaler(self.h, i, idx), (self.sites[i],)).copy()

    def eigen(self, i):
        S = self.S
        return (np.ctypeslib.as_array(self.L.ref_eigenvecs(self.h, i), (S * S,)).copy(),
                np.ctypeslib.as_array(self.L.ref_inv_eigenvecs(self.h, i), (S * S,)).copy(),
                np.ctypeslib.as_array(self.L.ref_eigenvals(self.h, i), (S,)).copy())

    def rates(self, i):
        return np.ctypeslib.as_array(self.L.ref_rates(self.h, i), (self.R,)).copy()

    def likelihood_vector(self, i):
        return np.ctypeslib.as_array(self.L.ref_likelihood_vector(self.h, i), (self.sites[i],)).copy()
