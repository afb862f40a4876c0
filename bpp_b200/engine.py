"""Host-side mirror of the reference's locus seam over the C-ABI (include/bpp_b200.h).

Names and argument meaning follow bpp v4.8.7 (src/bpp.h:2032-2090, src/locus.c):
locus_create, pll_set_tip_states, pll_set_tip_clv, pll_set_pattern_weights, pll_set_frequencies,
pll_set_subst_params, pll_set_category_rates, locus_update_matrices, locus_update_partials,
locus_root_loglikelihood -- plus `Batch`, the `for each locus` loop of the callers
(prop_mixing.c:71-214) turned into one launch, and `GeneTrees`, the gnode_t index bookkeeping
(bpp.h:715-717, flip macros locus.c:24-26) for N loci as numpy arrays.

Everything numerical happens in the CUDA library; this file only marshals arrays.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import PartialOp, BppGpuError  # noqa: F401

DATA_DNA, DATA_AA = 0, 1
DNA_MODEL_JC69, DNA_MODEL_GTR = 0, 7
# bpp.h:215-222
DNA_MODELS = {"JC69": 0, "K80": 1, "F81": 2, "HKY": 3, "T92": 4, "TN93": 5, "F84": 6, "GTR": 7}
SCALE_BUFFER_NONE = -1
MATH_EXACT, MATH_FMA = 0, 1
ATTRIB_ARCH_CUDA = 1 << 6
KERNELS = ("pmatrix", "plan", "tree", "finish")

OP_DTYPE = np.dtype([("parent_clv_index", "<u4"), ("left_clv_index", "<u4"), ("right_clv_index", "<u4"),
                     ("left_pmatrix_index", "<u4"), ("right_pmatrix_index", "<u4"),
                     ("parent_scaler_index", "<i4"), ("left_scaler_index", "<i4"), ("right_scaler_index", "<i4")])
assert OP_DTYPE.itemsize == C.sizeof(PartialOp) == 32


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _opp(a):
    return a.ctypes.data_as(C.POINTER(PartialOp))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class PinnedArray:
    """A numpy array backed by pinned host memory of the library (bppgpu_host_alloc): step inputs kept
    in such arrays are copied H2D without an intermediate staging copy."""

    def __init__(self, like):
        L = _lib.load()
        a = np.ascontiguousarray(like)
        self._L, self.ptr = L, L.bppgpu_host_alloc(max(1, a.nbytes))
        _lib.check()
        buf = (C.c_char * max(1, a.nbytes)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=a.dtype, count=a.size).reshape(a.shape)
        self.array[...] = a

    def free(self):
        if self.ptr:
            self.array = None
            self._L.bppgpu_host_free(self.ptr)
            self.ptr = None


def pin_step(step):
    """Copy the 7 arrays of a full-pass step into pinned memory; returns (pinned step tuple, holders)."""
    dt = [np.uint32, np.uint32, np.float64, np.uint32, OP_DTYPE, np.uint32, np.int32]
    holders = [PinnedArray(np.ascontiguousarray(a, dtype=t)) for a, t in zip(step, dt)]
    return tuple(h.array for h in holders), holders


def device_count():
    return _lib.load().bppgpu_device_count()


class Engine:
    """One per GPU.  math="exact" reproduces the AVX association order bit for bit (4 states),
    math="fma" allows fused multiply-add."""

    def __init__(self, device=0, math="exact"):
        self.L = _lib.load()
        self.h = self.L.bppgpu_engine_create(device, MATH_FMA if math == "fma" else MATH_EXACT)
        _lib.check()
        if not self.h:
            raise BppGpuError("engine creation failed")
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.bppgpu_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_math(self, math):
        self.L.bppgpu_engine_set_math(self.h, MATH_FMA if math == "fma" else MATH_EXACT)

    def synchronize(self):
        self.L.bppgpu_engine_synchronize(self.h)
        _lib.check()

    @property
    def launch_count(self):
        return int(self.L.bppgpu_engine_launch_count(self.h))

    @property
    def bytes_allocated(self):
        return int(self.L.bppgpu_engine_bytes_allocated(self.h))

    def set_profiling(self, on):
        self.L.bppgpu_engine_set_profiling(self.h, int(bool(on)))

    def reset_profile(self):
        self.L.bppgpu_engine_reset_profile(self.h)

    def profile(self):
        ms = (C.c_double * 4)()
        cnt = (C.c_ulonglong * 4)()
        self.L.bppgpu_engine_get_profile(self.h, ms, cnt)
        return {k: {"ms": ms[i], "launches": int(cnt[i])} for i, k in enumerate(KERNELS)}


class Comm:
    """The path's one exchange step: an NCCL sum over ranks (threads.c:544-558, 583-590).

    One process per GPU: rank 0 calls Comm.unique_id(), the host ships the 128 bytes to every rank
    (bench.py: torch.distributed broadcast), every rank builds Comm(engine, nranks, rank, id).
    One process with several engines: Comm.init_all(engines)."""

    def __init__(self, engine, nranks, rank, uid, _handle=None):
        self.L, self.e = engine.L, engine
        if _handle is None:
            buf = (C.c_char * 128).from_buffer_copy(bytes(uid))
            _handle = self.L.bppgpu_comm_init_rank(engine.h, nranks, rank, C.cast(buf, C.c_void_p))
            _lib.check()
        if not _handle:
            raise BppGpuError("comm init failed")
        self.h, self.nranks, self.rank = _handle, nranks, rank

    @staticmethod
    def unique_id():
        L = _lib.load()
        buf = (C.c_char * 128)()
        L.bppgpu_comm_get_unique_id(C.cast(buf, C.c_void_p))
        _lib.check()
        return bytes(buf)

    @classmethod
    def init_all(cls, engines):
        L = _lib.load()
        n = len(engines)
        arr = (C.c_void_p * n)(*[e.h for e in engines])
        out = (C.c_void_p * n)()
        L.bppgpu_comm_init_all(arr, n, out)
        _lib.check()
        return [cls(e, n, i, None, _handle=out[i]) for i, e in enumerate(engines)]

    def allreduce_sum(self, values):
        a = _f64(values).copy()
        self.L.bppgpu_allreduce_sum(self.h, _dp(a), a.size)
        _lib.check()
        return a

    @staticmethod
    def allreduce_sum_all(comms, values):
        """One host thread driving all engines: values[i] belongs to comms[i]."""
        L = comms[0].L
        arrs = [_f64(v).copy() for v in values]
        n = len(comms)
        ch = (C.c_void_p * n)(*[c.h for c in comms])
        vp = (C.POINTER(C.c_double) * n)(*[_dp(a) for a in arrs])
        L.bppgpu_allreduce_sum_all(ch, n, vp, arrs[0].size)
        _lib.check()
        return arrs

    @property
    def calls(self):
        return int(self.L.bppgpu_comm_calls(self.h))

    def destroy(self):
        if getattr(self, "h", None):
            self.L.bppgpu_comm_destroy(self.h)
            self.h = None


class Locus:
    """Device mirror of locus_t.  `locus_create` arguments as in locus.c:622."""

    def __init__(self, engine, dtype, model, tips, clv_buffers, states, sites, rate_matrices,
                 prob_matrices, rate_cats, scale_buffers, attributes=ATTRIB_ARCH_CUDA):
        self.e, self.L = engine, engine.L
        self.tips, self.states, self.sites, self.rate_cats = tips, states, sites, rate_cats
        self.clv_buffers, self.prob_matrices, self.scale_buffers = clv_buffers, prob_matrices, scale_buffers
        self.h = self.L.bppgpu_locus_create(engine.h, dtype, model, tips, clv_buffers, states, sites,
                                            rate_matrices, prob_matrices, rate_cats, scale_buffers, attributes)
        _lib.check()
        if not self.h:
            raise BppGpuError("locus_create failed")

    @classmethod
    def create_like_bpp(cls, engine, tips, sites, states=4, rate_cats=1, scaling=False, model=None):
        """method.c:4137-4147: clv_buffers = 2*inner, prob_matrices = 2*edges, scale_buffers = 2*inner|0."""
        dtype = DATA_DNA if states == 4 else DATA_AA
        if model is None:
            model = DNA_MODEL_JC69 if states == 4 else 1
        inner, edges = tips - 1, 2 * tips - 2
        return cls(engine, dtype, model, tips, 2 * inner, states, sites, 1, 2 * edges, rate_cats,
                   2 * inner if scaling else 0)

    def destroy(self):
        if getattr(self, "h", None):
            self.L.bppgpu_locus_destroy(self.h)
            self.h = None

    # -- setters (pll_set_*)
    def set_tip_states(self, tip, charmap, sequence):
        seq = bytes(sequence)
        if len(seq) != self.sites:
            raise ValueError("sequence length != sites")
        cm = _u32(charmap)
        rc = self.L.bppgpu_set_tip_states(self.h, tip, _up(cm), seq)
        _lib.check()
        return rc

    def set_tip_clv(self, tip, clv):
        a = _f64(clv)
        if a.size != self.sites * self.states:
            raise ValueError("clv must hold sites*states doubles")
        rc = self.L.bppgpu_set_tip_clv(self.h, tip, _dp(a), 0)
        _lib.check()
        return rc

    def set_pattern_weights(self, w):
        a = _u32(w)
        assert a.size == self.sites
        self.L.bppgpu_set_pattern_weights(self.h, _up(a))
        _lib.check()

    def set_frequencies(self, freqs, index=0):
        a = _f64(freqs)
        assert a.size == self.states
        self.L.bppgpu_set_frequencies(self.h, index, _dp(a))
        _lib.check()

    def set_subst_params(self, params, index=0):
        a = _f64(params)
        assert a.size == self.states * (self.states - 1) // 2
        self.L.bppgpu_set_subst_params(self.h, index, _dp(a))
        _lib.check()

    def set_category_rates(self, rates):
        a = _f64(rates)
        assert a.size == self.rate_cats
        self.L.bppgpu_set_category_rates(self.h, _dp(a))

    def set_category_weights(self, w):
        a = _f64(w)
        assert a.size == self.rate_cats
        self.L.bppgpu_set_category_weights(self.h, _dp(a))

    def set_eigen(self, eigenvecs, inv_eigenvecs, eigenvals):
        a, b, c = _f64(eigenvecs), _f64(inv_eigenvecs), _f64(eigenvals)
        self.L.bppgpu_set_eigen(self.h, 0, _dp(a), _dp(b), _dp(c))

    def get_eigen(self):
        S = self.states
        a, b, c = np.zeros(S * S), np.zeros(S * S), np.zeros(S)
        self.L.bppgpu_get_eigen(self.h, 0, _dp(a), _dp(b), _dp(c))
        return a, b, c

    def set_diploid(self, resolution_count, mapping, unphased_weights):
        """diploid.c / method.c:4172-4193: phase-resolution counts and mapping of the unphased sites and
        THEIR weights (unphased_length entries, which may differ from `sites`)."""
        rc = np.ascontiguousarray(resolution_count, dtype=np.uint64)
        mp = np.ascontiguousarray(mapping, dtype=np.uint64)
        uw = _u32(unphased_weights)
        assert uw.size == rc.size
        ul = C.POINTER(C.c_ulong)
        r = self.L.bppgpu_set_diploid(self.h, len(rc), rc.ctypes.data_as(ul), mp.ctypes.data_as(ul), len(mp), _up(uw))
        _lib.check()
        return r

    # -- the hot triplet
    def update_matrices(self, pmatrix_indices, branch_lengths):
        idx, bl = _u32(pmatrix_indices), _f64(branch_lengths)
        assert idx.size == bl.size
        rc = self.L.bppgpu_update_matrices(self.h, idx.size, _up(idx), _dp(bl))
        _lib.check()
        return rc

    def update_partials(self, ops):
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        rc = self.L.bppgpu_update_partials(self.h, ops.size, _opp(ops))
        _lib.check()
        return rc

    def root_loglikelihood(self, root_clv_index, root_scaler_index=SCALE_BUFFER_NONE, persite=False):
        out = np.zeros(self.sites) if persite else None
        v = self.L.bppgpu_root_loglikelihood(self.h, root_clv_index, root_scaler_index,
                                             _dp(out) if persite else None)
        _lib.check()
        return (v, out) if persite else v

    def root_likelihood_vector(self, root_clv_index):
        out = np.zeros(self.sites)
        self.L.bppgpu_root_likelihood_vector(self.h, root_clv_index, _dp(out))
        _lib.check()
        return out

    def root_loglikelihood_diploid(self, root_clv_index):
        v = self.L.bppgpu_root_loglikelihood_diploid(self.h, root_clv_index)
        _lib.check()
        return v

    # -- raw buffers
    def get_clv(self, clv_index):
        out = np.zeros(self.sites * self.rate_cats * self.states)
        self.L.bppgpu_get_clv(self.h, clv_index, _dp(out))
        _lib.check()
        return out

    def get_pmatrix(self, idx):
        out = np.zeros(self.rate_cats * self.states * self.states)
        self.L.bppgpu_get_pmatrix(self.h, idx, _dp(out))
        _lib.check()
        return out

    def set_pmatrix(self, idx, values):
        a = _f64(values)
        assert a.size == self.rate_cats * self.states * self.states
        self.L.bppgpu_set_pmatrix(self.h, idx, _dp(a))
        _lib.check()

    def get_scaler(self, idx):
        out = np.zeros(self.sites, dtype=np.uint32)
        self.L.bppgpu_get_scaler(self.h, idx, _up(out))
        _lib.check()
        return out


class Batch:
    """An ordered set of loci launched together (per-locus arrays concatenated in batch order)."""

    def __init__(self, engine, loci):
        self.e, self.L = engine, engine.L
        self.loci = list(loci)
        arr = (C.c_void_p * len(self.loci))(*[l.h for l in self.loci])
        self.h = self.L.bppgpu_batch_create(engine.h, len(self.loci), arr)
        _lib.check()
        if not self.h:
            raise BppGpuError("batch_create failed")
        self.n = len(self.loci)
        self._keep = None

    def destroy(self):
        if getattr(self, "h", None):
            self.L.bppgpu_batch_destroy(self.h)
            self.h = None

    def update_matrices(self, counts, pmatrix_indices, branch_lengths):
        c, i, b = _u32(counts), _u32(pmatrix_indices), _f64(branch_lengths)
        rc = self.L.bppgpu_batch_update_matrices(self.h, _up(c), _up(i), _dp(b))
        _lib.check()
        return rc

    def update_partials(self, counts, ops):
        c, o = _u32(counts), np.ascontiguousarray(ops, dtype=OP_DTYPE)
        rc = self.L.bppgpu_batch_update_partials(self.h, _up(c), _opp(o))
        _lib.check()
        return rc

    def root_loglikelihood(self, root_clv, root_scaler=None):
        r = _u32(root_clv)
        s = _i32(root_scaler if root_scaler is not None else np.full(self.n, -1))
        out = np.zeros(self.n)
        self.L.bppgpu_batch_root_loglikelihood(self.h, _up(r), _ip(s), _dp(out))
        _lib.check()
        return out

    @staticmethod
    def _step_args(step):
        mc, mi, mb, oc, ops, rc, rs = step
        return (_u32(mc), _u32(mi), _f64(mb), _u32(oc), np.ascontiguousarray(ops, dtype=OP_DTYPE),
                _u32(rc), _i32(rs))

    def prepare(self, step):
        """Resolve the ctypes pointers of a step's arrays once; the returned PreparedStep can be staged
        repeatedly (its arrays may be modified in place between steps, like a C caller's buffers)."""
        return PreparedStep(self._step_args(step))

    def set_waves(self, waves):
        self.L.bppgpu_batch_set_waves(self.h, int(waves))

    def full_pass(self, step):
        """step = (matrix_counts, pmatrix_indices, branch_lengths, op_counts, ops, root_clv, root_scaler)
        host arrays; returns (lnl[n], lnl_sum).  H2D + kernels + D2H in one call."""
        mc, mi, mb, oc, ops, rc, rs = self._step_args(step)
        out = np.zeros(self.n)
        tot = C.c_double(0)
        self.L.bppgpu_batch_full_pass(self.h, _up(mc), _up(mi), _dp(mb), _up(oc), _opp(ops), _up(rc), _ip(rs),
                                      _dp(out), C.byref(tot))
        _lib.check()
        return out, tot.value

    def stage(self, step):
        if isinstance(step, PreparedStep):
            self.L.bppgpu_batch_stage(self.h, *step.ptrs)
        else:
            mc, mi, mb, oc, ops, rc, rs = self._step_args(step)
            self.L.bppgpu_batch_stage(self.h, _up(mc), _up(mi), _dp(mb), _up(oc), _opp(ops), _up(rc), _ip(rs))
        _lib.check()

    def run(self):
        self.L.bppgpu_batch_run(self.h)
        _lib.check()

    def collect(self, out=None):
        """D2H of the n per-locus lnL values and their sum; `out` (float64[n]) is reused when given."""
        if out is None:
            out = np.zeros(self.n)
            ptr = _dp(out)
        else:
            ptr = getattr(self, "_out_ptr", None)
            if ptr is None or self._out_ref is not out:
                ptr = _dp(out)
                self._out_ptr, self._out_ref = ptr, out
        tot = C.c_double(0)
        self.L.bppgpu_batch_collect(self.h, ptr, C.byref(tot))
        _lib.check()
        return out, tot.value

    def synchronize(self):
        self.L.bppgpu_batch_synchronize(self.h)
        _lib.check()

    def wait_inputs(self):
        self.L.bppgpu_batch_wait_inputs(self.h)
        _lib.check()

    def flip_indices(self):
        """SWAP_CLV_INDEX / SWAP_SCALER_INDEX / SWAP_PMAT_INDEX of every inner node and edge of the staged step,
        on the device (a whole-tree proposal, prop_mixing.c:100-124)."""
        self.L.bppgpu_batch_flip_indices(self.h)
        _lib.check()

    def set_branch_lengths(self, bl):
        """New branch lengths for the staged matrix list (float64, same order; keep `bl` alive and unchanged
        until collect / wait_inputs when it is pinned)."""
        a = bl if (isinstance(bl, np.ndarray) and bl.dtype == np.float64 and bl.flags.c_contiguous) else _f64(bl)
        self._bl_keep = a
        self.L.bppgpu_batch_set_branch_lengths(self.h, _dp(a))
        _lib.check()

    def allreduce_lnl_sum(self, comm):
        """NCCL sum of the batch's lnL sum over ranks, on the batch stream behind run()."""
        self.L.bppgpu_batch_allreduce_lnl_sum(self.h, comm.h)
        _lib.check()

    def timer_start(self):
        self.L.bppgpu_batch_timer_start(self.h)

    def timer_stop_ms(self):
        return self.L.bppgpu_batch_timer_stop_ms(self.h)

    @property
    def kernel_name(self):
        return self.L.bppgpu_batch_kernel_name(self.h).decode()

    def plan_stats(self):
        """Loci per path of the last planned step (4-state batches): see bppgpu_batch_plan_stats."""
        out = (C.c_uint * 8)()
        if not self.L.bppgpu_batch_plan_stats(self.h, out):
            return None
        keys = ("fast", "lean", "scaled", "walker", "max_chunks", "slots", "cells_per_thread", "smem_bytes")
        return dict(zip(keys, [int(v) for v in out]))

    @property
    def lnl_sum_dev(self):
        return self.L.bppgpu_batch_lnl_sum_dev(self.h)

    @property
    def stream(self):
        return self.L.bppgpu_batch_stream(self.h)


class PreparedStep:
    """The 7 arrays of a full-pass step with their ctypes pointers resolved (see Batch.prepare)."""

    def __init__(self, arrays):
        self.arrays = arrays
        mc, mi, mb, oc, ops, rc, rs = arrays
        self.ptrs = (_up(mc), _up(mi), _dp(mb), _up(oc), _opp(ops), _up(rc), _ip(rs))


class GeneTrees:
    """gnode_t bookkeeping for N loci with T tips each: topology, ages and the three flipping
    buffer indices of every node (gtree.c:2395-2399,2664-2675; locus.c:24-26)."""

    def __init__(self, left, right, times, rate_mui, scaling):
        self.left = np.ascontiguousarray(left, dtype=np.int64)        # [N, T-1]
        self.right = np.ascontiguousarray(right, dtype=np.int64)
        self.times = np.array(times, dtype=np.float64)                 # [N, 2T-1]
        self.rate_mui = np.array(rate_mui, dtype=np.float64)           # [N]
        self.N, self.T = self.left.shape[0], self.left.shape[1] + 1
        N, T = self.N, self.T
        nn = 2 * T - 1
        self.scaling = bool(scaling)
        self.clv_index = np.tile(np.arange(nn, dtype=np.int64), (N, 1))
        self.pmatrix_index = np.tile(np.arange(nn, dtype=np.int64), (N, 1))
        sc = np.full(nn, SCALE_BUFFER_NONE, dtype=np.int64)
        if scaling:
            sc[T:] = np.arange(T - 1)
        self.scaler_index = np.tile(sc, (N, 1))
        rows = np.arange(N)[:, None]
        self.parent = np.full((N, nn), -1, dtype=np.int64)
        inner = np.arange(T, nn)[None, :].repeat(N, 0)
        self.parent[rows, self.left] = inner
        self.parent[rows, self.right] = inner
        self.root = nn - 1
        self.order = self._post_orders()                               # [N, T-1] node ids

    def _post_orders(self):
        """recursive left,right,node order of every locus (prop_mixing.c:28-50)."""
        N, T = self.N, self.T
        out = np.zeros((N, T - 1), dtype=np.int64)
        for i in range(N):
            L, R = self.left[i], self.right[i]
            k = 0
            stack = [(2 * T - 2, 0)]
            while stack:
                node, st = stack.pop()
                if node < T:
                    continue
                if st == 0:
                    stack.append((node, 1))
                    stack.append((int(R[node - T]), 0))
                    stack.append((int(L[node - T]), 0))
                else:
                    out[i, k] = node
                    k += 1
        return out

    # SWAP_CLV_INDEX / SWAP_SCALER_INDEX / SWAP_PMAT_INDEX, locus.c:24-26
    def flip_clv(self, nodes=None):
        T = self.T
        sel = slice(T, None) if nodes is None else nodes
        self.clv_index[:, sel] = T + (self.clv_index[:, sel] - 1) % (2 * T - 2)
        if self.scaling:
            self.scaler_index[:, sel] = (T + self.scaler_index[:, sel] - 1) % (2 * T - 2)

    def flip_pmatrix(self):
        e = 2 * self.T - 2
        self.pmatrix_index[:, :-1] = (e + self.pmatrix_index[:, :-1]) % (2 * e)

    def branch_lengths(self):
        """(parent.time - node.time) * rate_mui for the 2T-2 non-root nodes (core_pmatrix.c:711-715)."""
        rows = np.arange(self.N)[:, None]
        par = self.parent[:, :-1]
        return (self.times[rows, par] - self.times[:, :-1]) * self.rate_mui[:, None]

    def full_pass_step(self):
        """Host arrays of one full-tree pass over all loci: every branch's P-matrix, every inner
        CLV in post-order, the root lnL (what prop_mixing_update_gtrees does per locus)."""
        N, T = self.N, self.T
        rows = np.arange(N)[:, None]
        mc = np.full(N, 2 * T - 2, dtype=np.uint32)
        mi = self.pmatrix_index[:, :-1].astype(np.uint32).ravel()
        mb = self.branch_lengths().ravel()
        order = self.order
        l, r = self.left[rows, order - T], self.right[rows, order - T]
        ops = np.zeros((N, T - 1), dtype=OP_DTYPE)
        ops["parent_clv_index"] = self.clv_index[rows, order]
        ops["left_clv_index"] = self.clv_index[rows, l]
        ops["right_clv_index"] = self.clv_index[rows, r]
        ops["left_pmatrix_index"] = self.pmatrix_index[rows, l]
        ops["right_pmatrix_index"] = self.pmatrix_index[rows, r]
        ops["parent_scaler_index"] = self.scaler_index[rows, order]
        ops["left_scaler_index"] = self.scaler_index[rows, l]
        ops["right_scaler_index"] = self.scaler_index[rows, r]
        oc = np.full(N, T - 1, dtype=np.uint32)
        rc = self.clv_index[:, self.root].astype(np.uint32)
        rs = self.scaler_index[:, self.root].astype(np.int32)
        return mc, mi, mb, oc, ops.ravel(), rc, rs


def age_move_step(trees, nodes, new_ages):
    """Host arrays of one gene-tree age move per locus (gtree.c:5437-5467): node `nodes[i]` of locus i gets the age
    `new_ages[i]`; the P-matrix indices of the 2-3 edges around it and the CLV / scaler indices of its root path
    are flipped (in `trees`, like the reference does before it calls the seam), and the step recomputes exactly
    those: 2-3 matrices, the root path's partials, the root lnL.  Ragged per locus."""
    N, T = trees.N, trees.T
    rows = np.arange(N)
    nodes = np.asarray(nodes, dtype=np.int64)
    trees.times[rows, nodes] = new_ages
    e = 2 * T - 2
    l, r = trees.left[rows, nodes - T], trees.right[rows, nodes - T]
    not_root = trees.parent[rows, nodes] >= 0
    for sel, who in ((np.ones(N, bool), l), (np.ones(N, bool), r), (not_root, nodes)):
        trees.pmatrix_index[rows[sel], who[sel]] = (e + trees.pmatrix_index[rows[sel], who[sel]]) % (2 * e)
    # root paths
    paths, cur, alive = [], nodes.copy(), np.ones(N, bool)
    while alive.any():
        paths.append(np.where(alive, cur, -1))
        ci = rows[alive], cur[alive]
        trees.clv_index[ci] = T + (trees.clv_index[ci] - 1) % (2 * T - 2)
        if trees.scaling:
            trees.scaler_index[ci] = (T + trees.scaler_index[ci] - 1) % (2 * T - 2)
        nxt = np.where(alive, trees.parent[rows, np.maximum(cur, 0)], -1)
        alive = nxt >= 0
        cur = nxt
    path = np.stack(paths, axis=1)                                  # [N, depth], -1 padded
    oc = (path >= 0).sum(axis=1).astype(np.uint32)
    bl = trees.branch_lengths()
    mc = (2 + not_root).astype(np.uint32)
    mnodes = np.stack([l, r, np.where(not_root, nodes, -1)], axis=1)
    msel = mnodes >= 0
    mrows = np.repeat(rows, 3).reshape(N, 3)[msel]
    mi = trees.pmatrix_index[mrows, mnodes[msel]].astype(np.uint32)
    mb = bl[mrows, mnodes[msel]]
    osel = path >= 0
    orow = np.repeat(rows, path.shape[1]).reshape(path.shape)[osel]
    on = path[osel]
    ol, orr = trees.left[orow, on - T], trees.right[orow, on - T]
    ops = np.zeros(on.size, dtype=OP_DTYPE)
    ops["parent_clv_index"] = trees.clv_index[orow, on]
    ops["left_clv_index"] = trees.clv_index[orow, ol]
    ops["right_clv_index"] = trees.clv_index[orow, orr]
    ops["left_pmatrix_index"] = trees.pmatrix_index[orow, ol]
    ops["right_pmatrix_index"] = trees.pmatrix_index[orow, orr]
    ops["parent_scaler_index"] = trees.scaler_index[orow, on]
    ops["left_scaler_index"] = trees.scaler_index[orow, ol]
    ops["right_scaler_index"] = trees.scaler_index[orow, orr]
    rc = trees.clv_index[:, trees.root].astype(np.uint32)
    rs = trees.scaler_index[:, trees.root].astype(np.int32)
    return mc, mi, mb, oc, ops, rc, rs


def propose_ages(trees, rng):
    """A random inner node per locus and a new age strictly between its older child and its parent (the root: up to
    10 % older), the way the age move samples inside its bounds (gtree.c:4585-5436)."""
    N, T = trees.N, trees.T
    rows = np.arange(N)
    nodes = rng.integers(T, 2 * T - 1, size=N)
    l, r = trees.left[rows, nodes - T], trees.right[rows, nodes - T]
    lo = np.maximum(trees.times[rows, l], trees.times[rows, r])
    par = trees.parent[rows, nodes]
    hi = np.where(par >= 0, trees.times[rows, np.maximum(par, 0)], trees.times[rows, nodes] * 1.1)
    u = rng.uniform(0.05, 0.95, size=N)
    return nodes, lo + u * (hi - lo)


def load_workload(engine, w, charmap=None):
    """Create the loci of a synth.Workload on `engine`; returns (loci, GeneTrees)."""
    from . import synth
    if charmap is None:
        charmap = synth.iupac_nt_map() if w.states == 4 else synth.aa_map()
    model = 1 if w.model == "LG" else DNA_MODELS[w.model]
    loci = []
    for i in range(w.n_loci):
        l = Locus.create_like_bpp(engine, w.tips, w.sites, w.states, w.rate_cats, w.scaling, model)
        for t in range(w.tips):
            l.set_tip_states(t, charmap, w.tip_chars[i, t].tobytes())
        l.set_pattern_weights(w.weights[i])
        l.set_frequencies(w.freqs[i])
        if w.model != "JC69":
            l.set_subst_params(w.subst[i])
        l.set_category_rates(w.rates)
        loci.append(l)
    trees = GeneTrees(w.left, w.right, w.times, w.rate_mui, w.scaling)
    return loci, trees
