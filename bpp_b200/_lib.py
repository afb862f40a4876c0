"""ctypes loader of the C-ABI library (include/bpp_b200.h).

There is no CPU fallback: if bpp_b200/libbppgpu.so is missing or no B200 is visible, every
entry point raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbppgpu.so")

# every symbol include/bpp_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "bppgpu_device_count", "bppgpu_engine_create", "bppgpu_engine_destroy", "bppgpu_engine_device",
    "bppgpu_engine_set_math", "bppgpu_engine_synchronize", "bppgpu_engine_stream",
    "bppgpu_engine_launch_count", "bppgpu_engine_bytes_allocated", "bppgpu_last_error",
    "bppgpu_set_fatal_handler", "bppgpu_version", "bppgpu_host_alloc", "bppgpu_host_free", "bppgpu_engine_set_profiling",
    "bppgpu_engine_get_profile", "bppgpu_engine_reset_profile",
    "bppgpu_locus_create", "bppgpu_locus_destroy", "bppgpu_set_tip_states", "bppgpu_set_tip_clv",
    "bppgpu_set_pattern_weights", "bppgpu_set_frequencies", "bppgpu_set_subst_params",
    "bppgpu_set_category_rates", "bppgpu_set_category_weights", "bppgpu_set_eigen", "bppgpu_get_eigen",
    "bppgpu_update_matrices", "bppgpu_update_partials", "bppgpu_root_loglikelihood",
    "bppgpu_root_likelihood_vector", "bppgpu_set_diploid", "bppgpu_root_loglikelihood_diploid",
    "bppgpu_get_clv", "bppgpu_get_pmatrix", "bppgpu_set_pmatrix", "bppgpu_get_scaler",
    "bppgpu_batch_create", "bppgpu_batch_destroy", "bppgpu_batch_size", "bppgpu_batch_kernel_name", "bppgpu_batch_plan_stats",
    "bppgpu_batch_update_matrices", "bppgpu_batch_update_partials", "bppgpu_batch_root_loglikelihood",
    "bppgpu_batch_full_pass", "bppgpu_batch_stage", "bppgpu_batch_run", "bppgpu_batch_set_waves", "bppgpu_batch_collect",
    "bppgpu_batch_wait_inputs", "bppgpu_batch_flip_indices", "bppgpu_batch_set_branch_lengths",
    "bppgpu_batch_lnl_sum_dev", "bppgpu_batch_stream", "bppgpu_batch_timer_start",
    "bppgpu_batch_timer_stop_ms", "bppgpu_batch_synchronize",
    "bppgpu_comm_nccl_version", "bppgpu_comm_get_unique_id", "bppgpu_comm_init_rank", "bppgpu_comm_init_all",
    "bppgpu_comm_destroy", "bppgpu_comm_nranks", "bppgpu_comm_rank", "bppgpu_comm_calls",
    "bppgpu_allreduce_sum", "bppgpu_allreduce_sum_all", "bppgpu_batch_allreduce_lnl_sum",
]


class PartialOp(C.Structure):
    """struct bppgpu_partial_op"""
    _fields_ = [("parent_clv_index", C.c_uint), ("left_clv_index", C.c_uint), ("right_clv_index", C.c_uint),
                ("left_pmatrix_index", C.c_uint), ("right_pmatrix_index", C.c_uint),
                ("parent_scaler_index", C.c_int), ("left_scaler_index", C.c_int), ("right_scaler_index", C.c_int)]


class BppGpuError(RuntimeError):
    pass


_lib = None
_handler_ref = None


def load():
    """Load libbppgpu.so and declare the prototypes.  Raises if the library has not been built."""
    global _lib, _handler_ref
    if _lib is not None:
        return _lib
    path = os.environ.get("BPPGPU_LIB", LIB_PATH)        # tuning builds (tools/): another build of the same library
    if not os.path.exists(path):
        raise BppGpuError("%s not found: build it with `python -m bpp_b200.build` "
                          "(there is no CPU fallback)" % path)
    L = C.CDLL(path)
    vp, u, i, d = C.c_void_p, C.c_uint, C.c_int, C.c_double
    up, ip, dp = C.POINTER(C.c_uint), C.POINTER(C.c_int), C.POINTER(C.c_double)
    opp = C.POINTER(PartialOp)
    ull = C.c_ulonglong
    sig = {
        "bppgpu_device_count": (i, []),
        "bppgpu_engine_create": (vp, [i, u]),
        "bppgpu_engine_destroy": (None, [vp]),
        "bppgpu_engine_device": (i, [vp]),
        "bppgpu_engine_set_math": (None, [vp, u]),
        "bppgpu_engine_synchronize": (None, [vp]),
        "bppgpu_engine_stream": (vp, [vp]),
        "bppgpu_engine_launch_count": (ull, [vp]),
        "bppgpu_engine_bytes_allocated": (ull, [vp]),
        "bppgpu_last_error": (C.c_char_p, []),
        "bppgpu_set_fatal_handler": (None, [vp]),
        "bppgpu_version": (C.c_char_p, []),
        "bppgpu_host_alloc": (vp, [C.c_size_t]),
        "bppgpu_host_free": (None, [vp]),
        "bppgpu_engine_set_profiling": (None, [vp, i]),
        "bppgpu_engine_get_profile": (None, [vp, dp, C.POINTER(ull)]),
        "bppgpu_engine_reset_profile": (None, [vp]),
        "bppgpu_locus_create": (vp, [vp] + [u] * 11),
        "bppgpu_locus_destroy": (None, [vp]),
        "bppgpu_set_tip_states": (i, [vp, u, up, C.c_char_p]),
        "bppgpu_set_tip_clv": (i, [vp, u, dp, i]),
        "bppgpu_set_pattern_weights": (None, [vp, up]),
        "bppgpu_set_frequencies": (None, [vp, u, dp]),
        "bppgpu_set_subst_params": (None, [vp, u, dp]),
        "bppgpu_set_category_rates": (None, [vp, dp]),
        "bppgpu_set_category_weights": (None, [vp, dp]),
        "bppgpu_set_eigen": (None, [vp, u, dp, dp, dp]),
        "bppgpu_get_eigen": (None, [vp, u, dp, dp, dp]),
        "bppgpu_update_matrices": (i, [vp, u, up, dp]),
        "bppgpu_update_partials": (i, [vp, u, opp]),
        "bppgpu_root_loglikelihood": (d, [vp, u, i, dp]),
        "bppgpu_root_likelihood_vector": (i, [vp, u, dp]),
        "bppgpu_set_diploid": (i, [vp, u, C.POINTER(C.c_ulong), C.POINTER(C.c_ulong), C.c_ulong, up]),
        "bppgpu_root_loglikelihood_diploid": (d, [vp, u]),
        "bppgpu_get_clv": (i, [vp, u, dp]),
        "bppgpu_get_pmatrix": (i, [vp, u, dp]),
        "bppgpu_set_pmatrix": (i, [vp, u, dp]),
        "bppgpu_get_scaler": (i, [vp, u, up]),
        "bppgpu_batch_create": (vp, [vp, u, C.POINTER(vp)]),
        "bppgpu_batch_destroy": (None, [vp]),
        "bppgpu_batch_size": (u, [vp]),
        "bppgpu_batch_kernel_name": (C.c_char_p, [vp]),
        "bppgpu_batch_plan_stats": (C.c_int, [vp, C.POINTER(C.c_uint)]),
        "bppgpu_batch_update_matrices": (i, [vp, up, up, dp]),
        "bppgpu_batch_update_partials": (i, [vp, up, opp]),
        "bppgpu_batch_root_loglikelihood": (i, [vp, up, ip, dp]),
        "bppgpu_batch_full_pass": (i, [vp, up, up, dp, up, opp, up, ip, dp, dp]),
        "bppgpu_batch_stage": (i, [vp, up, up, dp, up, opp, up, ip]),
        "bppgpu_batch_run": (i, [vp]),
        "bppgpu_batch_set_waves": (None, [vp, u]),
        "bppgpu_batch_collect": (i, [vp, dp, dp]),
        "bppgpu_batch_wait_inputs": (i, [vp]),
        "bppgpu_batch_flip_indices": (i, [vp]),
        "bppgpu_batch_set_branch_lengths": (i, [vp, dp]),
        "bppgpu_batch_lnl_sum_dev": (vp, [vp]),
        "bppgpu_batch_stream": (vp, [vp]),
        "bppgpu_batch_timer_start": (None, [vp]),
        "bppgpu_batch_timer_stop_ms": (d, [vp]),
        "bppgpu_batch_synchronize": (None, [vp]),
        "bppgpu_comm_nccl_version": (i, []),
        "bppgpu_comm_get_unique_id": (i, [vp]),
        "bppgpu_comm_init_rank": (vp, [vp, i, i, vp]),
        "bppgpu_comm_init_all": (i, [C.POINTER(vp), i, C.POINTER(vp)]),
        "bppgpu_comm_destroy": (None, [vp]),
        "bppgpu_comm_nranks": (i, [vp]),
        "bppgpu_comm_rank": (i, [vp]),
        "bppgpu_comm_calls": (ull, [vp]),
        "bppgpu_allreduce_sum": (i, [vp, dp, i]),
        "bppgpu_allreduce_sum_all": (i, [C.POINTER(vp), i, C.POINTER(dp), i]),
        "bppgpu_batch_allreduce_lnl_sum": (i, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args

    # route the library's fatal() (default: exit(1) like util.c:30) into a Python exception
    @C.CFUNCTYPE(None, C.c_char_p)
    def _on_fatal(msg):
        _on_fatal.last = msg.decode(errors="replace")
    _on_fatal.last = None
    _handler_ref = _on_fatal
    L.bppgpu_set_fatal_handler(C.cast(_on_fatal, C.c_void_p))
    _lib = L
    return L


def check():
    """Raise the pending fatal error of the library, if any."""
    if _handler_ref is not None and _handler_ref.last:
        msg, _handler_ref.last = _handler_ref.last, None
        raise BppGpuError(msg)
