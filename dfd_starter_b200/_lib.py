"""ctypes binding of libdfd_b200.so (the C ABI declared in include/dfd_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device
is present when a context is created, the product path raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdfd_b200.so")
_lib = None

# every symbol include/dfd_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "dfd_abi_version", "dfd_last_error", "dfd_ctx_create", "dfd_ctx_destroy", "dfd_ctx_sm_count",
    "dfd_ctx_launch_count", "dfd_table_replica_stride", "dfd_table_scratch_bytes", "dfd_table_build",
    "dfd_perturb_members", "dfd_policy_num_params", "dfd_policy_num_buffers", "dfd_policy_out_width",
    "dfd_policy_forward", "dfd_impala_scratch_bytes", "dfd_impala_forward", "dfd_fd_prepare_scratch_bytes",
    "dfd_fd_prepare", "dfd_fd_reduce_scratch_bytes", "dfd_fd_reduce", "dfd_dsgd_step", "dfd_dsgd_scratch_bytes", "dfd_synthetic_reward",
    "dfd_fd_prepare_partial", "dfd_xchg_mailbox_bytes", "dfd_xchg_mailbox_create", "dfd_xchg_mailbox_open",
    "dfd_xchg_mailbox_close", "dfd_xchg_mailbox_destroy", "dfd_xchg_allreduce",
    "dfd_fd_step_fused_scratch_bytes", "dfd_fd_step_fused", "dfd_wire_count_returns", "dfd_wire_decode_returns",
    "dfd_strategy_distances", "dfd_host_stage", "dfd_normalize_obs", "dfd_member_obs_stats",
    "dfd_xchg_gather_f64", "dfd_table_scaled16_bytes", "dfd_table_build_scaled16", "dfd_table_drop_scaled16", "dfd_policy_direct_supported",
    "dfd_rng_scratch_bytes", "dfd_rng_normal_rows",
]


class DfdTable(C.Structure):
    _fields_ = [("replicas", C.c_void_p), ("replica_stride", C.c_int64), ("prefix_sq", C.c_void_p),
                ("size", C.c_int64)]


class DfdPolicyDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_in", C.c_int), ("h1", C.c_int), ("h2", C.c_int), ("n_act", C.c_int),
                ("precision", C.c_int)]


class DfdFdRows(C.Structure):
    _fields_ = [("row_ptr", C.c_void_p), ("row_coef", C.c_void_p), ("max_rows", C.c_int)]


class DfdReturnSoa(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("epoch", "idx", "sign", "reward", "novelty", "entropy", "timesteps", "is_eval",
                                          "key_off", "key_len", "states_off", "states_len", "shape_off", "shape_len",
                                          "stats_off", "stats_len")]


class DfdError(RuntimeError):
    pass


def load():
    """Load the library and declare every prototype.  Needs no GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DfdError("%s not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                       "dfd_starter_b200 has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, f32, f64, sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double, C.c_size_t
    P = C.POINTER

    def proto(name, res, args):
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args

    proto("dfd_abi_version", i32, [])
    proto("dfd_last_error", C.c_char_p, [])
    proto("dfd_ctx_create", i32, [i32, P(vp)])
    proto("dfd_ctx_destroy", i32, [vp])
    proto("dfd_ctx_sm_count", i32, [vp])
    proto("dfd_ctx_launch_count", i64, [vp])
    proto("dfd_table_replica_stride", i64, [i64])
    proto("dfd_table_scratch_bytes", sz, [i64])
    proto("dfd_table_build", i32, [vp, vp, i64, vp, i64, vp, vp, sz, vp])
    proto("dfd_table_scaled16_bytes", sz, [i64, i64])
    proto("dfd_table_build_scaled16", i32, [vp, P(DfdTable), f32, i64, vp, sz, vp])
    proto("dfd_table_drop_scaled16", i32, [vp])
    proto("dfd_policy_direct_supported", i32, [P(DfdPolicyDesc)])
    proto("dfd_perturb_members", i32, [vp, P(DfdTable), vp, i64, vp, vp, i32, f32, vp, i64, vp])
    proto("dfd_policy_num_params", i64, [P(DfdPolicyDesc)])
    proto("dfd_policy_num_buffers", i64, [P(DfdPolicyDesc)])
    proto("dfd_policy_out_width", i64, [P(DfdPolicyDesc)])
    proto("dfd_policy_forward", i32, [vp, P(DfdPolicyDesc), P(DfdTable), vp, vp, vp, vp, i32, f32, vp, i32, vp, vp])
    proto("dfd_impala_scratch_bytes", sz, [i32, i32])
    proto("dfd_impala_forward", i32, [vp, P(DfdPolicyDesc), P(DfdTable), vp, vp, vp, vp, i32, f32, vp, vp, vp, vp,
                                      vp, i32, vp, vp, vp, vp, sz, vp])
    proto("dfd_fd_prepare_scratch_bytes", sz, [i32, i32])
    proto("dfd_fd_prepare", i32, [vp, P(DfdTable), i64, vp, vp, vp, vp, i32, i32, f64, f32, vp, i64, i32, vp, i32,
                                  P(DfdFdRows), vp, sz, vp])
    proto("dfd_fd_reduce_scratch_bytes", sz, [vp, i64, i32])
    proto("dfd_fd_reduce", i32, [vp, P(DfdFdRows), i32, i64, vp, vp, sz, vp])
    proto("dfd_dsgd_scratch_bytes", sz, [i64])
    proto("dfd_dsgd_step", i32, [vp, vp, vp, i64, f64, f64, vp, vp, i64, i32, i32, vp, vp, sz, vp])
    proto("dfd_synthetic_reward", i32, [vp, vp, i32, i32, i32, vp, vp, vp])
    proto("dfd_fd_prepare_partial", i32, [vp, P(DfdTable), i64, vp, vp, vp, vp, i32, f64, f32, P(DfdFdRows), vp, vp, sz, vp])
    proto("dfd_xchg_mailbox_bytes", sz, [i64, i32])
    proto("dfd_xchg_mailbox_create", i32, [vp, sz, P(vp), C.c_char_p])
    proto("dfd_xchg_mailbox_open", i32, [vp, C.c_char_p, P(vp)])
    proto("dfd_xchg_mailbox_close", i32, [vp, vp])
    proto("dfd_xchg_mailbox_destroy", i32, [vp, vp])
    proto("dfd_xchg_allreduce", i32, [vp, vp, i32, i32, i64, vp, vp, vp, vp])
    proto("dfd_xchg_gather_f64", i32, [vp, vp, i32, i32, i64, vp, i32, vp, vp])
    proto("dfd_fd_step_fused_scratch_bytes", sz, [vp, i64, i32, i32])
    proto("dfd_fd_step_fused", i32, [vp, P(DfdTable), i64, vp, vp, vp, i32, i32, f64, f32, vp, vp, f64, f64, vp, vp, i64,
                                     i32, i32, vp, vp, i32, i32, vp, sz, vp])
    proto("dfd_strategy_distances", i32, [vp, vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp])
    proto("dfd_normalize_obs", i32, [vp, vp, i64, i32, vp, vp, i32, f32, vp, vp])
    proto("dfd_member_obs_stats", i32, [vp, vp, vp, i32, i32, i32, vp, vp])
    proto("dfd_rng_scratch_bytes", sz, [i32, i64, i64, f64])
    proto("dfd_rng_normal_rows", i32, [vp, vp, i32, i64, i64, vp, f64, vp, vp, vp, i64, vp, vp, vp, i32, i32, f64, vp, sz, vp])
    proto("dfd_host_stage", i32, [vp, vp, vp, sz, vp])
    proto("dfd_wire_count_returns", i64, [C.c_char_p, sz])
    proto("dfd_wire_decode_returns", i64, [C.c_char_p, sz, i32, i64, P(DfdReturnSoa)])
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = load().dfd_last_error().decode("utf-8", "replace")
        raise DfdError("%s failed (status %d): %s" % (what or "dfd call", rc, msg))
