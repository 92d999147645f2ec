"""ctypes binding of libphoskin_b200.so (include/phoskin_b200.h).

The product path has no CPU fallback: if the shared library is missing or a call is made
without a CUDA device, it raises.  The library itself links cudart statically, so it *loads*
on a CPU-only host (used by the `-m "not gpu"` symbol tests) but every compute call fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHOSKIN_LIB") or os.path.join(_HERE, "libphoskin_b200.so")   # PHOSKIN_LIB: debug builds

PK_HOST, PK_DEVICE = 0, 1
MODEL_IDS = {"distmod": 0, "succmod": 1, "randmod": 2}
Y_METRIC_IDS = {"total_signal": 0, "mean_activity": 1, "variance": 2, "dynamics": 3, "l2_norm": 4}
METHOD_IDS = {None: 0, "default": 0, "rodas4": 1, "ros5l": 2, "ros6l": 3}
STATUS_NAMES = {0: "ok", 1: "max_steps", 2: "step_underflow", 3: "non_finite"}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class PkLocalJob(C.Structure):
    """Mirror of `struct pk_local_job` (include/phoskin_b200.h) — pointers as void* so that
    host (numpy) and device (torch) addresses go through the same fields."""
    _fields_ = [
        ("model", C.c_int32), ("n_sites", C.c_int32), ("B", C.c_int64), ("T", C.c_int32),
        ("memspace", C.c_int32),
        ("params", C.c_void_p), ("y0", C.c_void_p), ("y0_stride", C.c_int64), ("t", C.c_void_p),
        ("rtol", C.c_double), ("atol", C.c_double),
        ("max_steps", C.c_int32), ("normalize", C.c_int32), ("log_params", C.c_int32),
        ("y_metric", C.c_int32), ("method", C.c_int32), ("reserved0", C.c_int32),
        ("out_sol", C.c_void_p), ("out_flat", C.c_void_p), ("out_Y", C.c_void_p),
        ("out_ssr", C.c_void_p), ("out_score", C.c_void_p), ("out_status", C.c_void_p),
        ("out_nsteps", C.c_void_p), ("out_nrej", C.c_void_p),
        ("target", C.c_void_p), ("sigma", C.c_void_p), ("group", C.c_void_p),
        ("n_groups", C.c_int32), ("sigma_len", C.c_int32), ("lam", C.c_double),
        ("score_w", C.c_double * 5), ("lam_group", C.c_void_p),
    ]


class PkNllsJob(C.Structure):
    """Mirror of `struct pk_nlls_job` — one batched call replacing a loop of curve_fit
    (reference paramest/normest.py:79-89, 278-290, 494-509)."""
    _fields_ = [
        ("model", C.c_int32), ("n_sites", C.c_int32), ("B", C.c_int64), ("T", C.c_int32), ("memspace", C.c_int32),
        ("theta", C.c_void_p), ("y0", C.c_void_p), ("y0_stride", C.c_int64), ("t", C.c_void_p),
        ("lb", C.c_void_p), ("ub", C.c_void_p), ("target", C.c_void_p), ("sigma", C.c_void_p), ("group", C.c_void_p),
        ("n_groups", C.c_int32), ("sigma_len", C.c_int32), ("lam", C.c_double),
        ("log_params", C.c_int32), ("max_iter", C.c_int32),
        ("ftol", C.c_double), ("xtol", C.c_double), ("gtol", C.c_double), ("fd_rel", C.c_double), ("mu0", C.c_double),
        ("rtol", C.c_double), ("atol", C.c_double), ("max_steps", C.c_int32), ("method", C.c_int32),
        ("score_w", C.c_double * 5),
        ("out_cost", C.c_void_p), ("out_score", C.c_void_p), ("out_status", C.c_void_p), ("out_iters", C.c_void_p),
        ("out_nfev", C.c_void_p), ("lam_group", C.c_void_p),
    ]


NLLS_STATUS_NAMES = {1: "gtol", 2: "ftol", 3: "xtol", 4: "max_iter", -1: "failed"}


class PkGlobalTopology(C.Structure):
    """Mirror of `struct pk_global_topology` — the static arrays of System.odeint_args()
    (reference global_model/network.py:508-526)."""
    _fields_ = [
        ("model", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("n_bins", C.c_int32),
        ("n_sites", C.c_void_p),
        ("W_indptr", C.c_void_p), ("W_indices", C.c_void_p), ("W_data", C.c_void_p),
        ("TF_indptr", C.c_void_p), ("TF_indices", C.c_void_p), ("TF_data", C.c_void_p),
        ("kin_grid", C.c_void_p), ("kin_Kmat", C.c_void_p), ("tf_deg", C.c_void_p),
        ("driver_map", C.c_void_p), ("force_generic_schur", C.c_int32), ("reserved0", C.c_int32),
    ]


class PkGlobalLossData(C.Structure):
    """Mirror of `struct pk_global_loss_data` — LOSS_FN's argument list minus Y
    (reference global_model/lossfn.py:113-121)."""
    _fields_ = [
        ("n_prot", C.c_int32), ("n_rna", C.c_int32), ("n_pho", C.c_int32),
        ("p_prot", C.c_void_p), ("t_prot", C.c_void_p), ("obs_prot", C.c_void_p), ("w_prot", C.c_void_p),
        ("p_rna", C.c_void_p), ("t_rna", C.c_void_p), ("obs_rna", C.c_void_p), ("w_rna", C.c_void_p),
        ("p_pho", C.c_void_p), ("s_pho", C.c_void_p), ("t_pho", C.c_void_p), ("obs_pho", C.c_void_p),
        ("w_pho", C.c_void_p),
        ("prot_base_idx", C.c_int32), ("rna_base_idx", C.c_int32), ("pho_base_idx", C.c_int32),
    ]


class PkGlobalJob(C.Structure):
    """Mirror of `struct pk_global_job`."""
    _fields_ = [
        ("topo", C.c_int32), ("memspace", C.c_int32), ("B", C.c_int64), ("T", C.c_int32),
        ("theta_mode", C.c_int32),
        ("params", C.c_void_p), ("y0", C.c_void_p), ("y0_stride", C.c_int64), ("t_eval", C.c_void_p),
        ("rtol", C.c_double), ("atol", C.c_double),
        ("max_steps", C.c_int32), ("loss_mode", C.c_int32), ("metric", C.c_int32),
        ("n_mt_prot", C.c_int32), ("n_mt_rna", C.c_int32), ("n_mt_pho", C.c_int32),
        ("mt_prot", C.c_void_p), ("mt_rna", C.c_void_p), ("mt_pho", C.c_void_p),
        ("mb_prot", C.c_int32), ("mb_rna", C.c_int32), ("mb_pho", C.c_int32), ("reserved0", C.c_int32),
        ("lambdas", C.c_double * 3), ("lambda_prior", C.c_double),
        ("out_Y", C.c_void_p), ("out_loss", C.c_void_p), ("out_F", C.c_void_p), ("out_metric", C.c_void_p),
        ("out_status", C.c_void_p), ("out_nsteps", C.c_void_p), ("out_nrej", C.c_void_p), ("out_fc", C.c_void_p),
    ]


GLOBAL_METRIC_IDS = {"total_signal": 0, "mean": 1, "variance": 2, "l2_norm": 3}

# every symbol include/phoskin_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pk_abi_version": (C.c_int, []),
    "pk_last_error": (C.c_char_p, []),
    "pk_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "pk_destroy": (C.c_int, [C.c_void_p]),
    "pk_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pk_device_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    "pk_local_dims": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int)]),
    "pk_local_job_init": (None, [C.POINTER(PkLocalJob)]),
    "pk_sizeof_local_job": (C.c_int, []),
    "pk_region_begin": (C.c_int, [C.c_void_p]),
    "pk_region_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "pk_local_solve_batch": (C.c_int, [C.c_void_p, C.POINTER(PkLocalJob)]),
    "pk_last_launch_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "pk_morris_ee": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                               C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pk_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "pk_host_free": (C.c_int, [C.c_void_p]),
    "pk_measure_fp64_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "pk_ros5l_coeffs": (C.c_int, [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "pk_ros6l_coeffs": (C.c_int, [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "pk_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "pk_nccl_init": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "pk_local_solve_allgather": (C.c_int, [C.c_void_p, C.POINTER(PkLocalJob), C.c_int32, C.c_int32, C.c_void_p]),
    "pk_allgather_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pk_global_solve_custom": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "pk_global_rhs_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pk_sym_alloc": (C.c_int, [C.c_void_p, C.c_int64, C.c_char_p]),
    "pk_sym_open": (C.c_int, [C.c_void_p, C.c_char_p]),
    "pk_sym_buffer": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "pk_sym_free": (C.c_int, [C.c_void_p]),
    "pk_local_solve_gather_p2p": (C.c_int, [C.c_void_p, C.POINTER(PkLocalJob), C.c_int32]),
    "pk_nlls_job_init": (None, [C.POINTER(PkNllsJob)]),
    "pk_sizeof_nlls_job": (C.c_int, []),
    "pk_local_nlls_batch": (C.c_int, [C.c_void_p, C.POINTER(PkNllsJob)]),
    "pk_global_upload": (C.c_int, [C.c_void_p, C.POINTER(PkGlobalTopology), C.POINTER(C.c_int32)]),
    "pk_global_set_loss_data": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(PkGlobalLossData)]),
    "pk_global_set_prior": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "pk_global_release": (C.c_int, [C.c_void_p, C.c_int32]),
    "pk_global_dims": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pk_global_counts": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pk_global_job_init": (None, [C.POINTER(PkGlobalJob)]),
    "pk_sizeof_global_job": (C.c_int, []),
    "pk_global_solve_batch": (C.c_int, [C.c_void_p, C.POINTER(PkGlobalJob)]),
    "pk_global_loss_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                       C.c_void_p]),
}

_lib = None


class PhoskinError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every prototype. Raises if it is missing:
    there is deliberately no pure-Python / CPU substitute on the product path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PhoskinError(
            f"{LIB_PATH} not found — build it with `make` (or __graft_entry__.build()); "
            "phoskintime_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.pk_abi_version() != 1:
        raise PhoskinError("libphoskin_b200.so ABI version mismatch")
    if lib.pk_sizeof_local_job() != C.sizeof(PkLocalJob):
        raise PhoskinError("pk_local_job layout mismatch between header and ctypes mirror")
    if lib.pk_sizeof_nlls_job() != C.sizeof(PkNllsJob):
        raise PhoskinError("pk_nlls_job layout mismatch between header and ctypes mirror")
    if lib.pk_sizeof_global_job() != C.sizeof(PkGlobalJob):
        raise PhoskinError("pk_global_job layout mismatch between header and ctypes mirror")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().pk_last_error()
        raise PhoskinError(msg.decode() if msg else f"pk call failed with {rc}")
