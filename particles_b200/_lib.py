"""ctypes binding of libsmcb.so (the C-ABI declared in include/smcb.h).

The product path has NO CPU fallback: if the library is missing or no CUDA device
is visible, the calls below raise instead of computing somewhere else.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SMCB_LIB") or os.path.join(HERE, "libsmcb.so")   # SMCB_LIB: kernel-variant experiments

SMCB_MAX_PARAMS = 256
SUMMARY_STRIDE = 4

RS_CODES = {"multinomial": 0, "stratified": 1, "systematic": 2, "residual": 3, "ssp": 4}
FUSED_SCHEMES = ("multinomial", "stratified", "systematic")     # schemes built into the fused step kernel
FK_BOOTSTRAP, FK_GUIDED, FK_APF, FK_AUXBOOT = 0, 1, 2, 3
MODEL_STOCHVOL, MODEL_LINGAUSS, MODEL_GORDON, MODEL_THETALOGISTIC = 0, 1, 2, 3
MODEL_BEARINGS, MODEL_MVLINGAUSS, MODEL_DISCRETECOX, MODEL_STOCHVOLLEV = 4, 5, 6, 7
LSE_SUM, LSE_MEAN, LSE_ESSL = 0, 1, 2

c_dp = C.c_void_p  # device pointers travel as integers


class FilterDesc(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("fk", C.c_int32), ("scheme", C.c_int32), ("dim", C.c_int32),
        ("dy", C.c_int32), ("n_params", C.c_int32), ("world", C.c_int32), ("rank", C.c_int32),
        ("n", C.c_int64), ("n_global", C.c_int64), ("index_offset", C.c_int64), ("T", C.c_int64),
        ("essrmin", C.c_double), ("seed", C.c_uint64),
        ("params", C.c_double * SMCB_MAX_PARAMS),
        ("X", c_dp * 2), ("lw", c_dp * 2), ("A", c_dp), ("cdf", c_dp), ("data", c_dp),
        ("summaries", c_dp), ("z_in", c_dp), ("u_in", c_dp), ("scratch", c_dp),
        ("step_consts", c_dp), ("local_stats", c_dp), ("gathered", c_dp),
        ("mail_local", c_dp), ("mail_peer", c_dp * 8),
        ("rs_global", C.c_int32), ("reserved0", C.c_int32), ("moments", c_dp), ("reserved1", c_dp),
        ("peer_X0", c_dp * 8), ("peer_X1", c_dp * 8), ("peer_cdf", c_dp * 8),
    ]


# name -> (restype, argtypes): every symbol include/smcb.h declares
PROTOTYPES = {
    "smcb_last_error": (C.c_char_p, []),
    "smcb_version": (C.c_int, []),
    "smcb_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_uint64]),
    "smcb_destroy": (C.c_int, [C.c_void_p]),
    "smcb_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "smcb_seed": (C.c_int, [C.c_void_p, C.c_uint64]),
    "smcb_launch_count": (C.c_int64, [C.c_void_p]),
    "smcb_normalise": (C.c_int, [C.c_void_p, c_dp, C.c_int64, c_dp, c_dp]),
    "smcb_weights_from_stats": (C.c_int, [C.c_void_p, c_dp, C.c_int64, c_dp, c_dp]),
    "smcb_lse": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, C.c_int64, c_dp]),
    "smcb_exp_and_normalise": (C.c_int, [C.c_void_p, c_dp, C.c_int64, c_dp]),
    "smcb_wmean_and_var": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int64, C.c_int, c_dp]),
    "smcb_cumsum": (C.c_int, [C.c_void_p, c_dp, C.c_int64, c_dp]),
    "smcb_searchsorted": (C.c_int, [C.c_void_p, c_dp, C.c_int64, c_dp, C.c_int64, c_dp]),
    "smcb_resample_scratch_doubles": (C.c_int64, [C.c_int64, C.c_int64]),
    "smcb_resample": (C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_int64, C.c_int64, c_dp, c_dp, c_dp]),
    "smcb_gather": (C.c_int, [C.c_void_p, c_dp, C.c_int64, c_dp, C.c_int64, C.c_int, c_dp]),
    "smcb_gather_rows": (C.c_int, [C.c_void_p, c_dp, C.c_int64, c_dp, C.c_int64, C.c_int, c_dp]),
    "smcb_normal_rvs": (C.c_int, [C.c_void_p, c_dp, C.c_double, c_dp, C.c_double, c_dp, c_dp, C.c_int64]),
    "smcb_normal_logpdf": (C.c_int, [C.c_void_p, c_dp, C.c_double, c_dp, C.c_double, c_dp, C.c_double,
                                     c_dp, C.c_int64]),
    "smcb_logpdf1": (C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_double, C.c_double, C.c_double, c_dp, C.c_double, c_dp,
                               C.c_double, c_dp, C.c_int64]),
    "smcb_mvnormal_rvs": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_int64]),
    "smcb_mvnormal_logpdf": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int, c_dp,
                                       C.c_int64]),
    "smcb_standard_normal": (C.c_int, [C.c_void_p, c_dp, C.c_int64]),
    "smcb_uniform": (C.c_int, [C.c_void_p, c_dp, C.c_int64]),
    "smcb_logistic_target": (C.c_int, [C.c_void_p, c_dp, C.c_int64, C.c_int, c_dp, C.c_int64, C.c_double,
                                       C.c_double, c_dp, c_dp, c_dp]),
    "smcb_logistic_wf_move": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp,
                                        C.c_int64, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                        c_dp, c_dp]),
    "smcb_rw_propose": (C.c_int, [C.c_void_p, c_dp, C.c_int64, C.c_int, c_dp, c_dp, c_dp]),
    "smcb_mh_accept": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                 c_dp, c_dp]),
    "smcb_next_annealing_epn": (C.c_int, [C.c_void_p, c_dp, C.c_int64, C.c_double, C.c_double, c_dp]),
    "smcb_rw_calibrate": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int64, C.c_int, C.c_double, c_dp]),
    "smcb_essl_grid": (C.c_int, [C.c_void_p, c_dp, C.c_int64, C.c_double, C.c_double, c_dp, c_dp]),
    "smcb_wcov_sums": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int64, C.c_int, c_dp, c_dp]),
    "smcb_chol_from_sums": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_double, c_dp]),
    "smcb_device_math": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, C.c_int64]),
    "smcb_filter_create": (C.c_int, [C.c_void_p, C.POINTER(FilterDesc), C.POINTER(C.c_void_p)]),
    "smcb_filter_destroy": (C.c_int, [C.c_void_p]),
    "smcb_filter_step": (C.c_int, [C.c_void_p, C.c_int64]),
    "smcb_filter_step_local": (C.c_int, [C.c_void_p]),
    "smcb_filter_step_finish": (C.c_int, [C.c_void_p]),
    "smcb_p2p_alloc": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    "smcb_p2p_open": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "smcb_p2p_close": (C.c_int, [C.c_void_p]),
    "smcb_p2p_free": (C.c_int, [C.c_void_p]),
    "smcb_filter_step_timed": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_double)]),
    "smcb_filter_state": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "smcb_measure_fp64_peak": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(C.c_double)]),
    "smcb_measure_stream_peak": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int64, C.POINTER(C.c_double)]),
}

_lib = None


class SmcbError(RuntimeError):
    pass


def load():
    """Load libsmcb.so and bind every prototype (no GPU needed for this)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise SmcbError(
                f"{SO_PATH} is missing: build it with `python -m particles_b200.build` "
                "(there is no CPU fallback for this path)")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)   # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc):
    if rc == 0:
        return
    msg = load().smcb_last_error().decode()
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise NotImplementedError(msg)
    raise SmcbError(msg)
