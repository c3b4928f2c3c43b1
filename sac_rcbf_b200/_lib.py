"""ctypes binding of librcbf_b200.so (C ABI in include/rcbf_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or does not load, importing the compute entry points
raises.  Nothing here (or anywhere in this package) imports the CPU oracle.
"""
import ctypes as C
import os

from . import _params as P

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RCBF_LIB_PATH", os.path.join(HERE, "librcbf_b200.so"))  # env override: A/B builds

_f = C.POINTER(C.c_float)
_d = C.POINTER(C.c_double)
_vp = C.c_void_p
_i64 = C.c_int64

# name -> argtypes; every symbol include/rcbf_b200.h declares (tests/test_abi.py cross-checks against the header)
SIGNATURES = {
    "rcbf_unicycle_assemble": [_vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams), _vp, _vp, _vp],
    "rcbf_cars_assemble": [_vp, _vp, _vp, _i64, C.POINTER(P.CarsParams), _vp, _vp, _vp],
    "rcbf_unicycle_safe_action": [_vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams), _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp],
    "rcbf_cars_safe_action": [_vp, _vp, _vp, _i64, C.POINTER(P.CarsParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "rcbf_unicycle_safe_action_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams), _vp,
                                      _vp],
    "rcbf_cars_safe_action_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(P.CarsParams), _vp, _vp],
    "rcbf_unicycle_safe_action_saved": [_vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams), _vp, _vp, _vp, _vp],
    "rcbf_cars_safe_action_saved": [_vp, _vp, _vp, _i64, C.POINTER(P.CarsParams), _vp, _vp, _vp, _vp],
    "rcbf_unicycle_safe_action_bwd_meta": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams), _vp, _vp],
    "rcbf_cars_safe_action_bwd_meta": [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(P.CarsParams), _vp, _vp],
    "rcbf_qp_solve": [_vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "rcbf_qp_solve_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp],
    "rcbf_unicycle_safe_step": [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams),
                                C.POINTER(P.UnicycleEnvParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "rcbf_cars_safe_step": [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(P.CarsParams), C.POINTER(P.CarsEnvParams), _vp,
                            _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "rcbf_unicycle_safe_action_host": [_vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams), _vp,
                                       C.POINTER(C.c_int32), C.c_int, C.c_int],
    "rcbf_cars_safe_action_host": [_vp, _vp, _vp, _i64, C.POINTER(P.CarsParams), _vp, C.POINTER(C.c_int32), C.c_int,
                                   C.c_int],
    "rcbf_unicycle_safe_step_host": [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams),
                                     C.POINTER(P.UnicycleEnvParams), _vp, _vp, _vp, _vp, _vp, _vp,
                                     C.POINTER(C.c_int32), C.c_int, C.c_int],
    "rcbf_unicycle_safe_step_host_gp": [_vp, _vp, _vp, C.POINTER(P.GpPosterior), _i64, C.POINTER(P.UnicycleParams),
                                        C.POINTER(P.UnicycleEnvParams), _vp, _vp, _vp, _vp, _vp, _vp,
                                        C.POINTER(C.c_int32), C.c_int, C.c_int],
    "rcbf_cars_safe_step_host": [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(P.CarsParams), C.POINTER(P.CarsEnvParams),
                                 _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_int32), C.c_int, C.c_int],
    "rcbf_unicycle_assemble_f64": [_vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams), C.c_int, _vp, _vp, _vp],
    "rcbf_cars_assemble_f64": [_vp, _vp, _vp, _i64, C.POINTER(P.CarsParams), C.c_int, _vp, _vp, _vp],
    "rcbf_unicycle_safe_action_general": [_vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams), _vp, C.c_int, _vp, _vp,
                                          _vp, _vp, _vp],
    "rcbf_unicycle_safe_action_bwd_general": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams), _vp,
                                              C.c_int, _vp, _vp],
    "rcbf_unicycle_assemble_general": [_vp, _vp, _vp, _vp, _i64, C.POINTER(P.UnicycleParams), _vp, C.c_int, _vp, _vp,
                                       _vp],
    "rcbf_counters_publish": [_vp, _vp, C.c_uint64, _vp],
    "rcbf_stream_synchronize": [_vp],
    "rcbf_counters_bind_mirror": [_vp, _vp, _vp],
    "rcbf_fp32_fma_probe": [_vp, C.c_int, C.c_int, C.c_int, _vp],
    "rcbf_fp64_fma_probe": [_vp, C.c_int, C.c_int, C.c_int, _vp],
    "rcbf_gp_predict_f32": [_vp, _i64, C.POINTER(P.GpPosterior), _vp, _vp, _vp],
    "rcbf_gp_predict_f64": [_vp, _i64, C.POINTER(P.GpPosterior), _vp, _vp, _vp],
    "rcbf_replay_push": [C.POINTER(P.ReplayRing), _i64, C.POINTER(_vp * 7), _i64, _vp],
    "rcbf_replay_sample": [C.POINTER(P.ReplayRing), _i64, _i64, C.c_uint64, C.POINTER(_vp * 7), _vp, _vp],
}
for _suf in ("f32", "f64"):
    SIGNATURES["rcbf_unicycle_env_reset_" + _suf] = [_vp, _vp, _vp, _i64, C.POINTER(P.UnicycleEnvParams), _vp, _vp]
    SIGNATURES["rcbf_unicycle_env_step_" + _suf] = [_vp, _vp, _vp, _i64, C.POINTER(P.UnicycleEnvParams), _vp, _vp, _vp,
                                                    _vp, _vp, _vp]
    SIGNATURES["rcbf_cars_env_reset_" + _suf] = [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]
    SIGNATURES["rcbf_cars_env_step_" + _suf] = [_vp, _vp, _vp, _vp, _i64, C.POINTER(P.CarsEnvParams), _vp, _vp, _vp,
                                                _vp, _vp]
    SIGNATURES["rcbf_unicycle_predict_next_" + _suf] = [_vp, _vp, _vp, _i64, C.c_double, _vp, _vp]
    SIGNATURES["rcbf_unicycle_rollout_step_" + _suf] = [_vp, _vp, _vp, _vp, _vp, _i64, C.c_double, C.c_double,
                                                        C.c_double, _vp, _vp, _vp, _vp]
    SIGNATURES["rcbf_cars_rollout_step_" + _suf] = [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.c_double, C.c_double,
                                                    C.c_double, C.c_int, _vp, _vp, _vp, _vp, _vp]
    SIGNATURES["rcbf_cars_predict_next_" + _suf] = [_vp, _vp, _vp, _vp, _i64, C.c_double, C.c_double, C.c_double, _vp,
                                                    _vp]

_lib = None


class RcbfLibraryError(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises RcbfLibraryError if it is absent (run `python -m sac_rcbf_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RcbfLibraryError(
            "librcbf_b200.so not found at %s -- build it with `python -m sac_rcbf_b200.build` "
            "(there is no CPU fallback for the safety path)" % LIB_PATH)
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:
        raise RcbfLibraryError("could not load %s: %s" % (LIB_PATH, e)) from e
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: loud by design
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.rcbf_version.restype = C.c_char_p
    lib.rcbf_version.argtypes = []
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed with CUDA error code %d" % (what, rc))


def ptr(t):
    """data pointer of a torch tensor as a plain int (or None -> NULL); ctypes converts it for the `void*` argtypes."""
    return None if t is None else t.data_ptr()


_raw_stream = None


def stream_ptr(device):
    """cudaStream_t of torch's current stream on `device` as an int (0 -> None: the legacy default stream)."""
    global _raw_stream
    if _raw_stream is None:
        import torch

        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or \
            (lambda idx: torch.cuda.current_stream(idx).cuda_stream)
    idx = device.index if hasattr(device, "index") else device
    if idx is None:
        import torch

        idx = torch.cuda.current_device()
    return _raw_stream(idx) or None


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RcbfLibraryError("sac_rcbf_b200 needs a CUDA device (sm_100a): torch.cuda.is_available() is False and "
                               "there is no CPU fallback for the safety path")
