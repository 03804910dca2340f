"""ctypes binding of libflemingo_b200.so (include/flemingo_b200.h).

The library is built in-tree by `__graft_entry__.build()` into flemingo_b200/_lib/.  Loading fails loudly:
there is no Python or CPU fallback for any entry point.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "_lib" / "libflemingo_b200.so"

FL_OK = 0
FL_E_CUDA, FL_E_ARG, FL_E_TOO_LARGE, FL_E_NAN_NULL, FL_E_UNSUPPORTED, FL_E_STATE = -1, -2, -3, -4, -5, -6
FL_WANT_TRACE = 1
FL_EXACT_ONLY = 2
FL_MAX_SETS = 8
FIT_METHODS = {"Welch": 0, "Yuen": 1, "Trim-left-M": 2, "Trim-left-M-SD": 3}

# every symbol include/flemingo_b200.h declares (tests check that the .so exports exactly these)
EXPORTED = [
    "fl_last_error", "fl_abi_version", "fl_device_count", "fl_ctx_create", "fl_ctx_destroy", "fl_ctx_stream", "fl_sync",
    "fl_upload_sequences", "fl_sequence_classes", "fl_upload_population", "fl_update_connectors", "fl_place_batch", "fl_fitness",
    "fl_fitness_dev",
    "fl_calculate", "fl_launch_count", "fl_path_stats", "fl_measure_dpx_rate",
    "fl_build_connector_tables", "fl_round_decimals", "fl_build_shape_masses", "fl_fitness_finalize",
]


class FlPopulation(C.Structure):
    _fields_ = [
        ("n_org", C.c_int), ("n_rec_total", C.c_int),
        ("rec_begin", C.c_void_p), ("rec_type", C.c_void_p), ("rec_len", C.c_void_p), ("rec_nb", C.c_void_p),
        ("rec_data", C.c_void_p), ("pool_begin", C.c_void_p), ("rec_pool", C.c_void_p), ("n_pool", C.c_longlong),
        ("con_M", C.c_void_p), ("con_pool", C.c_void_p), ("n_con", C.c_longlong),
        ("rec_class", C.c_void_p), ("n_classes", C.c_int), ("cls_kind", C.c_void_p), ("cls_len", C.c_void_p),
        ("cls_nb", C.c_void_p), ("cls_edges", C.c_void_p), ("edges_pool", C.c_void_p), ("n_edges", C.c_longlong),
    ]


class FlemingoError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[flemingo_b200 {code}] {message}")
        self.code = code
        self.message = message


_lib = None


def load_library() -> C.CDLL:
    """Loads the shared library (no CUDA call is made until a context is created)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). flemingo_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    vp, ip, ll, dp = C.c_void_p, C.c_int, C.c_longlong, C.c_double
    lib.fl_last_error.restype = C.c_char_p
    lib.fl_last_error.argtypes = []
    lib.fl_abi_version.restype = ip
    lib.fl_device_count.restype = ip
    lib.fl_device_count.argtypes = []
    lib.fl_ctx_create.argtypes = [ip, vp, ll, C.POINTER(vp)]
    lib.fl_ctx_destroy.argtypes = [vp]
    lib.fl_ctx_stream.restype = vp
    lib.fl_ctx_stream.argtypes = [vp]
    lib.fl_sync.argtypes = [vp]
    lib.fl_upload_sequences.argtypes = [vp, ip, vp, vp, ip]
    lib.fl_sequence_classes.argtypes = [vp, ip, C.POINTER(ip), vp]
    lib.fl_upload_population.argtypes = [vp, C.POINTER(FlPopulation)]
    lib.fl_update_connectors.argtypes = [vp, vp, ll]
    lib.fl_place_batch.argtypes = [vp, ip, vp, ip, vp, vp, vp, vp]
    lib.fl_fitness.argtypes = [vp, ip, ip, ip, dp, vp]
    lib.fl_fitness_dev.argtypes = [vp, ip, ip, ip, dp, ip, C.POINTER(vp)]
    lib.fl_calculate.argtypes = [vp, vp, ip, vp, ip, vp, vp, vp, ll, vp, vp, vp, ip, vp, vp, vp]
    lib.fl_measure_dpx_rate.argtypes = [vp, C.POINTER(dp), C.POINTER(dp), C.POINTER(ip), C.POINTER(ip)]
    lib.fl_build_connector_tables.argtypes = [vp, vp, ip, ip, dp, vp, vp, ip]
    lib.fl_round_decimals.argtypes = [vp, ll, ip]
    lib.fl_build_shape_masses.argtypes = [vp, vp, vp, vp, vp, ip, dp, vp, vp]
    lib.fl_fitness_finalize.argtypes = [vp, ll]
    lib.fl_path_stats.argtypes = [vp, C.POINTER(ll), C.POINTER(ll), C.POINTER(ll)]
    lib.fl_launch_count.restype = ll
    lib.fl_launch_count.argtypes = [vp]
    for name in EXPORTED:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("fl_abi_version",):
            fn.restype = ip
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Maps ABI return codes to exceptions the way the reference's callers expect them."""
    if rc == FL_OK:
        return
    msg = load_library().fl_last_error().decode("utf-8", "replace")
    if rc == FL_E_TOO_LARGE:
        # organism_object.py:1279-1282 raises ValueError before the C call would set RuntimeError
        raise ValueError(msg)
    raise FlemingoError(rc, msg)
