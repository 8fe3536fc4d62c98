"""ctypes binding of libkgb200.so (the C ABI declared in include/kgb200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry point is reached without an
sm_100 GPU, the call raises.  Only the CUDA path exists.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkgb200.so")

# dtype codes (kgb200.h)
F64, I64, F32, I32, B8 = 0, 1, 2, 3, 4
ARG_ARRAY, ARG_SCALAR, ARG_DEVSCALAR = 0, 1, 2
MODE_MAP, MODE_REDUCE, MODE_SCAN, MODE_SCANRED = 0, 1, 2, 3
LAUNCH_WS_CLEAN = 1  # kgb_expr_launch_ex: the workspace was zero-filled once and is only ever used by this library
RED_ADD, RED_MUL, RED_MAX, RED_MIN = 0, 1, 2, 3


class KgbError(RuntimeError):
    """A libkgb200 entry point returned a non-zero status."""


_c = ctypes
_vp, _i32, _i64, _sz = _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_size_t
_pi32, _pi64, _psz = _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int64), _c.POINTER(_c.c_size_t)

# name -> (restype, argtypes); every symbol include/kgb200.h declares
SIGNATURES = {
    "kgb_abi_version": (_i32, []),
    "kgb_last_error": (_c.c_char_p, []),
    "kgb_init": (_i32, [_i32]),
    "kgb_device_sm_count": (_i32, [_i32, _pi32]),
    "kgb_expr_compile": (_i32, [_pi32, _i32, _pi32, _i32, _pi32, _pi32, _i32, _i32, _i32, _i32,
                                _c.POINTER(_vp), _pi32]),
    "kgb_expr_compile_epi": (_i32, [_pi32, _i32, _pi32, _i32, _pi32, _i32, _pi32, _pi32, _i32, _i32, _i32, _i32,
                                    _c.POINTER(_vp), _pi32]),
    "kgb_expr_check": (_i32, [_pi32, _i32, _pi32, _i32, _pi32, _pi32, _i32, _i32, _i32, _i32, _pi32,
                              _c.POINTER(_c.c_char_p)]),
    "kgb_expr_workspace_bytes": (_i32, [_vp, _i64, _psz]),
    "kgb_expr_launch": (_i32, [_vp, _c.POINTER(_vp), _i64, _vp, _vp, _sz, _vp]),
    "kgb_expr_launch_ex": (_i32, [_vp, _c.POINTER(_vp), _i64, _vp, _vp, _sz, _vp, _i32]),
    "kgb_expr_infer": (_i32, [_pi32, _i32, _pi32, _pi32, _i32, _pi32]),
    "kgb_expr_out_dtype2": (_i32, [_vp]),
    "kgb_expr_source": (_c.c_char_p, [_vp]),
    "kgb_expr_kernel_name": (_c.c_char_p, [_vp]),
    "kgb_argsort_workspace_bytes": (_i32, [_i32, _i64, _psz]),
    "kgb_argsort": (_i32, [_vp, _i32, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "kgb_group_workspace_bytes": (_i32, [_i32, _i64, _psz]),
    "kgb_group": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "kgb_select_workspace_bytes": (_i32, [_i64, _psz]),
    "kgb_find_equal": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "kgb_nonzero": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp, _sz, _vp]),
    "kgb_unique_first_workspace_bytes": (_i32, [_i32, _i64, _psz]),
    "kgb_unique_first": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "kgb_iota_i64": (_i32, [_vp, _i64, _i64, _i64, _vp]),
    "kgb_gather": (_i32, [_vp, _i32, _i64, _vp, _i64, _vp, _vp, _vp]),
    "kgb_expand": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "kgb_reverse": (_i32, [_vp, _i32, _i64, _vp, _vp]),
    "kgb_array_equal": (_i32, [_vp, _vp, _i32, _i64, _vp, _vp]),
    "kgb_groupagg_workspace_bytes": (_i32, [_i64, _i64, _psz]),
    "kgb_groupagg_f64": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _sz, _vp]),
    "kgb_groupagg_f64_async": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _sz, _vp]),
    "kgb_groupagg_status": (_i32, [_vp, _i64, _vp]),
    "kgb_groupagg_merge": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "kgb_groupagg_fold": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "kgb_groupagg_to_owners": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "kgb_groupagg_from_owners": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "kgb_brc_workspace_bytes": (_i32, [_i64, _psz]),
    "kgb_brc_scan": (_i32, [_vp, _i64, _pi64, _vp, _sz, _vp]),
    "kgb_brc_estimate": (_i32, [_vp, _i64, _pi64, _vp, _sz, _vp]),
    "kgb_brc_parse": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _pi64, _pi64, _vp, _sz, _vp]),
}

_lib = None
_lock = threading.Lock()
_tls = threading.local()


def load():
    """dlopen libkgb200.so and bind every declared symbol.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise KgbError(
                f"{LIB_PATH} is missing: build it with `python -m klongpy_b200.csrc.build` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        if lib.kgb_abi_version() != 1:
            raise KgbError("libkgb200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


# Test seam: the CPU-only host-logic tests (tests/fake_lib.py) register a test double here so the provider,
# lowering and verb overrides can be exercised through the real KlongPy interpreter on a box without a GPU.
# The product never sets it: DeviceArrays on a non-CUDA device fail loudly.
_host_double = None


def load_for_host_device(device):
    if _host_double is None:
        raise KgbError(f"DeviceArray on '{device}': klongpy_b200 computes on CUDA devices only "
                       "(no CPU fallback)")
    return _host_double


def last_error():
    lib = _lib if _lib is not None else (_host_double or load())
    return lib.kgb_last_error().decode(errors="replace")


def check(rc, what=""):
    if rc != 0:
        raise KgbError(f"{what}: {last_error()} (status {rc})" if what else f"{last_error()} (status {rc})")


def bind_device(index):
    """kgb_init once per (thread, device): the library keeps a thread-local device binding."""
    lib = load()
    if getattr(_tls, "device", None) != index:
        check(lib.kgb_init(index), "kgb_init")
        _tls.device = index
    return lib
