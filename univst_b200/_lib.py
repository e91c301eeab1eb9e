"""ctypes binding of ``libunivst_b200.so`` (the C ABI declared in ``include/univst_b200.h``).

There is no CPU or PyTorch fallback: if the shared library is missing, or the device is not sm_100, every
entry point raises.  ``lib()`` loads lazily so that importing the package on a GPU-less host (CI, the build
check) works.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libunivst_b200.so")

_lib = None
_device_ok = None


class UnivstError(RuntimeError):
    pass


class Epilogue(C.Structure):
    """Mirror of ``univst_epilogue_t``."""

    _fields_ = [
        ("bias", C.c_void_p),
        ("rowvec", C.c_void_p),
        ("rows_per_group", C.c_int32),
        ("rowvec_ld", C.c_int32),
        ("act", C.c_int32),
        ("residual", C.c_void_p),
        ("ldr", C.c_int32),
        ("bias2", C.c_void_p),
        ("geglu", C.c_int32),
        ("out_scale", C.c_float),
    ]


_vp, _i32, _f32, _i64 = C.c_void_p, C.c_int32, C.c_float, C.c_int64

# name -> argtypes (restype is always int unless listed in _RESTYPES)
PROTOTYPES = {
    "univst_abi_version": [],
    "univst_last_error": [],
    "univst_device_check": [],
    "univst_gemm_f16": [_vp, _i32, _vp, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _i32, C.POINTER(Epilogue), _vp],
    "univst_conv3x3_f16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i32, C.POINTER(Epilogue), _vp],
}
_RESTYPES = {"univst_last_error": C.c_char_p}


def register(name, argtypes, restype=None):
    PROTOTYPES[name] = argtypes
    if restype is not None:
        _RESTYPES[name] = restype
    if _lib is not None:  # library already loaded: attach immediately
        fn = getattr(_lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)


def lib():
    """Load the shared library (once) and attach prototypes.  Raises if it was never built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UnivstError(
                f"{LIB_PATH} not found: build it with `python -m univst_b200.build` "
                "(there is no CPU / PyTorch fallback for the sm_100a kernels)"
            )
        handle = C.CDLL(LIB_PATH)
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the header and the library disagree
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = handle
    return _lib


def check(code: int, what: str = ""):
    if code != 0:
        msg = lib().univst_last_error()
        raise UnivstError(f"{what} failed ({code}): {msg.decode() if msg else '?'}")


def require_device():
    """Fail loudly unless the current CUDA device is a Blackwell (sm_100) part."""
    global _device_ok
    if _device_ok is None:
        check(lib().univst_device_check(), "univst_device_check")
        _device_ok = True
