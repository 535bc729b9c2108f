"""ctypes signatures of the native engine entry points (include/ntransformer.h, include/nt_b200_engine.h)."""
import ctypes as C

_vp, _i, _f = C.c_void_p, C.c_int, C.c_float

ENGINE_SIGNATURES = {}


def apply(lib):
    for name, (res, args) in ENGINE_SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype, fn.argtypes = res, args
