"""ctypes binding of libnt_b200.so (include/nt_b200.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# NT_B200_LIB overrides the library path (A/B runs of kernel variants); the default is the in-tree build.
LIB_PATH = Path(os.environ.get("NT_B200_LIB") or (Path(__file__).resolve().parent / "libnt_b200.so"))
_lib = None

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes): every symbol include/nt_b200.h declares
SIGNATURES = {
    "nt_b200_rmsnorm": (None, [_vp, _vp, _vp, _i, _i, _f, _vp]),
    "nt_b200_rmsnorm_f16": (None, [_vp, _vp, _vp, _i, _i, _f, _vp]),
    "nt_b200_rope": (None, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "nt_b200_softmax": (None, [_vp, _vp, _i, _i, _vp]),
    "nt_b200_masked_softmax": (None, [_vp, _vp, _vp, _i, _i, _vp]),
    "nt_b200_gemv": (None, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "nt_b200_gemv_add": (None, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "nt_b200_gemm_f32": (None, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "nt_b200_silu_mul": (None, [_vp, _vp, _vp, _i, _vp]),
    "nt_b200_add_bias": (None, [_vp, _vp, _i, _vp]),
    "nt_b200_attention_decode": (None, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "nt_b200_attention_prefill": (None, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "nt_b200_copy_to_kv_cache": (None, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "nt_b200_add": (None, [_vp, _vp, _vp, _i, _vp]),
    "nt_b200_add_inplace": (None, [_vp, _vp, _i, _vp]),
    "nt_b200_copy": (None, [_vp, _vp, _i, _vp]),
    "nt_b200_cosine_similarity": (None, [_vp, _vp, _vp, _i, _vp]),
    "nt_cuda_malloc": (_vp, [_sz]),
    "nt_cuda_free": (None, [_vp]),
    "nt_cuda_memcpy_h2d": (None, [_vp, _vp, _sz]),
    "nt_cuda_memcpy_d2h": (None, [_vp, _vp, _sz]),
    "nt_cuda_memcpy_d2d": (None, [_vp, _vp, _sz]),
    "nt_cuda_memset": (None, [_vp, _i, _sz]),
    "nt_cuda_malloc_host": (_vp, [_sz]),
    "nt_cuda_free_host": (None, [_vp]),
    "nt_b200_xq_bytes": (_sz, [_i]),
    "nt_b200_quantize_x": (None, [_vp, _vp, _i, _vp]),
    "nt_b200_gemv_fused": (_i, [_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i, _vp, _i, _vp]),
    "nt_b200_gemv_fused_f32": (_i, [_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i, _vp, _vp, _f, _i, _vp]),
    "nt_b200_attention_decode_scratch_floats": (_sz, [_i, _i, _i, _i]),
    "nt_b200_attention_decode_tickets": (_i, [_i, _i]),
    "nt_b200_attention_decode_fused": (None, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "nt_b200_embed_rows": (None, [_vp, _vp, _i, _vp, _i, _i, _vp]),
    "nt_b200_gemm_f16_tc_workspace_bytes": (_sz, [_i, _i]),
    "nt_b200_gemm_f16_tc": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "nt_b200_split_activations": (None, [_vp, _vp, _i, _i, _vp]),
    "nt_b200_gemm_f16_tc_ws": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nt_b200_gemm_f16_tc_swiglu_ws": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "nt_b200_rmsnorm_split": (None, [_vp, _vp, _vp, _i, _i, C.c_float, _vp]),
    "nt_b200_dequant_split": (None, [_vp, _vp, _vp, _i, _sz, _i, _i, _vp]),
    "nt_b200_launch_count": (C.c_ulonglong, []),
    "nt_b200_stream_sync": (_i, [_vp]),
    "nt_b200_version": (C.c_char_p, []),
}

# C++ launcher names the reference's host code links against (src/cuda/kernels.h:10-74), Itanium-mangled.
CXX_LAUNCHERS = [
    "launch_rmsnorm", "launch_rmsnorm_f16", "launch_rope", "launch_softmax", "launch_masked_softmax", "launch_gemv",
    "launch_gemv_add", "launch_gemm_f32", "launch_silu_mul", "launch_add_bias", "launch_attention_decode",
    "launch_attention_prefill", "launch_copy_to_kv_cache", "launch_add", "launch_add_inplace", "launch_copy",
    "launch_cosine_similarity",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "ntransformer_b200 has no CPU fallback.")
        _lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
        from . import _engine_sigs
        _engine_sigs.apply(_lib)
    return _lib
