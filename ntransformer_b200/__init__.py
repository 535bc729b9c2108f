"""ntransformer_b200 — B200-native (sm_100a) resident quantized-decode path, drop-in behind the
kernel-launcher interface of xaskasdf/ntransformer (reference src/cuda/kernels.h).

The product is the CUDA shared library `libnt_b200.so` (kernels + C-ABI + native C++ engine).  This
package is only the Python host-side mirror of that interface (ctypes; torch is used for device memory,
streams and torch.distributed plumbing).  There is no CPU fallback: importing the kernel bindings without
the built library raises.
"""
from .dtypes import DType, dtype_size, dtype_block_size, dtype_row_size  # noqa: F401

__all__ = ["DType", "dtype_size", "dtype_block_size", "dtype_row_size"]
__version__ = "0.1.0"
