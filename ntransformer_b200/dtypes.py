"""DType ids and block geometry — ABI values of the reference's nt::DType (src/core/types.h:24-88)."""
from enum import IntEnum


class DType(IntEnum):
    F32 = 0
    F16 = 1
    Q8_0 = 2
    Q4_0 = 3
    Q4_K_M = 4
    Q6_K = 5
    Q5_K = 6
    Q2_K = 7
    I32 = 8


_SIZE = {DType.F32: 4, DType.F16: 2, DType.I32: 4, DType.Q8_0: 34, DType.Q4_0: 18, DType.Q4_K_M: 144,
         DType.Q5_K: 176, DType.Q6_K: 210, DType.Q2_K: 84}
_BLOCK = {DType.Q8_0: 32, DType.Q4_0: 32, DType.Q4_K_M: 256, DType.Q5_K: 256, DType.Q6_K: 256, DType.Q2_K: 256}

# GGML tensor-type id -> DType (src/core/types.h:168-217)
GGML_TO_DTYPE = {0: DType.F32, 1: DType.F16, 8: DType.Q8_0, 2: DType.Q4_0, 12: DType.Q4_K_M, 13: DType.Q5_K,
                 14: DType.Q6_K, 10: DType.Q2_K, 26: DType.I32}
DTYPE_TO_GGML = {v: k for k, v in GGML_TO_DTYPE.items()}


def dtype_size(dt) -> int:
    """Bytes per element (plain types) or per quantisation block."""
    return _SIZE.get(DType(dt), 0)


def dtype_block_size(dt) -> int:
    return _BLOCK.get(DType(dt), 1)


def dtype_row_size(dt, n: int) -> int:
    bs = dtype_block_size(dt)
    assert n % bs == 0, "row length must be a multiple of the block size"
    return n // bs * dtype_size(dt)
