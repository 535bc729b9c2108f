"""Synthetic GGUF-layout weights: random *valid* quantisation blocks (numpy on the host, torch on the GPU).

Any byte pattern with finite fp16 scales is a valid block of every GGUF format used here, so blocks are
drawn directly: codes uniform, 6-bit scales/mins uniform, fp16 super-block scales chosen so
the dequantised weights have std ~ `std` and mean ~ 0.
"""
from __future__ import annotations

import numpy as np

from .dtypes import DType, dtype_row_size, dtype_size

# (d, dmin) magnitudes that give dequantised std ~= 1 for uniformly random codes, per format
_UNIT = {
    DType.Q4_K_M: (1.0 / 160.0, 7.5 / 160.0),   # w = d*sc*q - dmin*m, sc,m~U[0,63], q~U[0,15]
    DType.Q5_K: (1.0 / 330.0, 15.5 / 330.0),
    DType.Q6_K: (1.0 / 170.0, 0.0),             # scales drawn in [-16,15]
    DType.Q8_0: (1.0 / 73.6, 0.0),
    DType.Q4_0: (1.0 / 4.6, 0.0),
}


def random_blocks_np(dtype: DType, rows: int, cols: int, rng: np.random.Generator, std: float = 0.02) -> np.ndarray:
    """uint8 array [rows, row_bytes] of random valid blocks (or f16/f32 data viewed as bytes)."""
    dtype = DType(dtype)
    if dtype == DType.F32:
        return (rng.standard_normal((rows, cols), dtype=np.float32) * std).view(np.uint8).reshape(rows, -1)
    if dtype == DType.F16:
        return (rng.standard_normal((rows, cols), dtype=np.float32) * std).astype(np.float16).view(np.uint8).reshape(rows, -1)
    bs = dtype_size(dtype)
    nb = dtype_row_size(dtype, cols) // bs
    blk = rng.integers(0, 256, size=(rows, nb, bs), dtype=np.uint8)
    du, dminu = _UNIT[dtype]
    jitter = rng.uniform(0.5, 1.5, size=(rows, nb)).astype(np.float32)
    d = (jitter * du * std).astype(np.float16).view(np.uint16)
    if dtype in (DType.Q4_K_M, DType.Q5_K):
        dm = (jitter * dminu * std).astype(np.float16).view(np.uint16)
        blk[:, :, 0] = d & 0xFF
        blk[:, :, 1] = d >> 8
        blk[:, :, 2] = dm & 0xFF
        blk[:, :, 3] = dm >> 8
    elif dtype == DType.Q6_K:
        sc = rng.integers(-16, 16, size=(rows, nb, 16), dtype=np.int8)
        blk[:, :, 192:208] = sc.view(np.uint8)
        blk[:, :, 208] = d & 0xFF
        blk[:, :, 209] = d >> 8
    else:  # Q8_0 / Q4_0: d first
        blk[:, :, 0] = d & 0xFF
        blk[:, :, 1] = d >> 8
    return blk.reshape(rows, nb * bs)


def random_blocks_cuda(dtype: DType, rows: int, cols: int, seed: int, std: float = 0.02, device="cuda"):
    """Same distribution generated on the GPU with torch (for multi-GB synthetic models)."""
    import torch

    dtype = DType(dtype)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if dtype == DType.F32:
        return (torch.randn((rows, cols), generator=g, device=device, dtype=torch.float32) * std).view(torch.uint8)
    if dtype == DType.F16:
        return (torch.randn((rows, cols), generator=g, device=device, dtype=torch.float32) * std).to(torch.float16).view(torch.uint8)
    bs = dtype_size(dtype)
    nb = dtype_row_size(dtype, cols) // bs
    blk = torch.randint(0, 256, (rows, nb, bs), generator=g, device=device, dtype=torch.uint8)
    du, dminu = _UNIT[dtype]
    jitter = torch.rand((rows, nb), generator=g, device=device, dtype=torch.float32) + 0.5
    d = (jitter * (du * std)).to(torch.float16).view(torch.int16).to(torch.int32) & 0xFFFF
    lo, hi = (d & 0xFF).to(torch.uint8), (d >> 8).to(torch.uint8)
    if dtype in (DType.Q4_K_M, DType.Q5_K):
        dm = (jitter * (dminu * std)).to(torch.float16).view(torch.int16).to(torch.int32) & 0xFFFF
        blk[:, :, 0], blk[:, :, 1] = lo, hi
        blk[:, :, 2], blk[:, :, 3] = (dm & 0xFF).to(torch.uint8), (dm >> 8).to(torch.uint8)
    elif dtype == DType.Q6_K:
        sc = torch.randint(-16, 16, (rows, nb, 16), generator=g, device=device, dtype=torch.int8)
        blk[:, :, 192:208] = sc.view(torch.uint8)
        blk[:, :, 208], blk[:, :, 209] = lo, hi
    else:
        blk[:, :, 0], blk[:, :, 1] = lo, hi
    return blk.view(rows, nb * bs)
