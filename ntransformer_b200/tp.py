"""Tensor-parallel shard plan for GGUF block-quantised weights (new relative to the single-GPU reference).

Row split (attn_q/k/v by heads, ffn_gate/up by FFN columns, output.weight by vocab rows) needs no communication;
column split (attn_output, ffn_down: slices of the *input* dimension at quantisation-block boundaries) produces
partial sums that one all-reduce per sub-block combines.  Norms and the embedding table are replicated."""
from __future__ import annotations

import numpy as np

from .dtypes import DType, dtype_block_size, dtype_row_size

ROW_SPLIT = ("attn_q", "attn_k", "attn_v", "ffn_gate", "ffn_up", "output")
COL_SPLIT = ("attn_output", "ffn_down")


def split_kind(name: str) -> str:
    base = name.replace(".weight", "").split(".")[-1]
    if base in ROW_SPLIT:
        return "rows"
    if base in COL_SPLIT:
        return "cols"
    return "replicate"


def shard_tensor(raw: np.ndarray, dtype: DType, rows: int, cols: int, name: str, rank: int, size: int):
    """Returns (bytes[rows_l, row_bytes_l], rows_l, cols_l) of `rank`'s shard of a [rows, cols] GGUF tensor."""
    raw = np.ascontiguousarray(raw).view(np.uint8).reshape(rows, -1)
    kind = split_kind(name)
    if size == 1 or kind == "replicate":
        return raw, rows, cols
    if kind == "rows":
        per = -(-rows // size)
        r0, r1 = min(rows, rank * per), min(rows, (rank + 1) * per)
        return raw[r0:r1], r1 - r0, cols
    bs = dtype_block_size(dtype)
    assert cols % size == 0 and (cols // size) % bs == 0, "column shard must be quantisation-block aligned"
    c_l = cols // size
    rb = dtype_row_size(dtype, c_l)
    return np.ascontiguousarray(raw[:, rank * rb:(rank + 1) * rb]), rows, c_l
