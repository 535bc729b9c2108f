"""Host-side mirror of the reference's kernel-launcher interface (src/cuda/kernels.h:10-74).

Same names, argument order and meaning as nt::cuda::launch_*; arguments are torch CUDA tensors (used
only as owners of device memory) or raw device pointers (ints).  Everything calls the C-ABI of
libnt_b200.so; like the reference the launches are asynchronous and return nothing.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import lib
from .dtypes import DType


def _p(t):
    if t is None:
        return None
    if isinstance(t, int):
        return t
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return t.data_ptr()


def _s(stream):
    if stream is None:
        return torch.cuda.current_stream().cuda_stream
    return stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)


def launch_rmsnorm(output, input, weight, batch_size, hidden_size, eps, stream=None):
    lib().nt_b200_rmsnorm(_p(output), _p(input), _p(weight), batch_size, hidden_size, eps, _s(stream))


def launch_rmsnorm_f16(output, input, weight, batch_size, hidden_size, eps, stream=None):
    lib().nt_b200_rmsnorm_f16(_p(output), _p(input), _p(weight), batch_size, hidden_size, eps, _s(stream))


def launch_rope(q, k, positions, batch_size, seq_len, n_heads, n_kv_heads, head_dim, theta_base, freq_scale,
                interleaved, stream=None):
    lib().nt_b200_rope(_p(q), _p(k), _p(positions), batch_size, seq_len, n_heads, n_kv_heads, head_dim, theta_base,
                       freq_scale, int(bool(interleaved)), _s(stream))


def launch_softmax(output, input, rows, cols, stream=None):
    lib().nt_b200_softmax(_p(output), _p(input), rows, cols, _s(stream))


def launch_masked_softmax(output, input, mask, rows, cols, stream=None):
    lib().nt_b200_masked_softmax(_p(output), _p(input), _p(mask), rows, cols, _s(stream))


def launch_gemv(y, W, x, out_features, in_features, weight_dtype, stream=None):
    lib().nt_b200_gemv(_p(y), _p(W), _p(x), out_features, in_features, int(weight_dtype), _s(stream))


def launch_gemv_add(y, W, x, out_features, in_features, weight_dtype, stream=None):
    lib().nt_b200_gemv_add(_p(y), _p(W), _p(x), out_features, in_features, int(weight_dtype), _s(stream))


def launch_gemm_f32(Cm, A, B, M, N, K, stream=None):
    lib().nt_b200_gemm_f32(_p(Cm), _p(A), _p(B), M, N, K, _s(stream))


def launch_silu_mul(output, gate, up, size, stream=None):
    lib().nt_b200_silu_mul(_p(output), _p(gate), _p(up), size, _s(stream))


def launch_add_bias(y, bias, size, stream=None):
    lib().nt_b200_add_bias(_p(y), _p(bias), size, _s(stream))


def launch_attention_decode(output, q, k_cache, v_cache, seq_len, n_heads, n_kv_heads, head_dim, max_seq, scale,
                            stream=None):
    lib().nt_b200_attention_decode(_p(output), _p(q), _p(k_cache), _p(v_cache), seq_len, n_heads, n_kv_heads,
                                   head_dim, max_seq, scale, _s(stream))


def launch_attention_prefill(output, Q, k_cache, v_cache, seq_len, start_pos, n_heads, n_kv_heads, head_dim,
                             max_seq, scale, stream=None):
    lib().nt_b200_attention_prefill(_p(output), _p(Q), _p(k_cache), _p(v_cache), seq_len, start_pos, n_heads,
                                    n_kv_heads, head_dim, max_seq, scale, _s(stream))


def launch_copy_to_kv_cache(k_cache, v_cache, k, v, seq_len, n_kv_heads, head_dim, start_pos, max_seq, stream=None):
    lib().nt_b200_copy_to_kv_cache(_p(k_cache), _p(v_cache), _p(k), _p(v), seq_len, n_kv_heads, head_dim, start_pos,
                                   max_seq, _s(stream))


def launch_add(out, a, b, size, stream=None):
    lib().nt_b200_add(_p(out), _p(a), _p(b), size, _s(stream))


def launch_add_inplace(a, b, size, stream=None):
    lib().nt_b200_add_inplace(_p(a), _p(b), size, _s(stream))


def launch_copy(dst, src, size, stream=None):
    lib().nt_b200_copy(_p(dst), _p(src), size, _s(stream))


def launch_cosine_similarity(result, a, b, size, stream=None):
    lib().nt_b200_cosine_similarity(_p(result), _p(a), _p(b), size, _s(stream))


# ---- additions without a reference counterpart ----
def xq_bytes(K: int) -> int:
    return lib().nt_b200_xq_bytes(K)


def quantize_x(x, xq, K, stream=None):
    lib().nt_b200_quantize_x(_p(x), _p(xq), K, _s(stream))


def gemv_fused(ys, Ws, outs, dtypes, in_features, xq, epilogue=0, stream=None):
    """Fused K-quant GEMV over up to 3 matrices sharing the quantised activations `xq`.
    epilogue: 0 store, 1 y += W.x, 2 ys[0] = silu(W0.x) * (W1.x)."""
    n = len(Ws)
    yp = (C.c_void_p * n)(*[_p(y) for y in ys])
    wp = (C.c_void_p * n)(*[_p(w) for w in Ws])
    op = (C.c_int * n)(*outs)
    dp = (C.c_int * n)(*[int(d) for d in dtypes])
    rc = lib().nt_b200_gemv_fused(n, yp, wp, op, dp, in_features, _p(xq), epilogue, _s(stream))
    if rc != 0:
        raise ValueError(f"nt_b200_gemv_fused rejected the launch (code {rc})")


def gemv_fused_f32(ys, Ws, outs, dtypes, in_features, x, epilogue=0, norm_w=None, eps=0.0, stream=None):
    """gemv_fused fed with the F32 vector: quantised in the kernel's prologue, optionally as RMSNorm(x) * norm_w
    (include/nt_b200.h nt_b200_gemv_fused_f32)."""
    n = len(Ws)
    yp = (C.c_void_p * n)(*[_p(y) for y in ys])
    wp = (C.c_void_p * n)(*[_p(w) for w in Ws])
    op = (C.c_int * n)(*outs)
    dp = (C.c_int * n)(*[int(d) for d in dtypes])
    rc = lib().nt_b200_gemv_fused_f32(n, yp, wp, op, dp, in_features, _p(x), _p(norm_w), eps, epilogue, _s(stream))
    if rc != 0:
        raise ValueError(f"nt_b200_gemv_fused_f32 rejected the launch (code {rc})")


def attention_decode_scratch_floats(max_seq, n_heads, n_kv_heads, head_dim) -> int:
    return int(lib().nt_b200_attention_decode_scratch_floats(max_seq, n_heads, n_kv_heads, head_dim))


def attention_decode_tickets(n_heads, n_kv_heads) -> int:
    return int(lib().nt_b200_attention_decode_tickets(n_heads, n_kv_heads))


def attention_decode_fused(out, q, k, v, k_cache, v_cache, pos_dev, max_seq, n_heads, n_kv_heads, head_dim, theta_base, freq_scale,
                           scale, scratch, tickets, xq_out=None, stream=None):
    """RoPE + KV-cache write at row *pos_dev + attention over *pos_dev + 1 keys + split merge (+ xq) in one launch."""
    lib().nt_b200_attention_decode_fused(_p(out), _p(q), _p(k), _p(v), _p(k_cache), _p(v_cache), _p(pos_dev), max_seq, n_heads,
                                         n_kv_heads, head_dim, theta_base, freq_scale, scale, _p(scratch), _p(tickets), _p(xq_out),
                                         _s(stream))


def embed_rows(out, table, dtype, tokens_dev, n_tokens, hidden, stream=None):
    lib().nt_b200_embed_rows(_p(out), _p(table), int(dtype), _p(tokens_dev), n_tokens, hidden, _s(stream))


def gemm_f16_tc_workspace_bytes(M: int, K: int) -> int:
    return lib().nt_b200_gemm_f16_tc_workspace_bytes(M, K)


def gemm_f16_tc(Cm, A, W_f16, M, N, K, workspace, stream=None):
    """Prefill GEMM on tcgen05 tensor cores: Cm[M,N] (F32) = A[M,K] (F32) . W_f16[N,K]^T."""
    rc = lib().nt_b200_gemm_f16_tc(_p(Cm), _p(A), _p(W_f16), M, N, K, _p(workspace), _s(stream))
    if rc != 0:
        raise ValueError(f"nt_b200_gemm_f16_tc rejected the shape M={M} N={N} K={K}")


def split_activations(workspace, A, M, K, stream=None):
    lib().nt_b200_split_activations(_p(workspace), _p(A), M, K, _s(stream))


def gemm_f16_tc_ws(Cm, workspace, W_f16, M, N, K, add=False, stream=None):
    """GEMM over an already split activation workspace; add=True accumulates into Cm (residual epilogue)."""
    rc = lib().nt_b200_gemm_f16_tc_ws(_p(Cm), _p(workspace), _p(W_f16), M, N, K, int(add), _s(stream))
    if rc != 0:
        raise ValueError(f"nt_b200_gemm_f16_tc_ws rejected the shape M={M} N={N} K={K}")


def gemm_f16_tc_swiglu_ws(workspace_out, workspace_in, Wgate_f16, Wup_f16, M, N, K, stream=None):
    """workspace_out = split(silu(A.Wgate^T) * (A.Wup^T)) for A pre-split in workspace_in (SwiGLU fused in the epilogue)."""
    rc = lib().nt_b200_gemm_f16_tc_swiglu_ws(_p(workspace_out), _p(workspace_in), _p(Wgate_f16), _p(Wup_f16), M, N, K, _s(stream))
    if rc != 0:
        raise ValueError(f"nt_b200_gemm_f16_tc_swiglu_ws rejected the shape M={M} N={N} K={K}")


def rmsnorm_split(workspace, x, w, rows, hidden, eps, stream=None):
    lib().nt_b200_rmsnorm_split(_p(workspace), _p(x), _p(w), rows, hidden, eps, _s(stream))


def dequant_split(w_hi, w_lo, W, dtype, rows, cols, row_pitch=0, stream=None):
    """GGUF blocks -> dense F16 pair with W = w_hi + w_lo (operands of the tensor-core prefill GEMM for quantised models)."""
    lib().nt_b200_dequant_split(_p(w_hi), _p(w_lo), _p(W), int(dtype), row_pitch, rows, cols, _s(stream))


def launch_count() -> int:
    return int(lib().nt_b200_launch_count())
