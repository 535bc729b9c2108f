"""Llama-3 shapes and per-tensor quantisation mixes of the benchmark configs.

The Q4_K_M recipe follows llama.cpp ftype 15 as SURVEY §8d states it: everything Q4_K except attn_v and
ffn_down, which are Q6_K on layers with  i < L/8  or  i >= 7L/8  or  (i - L/8) % 3 == 2 ;  on the 70B shape the
remaining attn_v are Q5_K ; output.weight is Q6_K, token_embd Q4_K, norms F32."""
from __future__ import annotations

from dataclasses import dataclass, asdict

from .dtypes import DType, dtype_row_size


@dataclass
class LlamaConfig:
    vocab_size: int = 128256
    hidden_size: int = 4096
    intermediate_size: int = 14336
    n_layers: int = 32
    n_heads: int = 32
    n_kv_heads: int = 8
    head_dim: int = 128
    max_seq_len: int = 4096
    norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    bos_token_id: int = 128000
    eos_token_id: int = 128009

    def dict(self):
        return asdict(self)


LLAMA3_8B = LlamaConfig()
LLAMA3_70B = LlamaConfig(hidden_size=8192, intermediate_size=28672, n_layers=80, n_heads=64)
TINY = LlamaConfig(vocab_size=512, hidden_size=512, intermediate_size=1024, n_layers=3, n_heads=8, n_kv_heads=2, head_dim=64,
                   max_seq_len=128, bos_token_id=1, eos_token_id=2)

LAYER_TENSORS = ("attn_norm", "attn_q", "attn_k", "attn_v", "attn_output", "ffn_norm", "ffn_gate", "ffn_up", "ffn_down")


def use_more_bits(i: int, n: int) -> bool:
    return i < n // 8 or i >= 7 * n // 8 or (i - n // 8) % 3 == 2


def tensor_dtype(mix: str, name: str, layer: int, n_layers: int, big: bool) -> DType:
    """dtype of GGUF tensor `name` ('attn_q', 'output', 'token_embd', ...) under quantisation mix `mix`."""
    if name.endswith("norm"):
        return DType.F32
    mix = mix.upper()
    if mix == "Q4_K_M":
        if name == "output":
            return DType.Q6_K
        if name == "token_embd":
            return DType.Q4_K_M
        if name in ("attn_v", "ffn_down"):
            if use_more_bits(layer, n_layers):
                return DType.Q6_K
            return DType.Q5_K if (big and name == "attn_v") else DType.Q4_K_M
        return DType.Q4_K_M
    return {"Q8_0": DType.Q8_0, "Q4_0": DType.Q4_0, "Q6_K": DType.Q6_K, "Q5_K": DType.Q5_K, "Q4_K": DType.Q4_K_M,
            "F16": DType.F16, "F32": DType.F32}[mix]


def tensor_table(cfg: LlamaConfig, mix: str, tp_rank: int = 0, tp_size: int = 1):
    """[(gguf_name, dtype, rows, cols)] for this rank's shard, in file order."""
    big = cfg.n_layers >= 64
    hd, h = cfg.head_dim, cfg.hidden_size
    nh, nkv, inter = cfg.n_heads // tp_size, cfg.n_kv_heads // tp_size, cfg.intermediate_size // tp_size
    vl = -(-cfg.vocab_size // tp_size)
    vrows = cfg.vocab_size if tp_size == 1 else max(0, min(vl, cfg.vocab_size - tp_rank * vl))
    t = [("token_embd.weight", tensor_dtype(mix, "token_embd", 0, cfg.n_layers, big), cfg.vocab_size, h),
         ("output_norm.weight", DType.F32, 1, h),
         ("output.weight", tensor_dtype(mix, "output", 0, cfg.n_layers, big), vrows, h)]
    shapes = {"attn_norm": (1, h), "ffn_norm": (1, h), "attn_q": (nh * hd, h), "attn_k": (nkv * hd, h), "attn_v": (nkv * hd, h),
              "attn_output": (h, nh * hd), "ffn_gate": (inter, h), "ffn_up": (inter, h), "ffn_down": (h, inter)}
    for i in range(cfg.n_layers):
        for n in LAYER_TENSORS:
            r, c = shapes[n]
            t.append((f"blk.{i}.{n}.weight", tensor_dtype(mix, n, i, cfg.n_layers, big), r, c))
    return t


def bytes_per_token(cfg: LlamaConfig, mix: str, ctx: int, tp_size: int = 1) -> int:
    """Algorithmic bytes read per decoded token: 7 projections x L + output.weight + norms + KV."""
    b = 0
    for name, dt, r, c in tensor_table(cfg, mix, 0, tp_size):
        if name == "token_embd.weight":
            continue
        b += r * (c * 4 if dt == DType.F32 and r == 1 else dtype_row_size(dt, c))
    return b + 2 * cfg.n_layers * ctx * (cfg.n_kv_heads // tp_size) * cfg.head_dim * 2
