"""Minimal GGUF v3 writer for synthetic models (tests, CLI runs, reference-side parity probes).

Writes exactly the keys the reference reads (src/model/config.cpp:30-49, src/model/loader.cpp:88-141) and the
tensor names it looks up (src/model/transformer.cpp:89-104, 286-322); data section aligned to 32 bytes."""
from __future__ import annotations

import struct

import numpy as np

from .dtypes import DTYPE_TO_GGML, DType
from .model_spec import LlamaConfig, tensor_table
from .synth import random_blocks_np

_T_U32, _T_F32, _T_STR, _T_ARR, _T_I32 = 4, 6, 8, 9, 5


def _s(b: str) -> bytes:
    e = b.encode("utf-8")
    return struct.pack("<Q", len(e)) + e


def gpt2_byte_tokens():
    """The 256 single-'byte' tokens of the GPT-2 byte<->unicode map (so the tokenizer detects GPT2-BPE)."""
    keep = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    out, extra = [], 0
    for b in range(256):
        if b in keep:
            out.append(chr(b))
        else:
            out.append(chr(256 + extra))
            extra += 1
    return out


def write_gguf(path, cfg: LlamaConfig, tensors: dict, vocab_tokens=None, name="synthetic", vocab_scores=None, vocab_types=None,
               merges=None):
    """tensors: name -> (np.ndarray bytes/f32, DType, rows, cols) in GGUF block layout."""
    table = [(n, dt, r, c) for n, (a, dt, r, c) in tensors.items()]
    write_gguf_streaming(path, cfg, ((n, a, dt, r, c) for n, (a, dt, r, c) in tensors.items()), table, vocab_tokens, name,
                         vocab_scores, vocab_types, merges)


def _tensor_nbytes(dt, rows, cols):
    from .dtypes import dtype_row_size
    return rows * dtype_row_size(dt, cols)


def write_gguf_streaming(path, cfg: LlamaConfig, gen, table=None, vocab_tokens=None, name="synthetic", vocab_scores=None,
                         vocab_types=None, merges=None):
    """Like write_gguf but `gen` yields (name, array, DType, rows, cols) one tensor at a time, in the order of
    `table` ([(name, DType, rows, cols)], default tensor_table(cfg-independent order of gen is NOT allowed))."""
    if table is None:
        table = tensor_table(cfg, "F32")      # only valid when the generator follows the standard order and dtypes
        raise ValueError("write_gguf_streaming needs the (name, dtype, rows, cols) table up front")
    if vocab_tokens is None:
        vocab_tokens = gpt2_byte_tokens() + [f"<t{i}>" for i in range(256, cfg.vocab_size)]
        vocab_tokens = vocab_tokens[: cfg.vocab_size]
    kv = []

    def kv_u32(k, v): kv.append(_s(k) + struct.pack("<II", _T_U32, v))
    def kv_f32(k, v): kv.append(_s(k) + struct.pack("<If", _T_F32, v))
    def kv_str(k, v): kv.append(_s(k) + struct.pack("<I", _T_STR) + _s(v))

    kv_str("general.architecture", "llama")
    kv_str("general.name", name)
    kv_u32("general.alignment", 32)
    kv_u32("llama.embedding_length", cfg.hidden_size)
    kv_u32("llama.feed_forward_length", cfg.intermediate_size)
    kv_u32("llama.block_count", cfg.n_layers)
    kv_u32("llama.attention.head_count", cfg.n_heads)
    kv_u32("llama.attention.head_count_kv", cfg.n_kv_heads)
    kv_u32("llama.context_length", cfg.max_seq_len)
    kv_f32("llama.attention.layer_norm_rms_epsilon", cfg.norm_eps)
    kv_f32("llama.rope.freq_base", cfg.rope_theta)
    kv_u32("tokenizer.ggml.bos_token_id", cfg.bos_token_id)
    kv_u32("tokenizer.ggml.eos_token_id", cfg.eos_token_id)
    kv.append(_s("tokenizer.ggml.tokens") + struct.pack("<IIQ", _T_ARR, _T_STR, len(vocab_tokens)) + b"".join(_s(t) for t in vocab_tokens))
    scores = (np.arange(len(vocab_tokens), 0, -1, dtype=np.float32) if vocab_scores is None
              else np.asarray(vocab_scores, dtype=np.float32))
    assert len(scores) == len(vocab_tokens)
    kv.append(_s("tokenizer.ggml.scores") + struct.pack("<IIQ", _T_ARR, _T_F32, len(scores)) + scores.tobytes())
    types = np.ones(len(vocab_tokens), dtype=np.int32) if vocab_types is None else np.asarray(vocab_types, dtype=np.int32)
    assert len(types) == len(vocab_tokens)
    kv.append(_s("tokenizer.ggml.token_type") + struct.pack("<IIQ", _T_ARR, _T_I32, len(types)) + types.tobytes())
    if merges:              # "left right" in rank order, as llama.cpp writes them (the reference skips this array)
        kv.append(_s("tokenizer.ggml.merges") + struct.pack("<IIQ", _T_ARR, _T_STR, len(merges)) + b"".join(_s(m) for m in merges))

    infos, offset, offsets = [], 0, {}
    for tname, dt, rows, cols in table:
        nbytes = _tensor_nbytes(dt, rows, cols)
        dims = (cols,) if rows == 1 and tname.endswith("norm.weight") else (cols, rows)
        infos.append(_s(tname) + struct.pack("<I", len(dims)) + b"".join(struct.pack("<Q", d) for d in dims)
                     + struct.pack("<IQ", DTYPE_TO_GGML[DType(dt)], offset))
        offsets[tname] = (offset, nbytes)
        offset = (offset + nbytes + 31) & ~31
    header = struct.pack("<IIQQ", 0x46554747, 3, len(infos), len(kv)) + b"".join(kv) + b"".join(infos)
    pad = (-len(header)) % 32
    with open(path, "wb") as f:
        f.write(header + b"\0" * pad)
        base = f.tell()
        for tname, arr, dt, rows, cols in gen:
            raw = np.ascontiguousarray(arr).view(np.uint8).ravel()
            off, nbytes = offsets[tname]
            assert raw.size == nbytes, (tname, raw.size, nbytes)
            f.seek(base + off)
            raw.tofile(f)
        f.truncate(base + offset)


def synthetic_tensors_np(cfg: LlamaConfig, mix: str, seed: int = 1234) -> dict:
    """name -> (array, DType, rows, cols): random valid blocks for every tensor the engine needs."""
    out = {}
    for idx, (name, dt, rows, cols) in enumerate(tensor_table(cfg, mix)):
        rng = np.random.default_rng(seed + idx)
        if name.endswith("norm.weight"):
            arr = (1.0 + 0.1 * rng.standard_normal(cols)).astype(np.float32)
        else:
            arr = random_blocks_np(dt, rows, cols, rng)
        out[name] = (arr, dt, rows, cols)
    return out
