"""CPU tests of the host logic behind the C-ABI: GGUF parser, tokenizer, sampler, model specs."""
import ctypes as C
import json

import numpy as np
import pytest

from ntransformer_b200._lib import lib
from ntransformer_b200.dtypes import DType
from ntransformer_b200.gguf_write import gpt2_byte_tokens, synthetic_tensors_np, write_gguf
from ntransformer_b200.model_spec import LLAMA3_8B, LLAMA3_70B, TINY, bytes_per_token, tensor_dtype, tensor_table


@pytest.fixture(scope="module")
def tiny_gguf(tmp_path_factory):
    p = tmp_path_factory.mktemp("gguf") / "tiny.gguf"
    tensors = synthetic_tensors_np(TINY, "Q4_K_M", seed=5)
    vocab = gpt2_byte_tokens() + ["He", "llo", "Hello", "Ġwor", "ld", "Ġworld"] + [f"<t{i}>" for i in range(262, TINY.vocab_size)]
    write_gguf(p, TINY, tensors, vocab_tokens=vocab)
    return p, tensors, vocab


def describe(path):
    buf = C.create_string_buffer(4 << 20)
    n = lib().nt_gguf_describe(str(path).encode(), buf, len(buf))
    assert n > 0
    return json.loads(buf.value.decode())


def test_gguf_parser_matches_writer_and_gguf_package(tiny_gguf):
    path, tensors, vocab = tiny_gguf
    d = describe(path)
    for k in ("vocab_size", "hidden_size", "intermediate_size", "n_layers", "n_heads", "n_kv_heads", "head_dim", "max_seq_len",
              "bos_token_id", "eos_token_id"):
        assert d[k] == getattr(TINY, k), k
    assert abs(d["rope_theta"] - TINY.rope_theta) < 1e-3 and abs(d["norm_eps"] - TINY.norm_eps) < 1e-9
    assert d["n_vocab_tokens"] == len(vocab) and d["data_offset"] % 32 == 0
    by_name = {t["name"]: t for t in d["tensors"]}
    assert set(by_name) == set(tensors)
    for name, (arr, dt, rows, cols) in tensors.items():
        t = by_name[name]
        assert t["dtype"] == int(dt) and t["nbytes"] == np.asarray(arr).nbytes and t["offset"] % 32 == 0
    gguf = pytest.importorskip("gguf")
    r = gguf.GGUFReader(str(path))                      # independent parser agrees on offsets and bytes
    raw_file = np.fromfile(path, dtype=np.uint8)
    for t in r.tensors:
        ours = by_name[t.name]
        assert int(t.data_offset) == d["data_offset"] + ours["offset"]
        want = np.ascontiguousarray(tensors[t.name][0]).view(np.uint8).ravel()
        got = raw_file[d["data_offset"] + ours["offset"]: d["data_offset"] + ours["offset"] + ours["nbytes"]]
        np.testing.assert_array_equal(got, want)


def test_gguf_rejects_garbage(tmp_path):
    bad = tmp_path / "bad.gguf"
    bad.write_bytes(b"NOPE" + b"\0" * 64)
    buf = C.create_string_buffer(1024)
    assert lib().nt_gguf_describe(str(bad).encode(), buf, 1024) == -1
    trunc = tmp_path / "trunc.gguf"
    trunc.write_bytes(b"GGUF" + (3).to_bytes(4, "little") + (10).to_bytes(8, "little") + (5).to_bytes(8, "little"))
    assert lib().nt_gguf_describe(str(trunc).encode(), buf, 1024) == -1
    assert lib().nt_gguf_describe(str(tmp_path / "missing.gguf").encode(), buf, 1024) == -1


def tok(path, text, bos=True):
    ids = (C.c_int * 256)()
    n = lib().nt_tokenize(str(path).encode(), text.encode(), int(bos), ids, 256)
    return list(ids[:n])


def detok(path, ids):
    arr = (C.c_int * len(ids))(*ids)
    out = C.create_string_buffer(4096)
    n = lib().nt_detokenize(str(path).encode(), arr, len(ids), out, 4096)
    assert n >= 0
    return out.raw[:n]


def test_tokenizer_longest_match_merges_and_roundtrip(tiny_gguf):
    path, _, vocab = tiny_gguf
    ids = tok(path, "Hello world")
    assert ids[0] == TINY.bos_token_id
    assert ids[1:] == [vocab.index("Hello"), vocab.index("Ġworld")]     # greedy longest match wins
    assert tok(path, "", bos=True) == [TINY.bos_token_id] and tok(path, "", bos=False) == []
    for text in ["Hello world", "a\tb\nc", "café ☃ \U0001F600", "x" * 100]:
        ids = tok(path, text, bos=False)
        assert detok(path, ids) == text.encode()                               # byte-level fallback is lossless


def test_sampler_greedy_penalty_and_seeded_sampling():
    L = lib()
    logits = np.array([0.1, 2.0, 1.9, -1.0, 0.5], dtype=np.float32)
    p = logits.ctypes.data_as(C.c_void_p)
    assert L.nt_sample_token(p, 5, 0.0, 40, 0.9, 1.0, 64, None, 0, 42) == 1
    recent = (C.c_int * 2)(1, 3)                       # penalising token 1 (2.0/1.1 < 1.9) flips the argmax (sampler.cpp:30-45)
    assert L.nt_sample_token(p, 5, 0.0, 40, 0.9, 1.1, 64, recent, 2, 42) == 2
    draws = [L.nt_sample_token(p, 5, 0.8, 3, 0.95, 1.0, 64, None, 0, s) for s in range(200)]
    assert set(draws) <= {1, 2, 4} and len(set(draws)) >= 2                     # top-k = 3 restricts the support
    assert draws == [L.nt_sample_token(p, 5, 0.8, 3, 0.95, 1.0, 64, None, 0, s) for s in range(200)]   # seeded => reproducible
    assert all(L.nt_sample_token(p, 5, 0.8, 40, 1e-6, 1.0, 64, None, 0, s) == 1 for s in range(20))  # tiny top-p => argmax


def test_q4_k_m_mix_and_byte_counts_match_survey():
    assert tensor_dtype("Q4_K_M", "attn_v", 0, 80, True) == DType.Q6_K
    assert tensor_dtype("Q4_K_M", "attn_v", 11, 80, True) == DType.Q5_K
    assert tensor_dtype("Q4_K_M", "ffn_down", 11, 80, True) == DType.Q4_K_M
    assert tensor_dtype("Q4_K_M", "ffn_down", 12, 80, True) == DType.Q6_K
    assert tensor_dtype("Q4_K_M", "output", 0, 32, False) == DType.Q6_K
    # SURVEY §8d expected values (GB)
    for cfg, mix, ctx, want in [(LLAMA3_8B, "Q8_0", 0, 7.974), (LLAMA3_8B, "Q4_K_M", 0, 4.616), (LLAMA3_70B, "Q4_K_M", 0, 41.92),
                                (LLAMA3_70B, "Q6_K", 0, 57.01), (LLAMA3_8B, "F16", 0, 15.01)]:
        assert abs(bytes_per_token(cfg, mix, ctx) / 1e9 - want) < 0.01
    assert abs(bytes_per_token(LLAMA3_70B, "Q6_K", 0, tp_size=8) / 1e9 - 7.13) < 0.01
    # shards tile the full model
    full = {n: (r, c) for n, _, r, c in tensor_table(LLAMA3_70B, "Q4_K_M")}
    for tp in (2, 4, 8):
        rows = {}
        for rank in range(tp):
            for n, _, r, c in tensor_table(LLAMA3_70B, "Q4_K_M", rank, tp):
                rows.setdefault(n, []).append((r, c))
        for n, parts in rows.items():
            R, Cc = full[n]
            if "attn_output" in n or "ffn_down" in n:
                assert sum(c for _, c in parts) == Cc and all(r == R for r, _ in parts)
            elif n.endswith("norm.weight") or n == "token_embd.weight":
                assert all(p == (R, Cc) for p in parts)
            else:
                assert sum(r for r, _ in parts) == R and all(c == Cc for _, c in parts)


def _valid_gguf(tmp_path):
    from ntransformer_b200.gguf_write import synthetic_tensors_np, write_gguf
    from ntransformer_b200.model_spec import TINY
    p = tmp_path / "ok.gguf"
    write_gguf(p, TINY, synthetic_tensors_np(TINY, "Q4_K_M", seed=2))
    return p


def _describe(path):
    import ctypes as C
    from ntransformer_b200._lib import lib
    buf = C.create_string_buffer(4 << 20)
    return lib().nt_gguf_describe(str(path).encode(), buf, len(buf))


def test_gguf_parser_rejects_malformed_files_instead_of_trusting_them(tmp_path, capfd):
    """counts, shapes and offsets read from the file are checked against the bytes that are there (the reference's
    loader.cpp:23-276 trusts them).  Every mutation must make nt_gguf_describe fail cleanly — no crash, no exception escaping the
    C-ABI — while the untouched file parses."""
    import struct
    ok = _valid_gguf(tmp_path)
    assert _describe(ok) > 0
    raw = bytearray(ok.read_bytes())

    def variant(name, edit):
        b = bytearray(raw)
        edit(b)
        p = tmp_path / name
        p.write_bytes(bytes(b))
        return p

    # header: magic, version, absurd tensor / key counts
    assert _describe(variant("magic.gguf", lambda b: b.__setitem__(slice(0, 4), b"GGUX"))) <= 0
    assert _describe(variant("version.gguf", lambda b: b.__setitem__(slice(4, 8), struct.pack("<I", 9)))) <= 0
    assert _describe(variant("ntensors.gguf", lambda b: b.__setitem__(slice(8, 16), struct.pack("<Q", 1 << 60)))) <= 0
    assert _describe(variant("nkv.gguf", lambda b: b.__setitem__(slice(16, 24), struct.pack("<Q", 1 << 61)))) <= 0
    # truncation inside the metadata and inside the tensor data
    assert _describe(variant("trunc_meta.gguf", lambda b: b.__delitem__(slice(200, len(b))))) <= 0
    assert _describe(variant("trunc_data.gguf", lambda b: b.__delitem__(slice(len(b) - 4096, len(b))))) <= 0
    # a string length that runs past the end of the file (first key's length field sits right after the 24-byte header)
    assert _describe(variant("strlen.gguf", lambda b: b.__setitem__(slice(24, 32), struct.pack("<Q", 1 << 40)))) <= 0
    # tensor table: find the first tensor record and corrupt its dimension count, a dimension, and its offset
    name = b"token_embd.weight"
    at = raw.find(struct.pack("<Q", len(name)) + name)
    assert at > 0
    nd_at = at + 8 + len(name)
    assert _describe(variant("nd0.gguf", lambda b: b.__setitem__(slice(nd_at, nd_at + 4), struct.pack("<I", 0)))) <= 0
    assert _describe(variant("nd9.gguf", lambda b: b.__setitem__(slice(nd_at, nd_at + 4), struct.pack("<I", 9)))) <= 0
    assert _describe(variant("dim0.gguf", lambda b: b.__setitem__(slice(nd_at + 4, nd_at + 12), struct.pack("<Q", 0)))) <= 0
    assert _describe(variant("dimhuge.gguf", lambda b: b.__setitem__(slice(nd_at + 4, nd_at + 12), struct.pack("<Q", 1 << 62)))) <= 0
    assert _describe(variant("dimodd.gguf", lambda b: b.__setitem__(slice(nd_at + 4, nd_at + 12), struct.pack("<Q", 100)))) <= 0   # not a multiple of the 256-weight block
    off_at = nd_at + 4 + 16 + 4
    assert _describe(variant("offset.gguf", lambda b: b.__setitem__(slice(off_at, off_at + 8), struct.pack("<Q", (1 << 64) - 64)))) <= 0
    capfd.readouterr()
