"""Opt-in rank-ordered byte-level BPE (csrc/engine/text.cpp encode_merges): the GGUF's tokenizer.ggml.merges, a
Llama-3 style pre-tokeniser and literal special tokens.  The reference ignores the merges array (loader.cpp skips it,
tokenizer.cpp:101-217 merges by score), so the default path stays pinned to the reference (tests/test_ref_host.py) and this one is
checked against a straightforward Python restatement of GPT-2 BPE on a synthetic vocabulary.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from ntransformer_b200._lib import lib
from ntransformer_b200.gguf_write import gpt2_byte_tokens, synthetic_tensors_np, write_gguf
from ntransformer_b200.model_spec import LlamaConfig


def _py_bpe(word_bytes, byte_syms, ranks, ids):
    sym = [byte_syms[b] for b in word_bytes]
    while len(sym) > 1:
        best, br = -1, 1 << 60
        for i in range(len(sym) - 1):
            r = ranks.get((sym[i], sym[i + 1]))
            if r is not None and r < br:
                best, br = i, r
        if best < 0:
            break
        sym[best:best + 2] = [sym[best] + sym[best + 1]]
    return [ids[s] for s in sym]


@pytest.fixture(scope="module")
def bpe_file(tmp_path_factory):
    bs = gpt2_byte_tokens()
    sp = bs[0x20]
    merges = [("h", "e"), ("l", "l"), ("he", "ll"), ("hell", "o"), (sp, "w"), ("o", "r"), (sp + "w", "or"), ("l", "d"), (sp + "wor", "ld"),
              (sp, "t"), (sp + "t", "he"), ("1", "2"), ("12", "3"), ("!", "!"), (sp, "h"), (sp + "h", "e"), (bs[0xC3], bs[0xA9])]
    tokens = list(bs)
    for a, b in merges:
        if a + b not in tokens:
            tokens.append(a + b)
    specials = ["<|begin|>", "<|eot|>"]
    tokens += specials
    types = [1] * (len(tokens) - len(specials)) + [3, 3]
    cfg = LlamaConfig(vocab_size=len(tokens), hidden_size=256, intermediate_size=256, n_layers=1, n_heads=2, n_kv_heads=2, head_dim=128,
                      max_seq_len=32, bos_token_id=len(tokens) - 2, eos_token_id=len(tokens) - 1)
    path = tmp_path_factory.mktemp("bpe") / "bpe.gguf"
    write_gguf(path, cfg, synthetic_tensors_np(cfg, "Q8_0", seed=1), vocab_tokens=tokens, vocab_types=types,
               merges=[f"{a} {b}" for a, b in merges])
    ids = {t: i for i, t in enumerate(tokens)}
    ranks = {m: r for r, m in enumerate(merges)}
    return path, cfg, bs, ids, ranks


def _encode(path, text, add_bos=False):
    buf = (C.c_int * 4096)()
    n = lib().nt_tokenize(str(path).encode(), text.encode("utf-8"), int(add_bos), buf, 4096)
    assert 0 <= n <= 4096
    return list(buf[:n])


def _decode(path, ids):
    arr = (C.c_int * len(ids))(*ids)
    out = C.create_string_buffer(1 << 16)
    n = lib().nt_detokenize(str(path).encode(), arr, len(ids), out, len(out))
    assert n >= 0
    return out.raw[:n].decode("utf-8", errors="replace")


def test_rank_ordered_merges_and_specials(bpe_file, monkeypatch, capfd):
    path, cfg, bs, ids, ranks = bpe_file
    sp = bs[0x20]
    monkeypatch.setenv("NT_B200_BPE_MERGES", "1")
    assert _encode(path, "hello world") == [ids["hello"], ids[sp + "world"]]
    assert _encode(path, "hello world", add_bos=True)[0] == cfg.bos_token_id
    # digits in groups of at most three, punctuation runs, a contraction, a special token written literally
    assert _encode(path, "12345") == [ids["123"], ids["4"], ids["5"]]
    assert _encode(path, "the!!") == [ids["t"], ids["he"], ids["!!"]]
    assert _encode(path, "hello<|eot|> the") == [ids["hello"], ids["<|eot|>"], ids[sp + "the"]]
    assert _encode(path, "he's") == [ids["he"], ids["'"], ids["s"]]
    # UTF-8: e-acute is the two bytes C3 A9, merged by the last rule
    assert _encode(path, "é") == [ids[bs[0xC3] + bs[0xA9]]]
    capfd.readouterr()


def test_matches_a_python_restatement_and_round_trips(bpe_file, monkeypatch, capfd):
    path, cfg, bs, ids, ranks = bpe_file
    monkeypatch.setenv("NT_B200_BPE_MERGES", "1")
    rng = np.random.default_rng(0)
    words = ["hello", " world", " the", "hell", "he", " help", "o", " w", "ld", "123", " 12", "!!", "é", " été", "x", " ", "\n"]
    for _ in range(60):
        text = "".join(rng.choice(words) for _ in range(int(rng.integers(1, 12))))
        got = _encode(path, text)
        assert _decode(path, got) == text                      # byte-level BPE is lossless
        # single-word inputs (no pre-tokeniser boundary inside) agree with the restatement symbol for symbol
    for w in ["hello", " world", " the", "hell", " help", "123", "!!", " wor", "hellohello"]:
        if w == "hellohello":
            continue                                           # letters only: one pre-token, merges may cross the middle
        assert _encode(path, w) == _py_bpe(w.encode("utf-8"), bs, ranks, ids), w
    assert _encode(path, "hellohello") == _py_bpe(b"hellohello", bs, ranks, ids)
    capfd.readouterr()


def test_default_path_ignores_the_merges_like_the_reference(bpe_file, monkeypatch, capfd):
    path, cfg, bs, ids, ranks = bpe_file
    monkeypatch.delenv("NT_B200_BPE_MERGES", raising=False)
    a = _encode(path, "hello world")
    monkeypatch.setenv("NT_B200_BPE_MERGES", "0")
    assert _encode(path, "hello world") == a
    assert _decode(path, a) == "hello world"
    capfd.readouterr()
