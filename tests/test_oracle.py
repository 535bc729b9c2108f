"""CPU tests: pin the oracle (oracle/nt_oracle.c) against
  (a) the reference's own known-answer tests (tests/test_gemm.cpp, tests/test_tensor.cpp),
  (b) golden fixtures produced by importing the reference's numpy dequantisers (tests/golden/),
  (c) the independent `gguf` package, and (d) internal consistency (gemv == dequant . x)."""
from pathlib import Path

import numpy as np
import pytest

from ntransformer_b200.dtypes import DType, dtype_row_size, dtype_size
from ntransformer_b200.synth import random_blocks_np
from oracle import oracle as O

GOLD = Path(__file__).resolve().parent / "golden"
ALL_Q = [DType.Q8_0, DType.Q4_0, DType.Q4_K_M, DType.Q5_K, DType.Q6_K, DType.F16, DType.F32]


def f16bits(v):
    return int(np.array([v], dtype=np.float16).view(np.uint16)[0])


# ---------- (a) reference known-answer tests ----------
def test_kat_dtype_sizes():
    # tests/test_tensor.cpp:125-135 (its Q8_0==36 assert is stale; types.h:43 returns 34)
    assert O.lib().nto_dtype_size(O.F32) == 4 and O.lib().nto_dtype_size(O.F16) == 2
    assert O.lib().nto_dtype_size(O.Q4_0) == 18 and O.lib().nto_dtype_size(O.Q8_0) == 34
    assert O.row_bytes(O.Q4_0, 1024) == 576
    for dt in ALL_Q:
        assert O.row_bytes(int(dt), 1024) == dtype_row_size(dt, 1024)
        assert O.lib().nto_dtype_size(int(dt)) == dtype_size(dt)


def test_kat_gemv_f32():
    # tests/test_gemm.cpp:26-61: W = [[1..3],[4..6],[7..9],[10..12]], x = ones -> {6,15,24,33}
    W = np.arange(1, 13, dtype=np.float32).reshape(4, 3)
    y = O.gemv(W, np.ones(3, np.float32), 4, 3, O.F32)
    np.testing.assert_allclose(y, [6, 15, 24, 33], atol=1e-3)


def q4_0_const_block(d, nib):
    b = np.zeros(18, np.uint8)
    b[0:2] = np.array([f16bits(d)], np.uint16).view(np.uint8)
    b[2:] = (nib << 4) | nib
    return b


def test_kat_gemv_q4_0():
    # tests/test_gemm.cpp:76-158: d=0.5, nibbles 10 -> +1.0 ; nibbles 7 -> -0.5 ; x = ones -> {32,-16}
    W = np.stack([q4_0_const_block(0.5, 10), q4_0_const_block(0.5, 7)])
    y = O.gemv(W, np.ones(32, np.float32), 2, 32, O.Q4_0)
    np.testing.assert_allclose(y, [32.0, -16.0], atol=0.1)


def q6_k_const_row(nblocks, ql, qh):
    blk = np.zeros(210, np.uint8)
    blk[0:128] = ql
    blk[128:192] = qh
    blk[192:208] = 1
    blk[208:210] = np.array([f16bits(1.0)], np.uint16).view(np.uint8)
    return np.tile(blk, nblocks)


@pytest.mark.parametrize("K", [256, 32768])
def test_kat_gemv_q6_k(K):
    # tests/test_gemm.cpp:266-393: (ql 0x11, qh 0xAA) -> +1, (ql 0xFF, qh 0x55) -> -1 ; x = ones -> +-K
    nb = K // 256
    W = np.stack([q6_k_const_row(nb, 0x11, 0xAA), q6_k_const_row(nb, 0xFF, 0x55)])
    y = O.gemv(W, np.ones(K, np.float32), 2, K, O.Q6_K)
    np.testing.assert_allclose(y, [K, -K], atol=0.5 if K == 256 else 1.0)


def test_kat_silu_mul():
    # tests/test_gemm.cpp:172-199
    out = O.silu_mul(np.array([0, 1, -1, 2], np.float32), np.ones(4, np.float32))
    np.testing.assert_allclose(out, [0.0, 0.731, -0.269, 1.762], atol=0.01)


def test_kat_rmsnorm():
    # tests/test_gemm.cpp:212-241: [1,2,3,4], w=1, eps 1e-5 -> x / sqrt(7.5)
    x = np.array([1, 2, 3, 4], np.float32)
    np.testing.assert_allclose(O.rmsnorm(x, np.ones(4, np.float32), 1e-5), x / np.sqrt(7.5), atol=0.01)


# ---------- (b) golden fixtures from the reference's numpy dequantisers ----------
@pytest.mark.parametrize("name", ["q6_k", "q8_0", "q4_k", "q5_k", "f16"])
def test_golden_dequant(name):
    g = np.load(GOLD / f"dequant_{name}.npz")
    rows, cols, dt = int(g["rows"]), int(g["cols"]), int(g["dtype"])
    deq = O.dequant_rows(dt, g["raw"], rows, cols)
    # same F32 arithmetic order up to one rounding: allow 2 ulp-ish relative
    np.testing.assert_allclose(deq, g["deq"], rtol=3e-6, atol=1e-9)


def test_golden_q6_k_quantizer_bytes():
    g = np.load(GOLD / "quantize_q6_k.npz")
    deq = O.dequant_rows(O.Q6_K, g["raw"], 2, 256)
    np.testing.assert_allclose(deq, g["deq"], rtol=3e-6, atol=1e-9)
    assert np.abs(deq - g["w"]).max() < 4e-3    # the reference quantiser is coarse; bytes are what matter


# ---------- (c) independent cross-check: gguf package ----------
@pytest.mark.parametrize("dt,gname", [(DType.Q8_0, "Q8_0"), (DType.Q4_0, "Q4_0"), (DType.Q4_K_M, "Q4_K"),
                                       (DType.Q5_K, "Q5_K"), (DType.Q6_K, "Q6_K")])
def test_dequant_matches_gguf_package(dt, gname):
    gguf = pytest.importorskip("gguf")
    from gguf import quants

    rng = np.random.default_rng(5)
    raw = random_blocks_np(dt, 4, 1024, rng)
    ours = O.dequant_rows(int(dt), raw, 4, 1024)
    theirs = quants.dequantize(raw, getattr(gguf.GGMLQuantizationType, gname)).reshape(4, 1024)
    np.testing.assert_allclose(ours, theirs, rtol=3e-6, atol=1e-9)


# ---------- (d) internal consistency ----------
@pytest.mark.parametrize("dt", ALL_Q)
@pytest.mark.parametrize("shape", [(5, 256), (3, 1024)])
def test_gemv_equals_dequant_dot(dt, shape):
    rows, cols = shape
    rng = np.random.default_rng(11)
    raw = random_blocks_np(dt, rows, cols, rng)
    x = rng.standard_normal(cols).astype(np.float32)
    y = O.gemv(raw, x, rows, cols, int(dt))
    ref = O.dequant_rows(int(dt), raw, rows, cols).astype(np.float64) @ x.astype(np.float64)
    np.testing.assert_allclose(y, ref, rtol=2e-5, atol=2e-6 * np.abs(ref).max())


def test_fp16_roundtrip_all_bit_patterns():
    bits = np.arange(0, 65536, 7, dtype=np.uint16)
    ref = bits.view(np.float16).astype(np.float32)
    got = O.fp16_bits_to_f32(bits)
    m = ~np.isnan(ref)
    np.testing.assert_array_equal(got[m], ref[m])
    back = O.f32_to_fp16_bits(ref[m])
    np.testing.assert_array_equal(back, bits[m])


def test_fp32_to_fp16_matches_numpy_rn():
    rng = np.random.default_rng(3)
    f = np.concatenate([rng.standard_normal(4000).astype(np.float32) * s for s in (1e-7, 1e-4, 1.0, 300.0, 7e4)])
    got = O.f32_to_fp16_bits(f)
    with np.errstate(over="ignore"):
        ref = f.astype(np.float16).view(np.uint16)
    np.testing.assert_array_equal(got, ref)


def test_rope_pairs_and_kv_write():
    rng = np.random.default_rng(9)
    nh, nkv, hd = 4, 2, 8
    q = rng.standard_normal((2, nh, hd)).astype(np.float32)
    k = rng.standard_normal((2, nkv, hd)).astype(np.float32)
    q2, k2 = O.rope(q, k, [0, 3], nh, nkv, hd, 10000.0)
    np.testing.assert_allclose(q2[0], q[0], atol=1e-7)          # position 0 is the identity
    i = 1
    freq = 1.0 / 10000.0 ** (2.0 * i / hd)
    c, s = np.cos(3 * freq), np.sin(3 * freq)
    np.testing.assert_allclose(q2[1, 0, i], q[1, 0, i] * c - q[1, 0, i + hd // 2] * s, rtol=1e-5)
    np.testing.assert_allclose(q2[1, 0, i + hd // 2], q[1, 0, i + hd // 2] * c + q[1, 0, i] * s, rtol=1e-5)
    kc = np.zeros((4, nkv, hd), np.uint16)
    vc = np.zeros((4, nkv, hd), np.uint16)
    O.copy_to_kv_cache(kc, vc, k2, k, 2, nkv, hd, 3, 4)          # second row falls off max_seq and is dropped
    np.testing.assert_array_equal(kc[3].view(np.float16), k2[0].astype(np.float16))
    assert not kc[:3].any()


def test_attention_decode_matches_numpy():
    rng = np.random.default_rng(21)
    nh, nkv, hd, ctx = 8, 2, 16, 37
    q = rng.standard_normal((nh, hd)).astype(np.float32)
    K = rng.standard_normal((ctx, nkv, hd)).astype(np.float16)
    V = rng.standard_normal((ctx, nkv, hd)).astype(np.float16)
    out = O.attention_decode(q, K.view(np.uint16), V.view(np.uint16), ctx, nh, nkv, hd, ctx, 0.25)
    for h in range(nh):
        kv = h // (nh // nkv)
        s = (K[:, kv].astype(np.float64) @ q[h].astype(np.float64)) * 0.25
        p = np.exp(s - s.max())
        p /= p.sum()
        np.testing.assert_allclose(out[h], p @ V[:, kv].astype(np.float64), rtol=1e-5, atol=1e-6)
