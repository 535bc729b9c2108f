"""The decode chain's fusions (round 2) against the oracle, through the C-ABI:
  * nt_b200_attention_decode_fused = launch_rope + launch_copy_to_kv_cache + launch_attention_decode in one launch
    (reference rotary.cu:16-62, attention.cu:316-342, attention.cu:108-202);
  * nt_b200_gemv_fused_f32 = RMSNorm + activation quantiser + GEMV in one launch
    (reference launch pairs of attention.cpp:144-162 / ffn.cpp:96-133);
  * the engine's 6-launch layer against its unfused 8-10-launch twin on the same GGUF.
Tolerances as tests/test_kernels_gpu.py: GEMV <= 2e-5 relative, attention <= 2e-5 absolute; KV-cache rows bit-exact."""
import numpy as np
import pytest
import torch

from ntransformer_b200 import kernels as K
from ntransformer_b200.dtypes import DType
from ntransformer_b200.synth import random_blocks_np
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def sync():
    torch.cuda.synchronize()


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


ATTN_CFGS = [(64, 8, 128), (32, 8, 128), (8, 8, 128), (4, 2, 64), (8, 1, 128), (16, 1, 256)]


@pytest.mark.parametrize("cfg", ATTN_CFGS)
@pytest.mark.parametrize("pos", [0, 4, 31, 32, 63, 64, 299, 2047])
def test_fused_decode_attention_vs_oracle(cfg, pos):
    nh, nkv, hd = cfg
    rng = np.random.default_rng(nh * 131 + pos)
    max_seq, theta = 2048, 500000.0
    kc = (rng.standard_normal((max_seq, nkv, hd)) * 0.7).astype(np.float16)
    vc = rng.standard_normal((max_seq, nkv, hd)).astype(np.float16)
    q = rng.standard_normal((nh, hd)).astype(np.float32)
    k = (rng.standard_normal((nkv, hd)) * 0.7).astype(np.float32)
    v = rng.standard_normal((nkv, hd)).astype(np.float32)
    scale = float(1.0 / np.sqrt(hd))
    # expected: the stand-alone RoPE kernel (bit-identical to the reference's fast-math kernel, tests/test_kernels_gpu.py; the
    # oracle's libm sin/cos differ from it by ~|angle| * 2^-21) -> oracle cache write at pos -> oracle attention over pos + 1 keys
    qr_d, kr_d = dev(q.reshape(1, nh, hd)), dev(k.reshape(1, nkv, hd))
    K.launch_rope(qr_d, kr_d, torch.tensor([pos], device=DEV, dtype=torch.int32), 1, 1, nh, nkv, hd, theta, 1.0, False)
    sync()
    qr, kr = qr_d.cpu().numpy(), kr_d.cpu().numpy()
    qo, ko = O.rope(q.reshape(1, nh, hd), k.reshape(1, nkv, hd), [pos], nh, nkv, hd, theta)
    np.testing.assert_allclose(qr, qo, atol=5e-3)
    kc_o, vc_o = kc.view(np.uint16).copy(), vc.view(np.uint16).copy()
    O.copy_to_kv_cache(kc_o, vc_o, kr, v, 1, nkv, hd, pos, max_seq)
    want = O.attention_decode(qr.reshape(nh, hd), kc_o, vc_o, pos + 1, nh, nkv, hd, max_seq, scale)

    out = torch.zeros(nh, hd, device=DEV)
    kd, vd = dev(kc), dev(vc)
    scratch = torch.zeros(K.attention_decode_scratch_floats(max_seq, nh, nkv, hd), device=DEV)
    tickets = torch.zeros(max(1, K.attention_decode_tickets(nh, nkv)), device=DEV, dtype=torch.int32)
    pos_dev = torch.tensor([pos], device=DEV, dtype=torch.int32)
    xq = torch.zeros(K.xq_bytes(nh * hd), device=DEV, dtype=torch.uint8) if (nh * hd) % 128 == 0 else None
    qd, kn, vn = dev(q), dev(k), dev(v)
    for _ in range(2):          # twice: the tickets must come back zeroed
        K.attention_decode_fused(out, qd, kn, vn, kd, vd, pos_dev, max_seq, nh, nkv, hd, theta, 1.0, scale, scratch, tickets, xq)
        sync()
        np.testing.assert_allclose(out.cpu().numpy(), want, atol=2e-5, rtol=1e-4)
    assert int(tickets.abs().sum()) == 0
    # the cache row written by the kernel is bit-identical to the reference's F16 rounding; q and k are left untouched
    assert np.array_equal(kd.cpu().numpy().view(np.uint16)[pos], kc_o[pos])
    assert np.array_equal(vd.cpu().numpy().view(np.uint16)[pos], vc_o[pos])
    assert np.array_equal(kd.cpu().numpy().view(np.uint16)[:pos], kc.view(np.uint16)[:pos])
    assert np.array_equal(qd.cpu().numpy(), q) and np.array_equal(kn.cpu().numpy(), k)
    if xq is not None:
        # the xq of the output feeds a GEMV: check it through one
        n = nh * hd
        W = random_blocks_np(DType.Q4_K_M, 256, n, rng) if n % 256 == 0 else None
        if W is not None:
            y = torch.zeros(256, device=DEV)
            K.gemv_fused([y], [dev(W)], [256], [DType.Q4_K_M], n, xq, epilogue=0)
            sync()
            assert rel_err(y.cpu().numpy(), O.gemv(W, want.reshape(-1), 256, n, int(DType.Q4_K_M))) <= 5e-5


@pytest.mark.parametrize("dt", [DType.Q4_K_M, DType.Q6_K, DType.Q8_0, DType.Q5_K, DType.Q4_0])
@pytest.mark.parametrize("inn,out", [(4096, 1024), (8192, 4096), (2048, 640), (14336, 512), (28672, 256)])
@pytest.mark.parametrize("norm", [False, True])
def test_gemv_with_the_norm_and_quantiser_in_its_prologue(dt, inn, out, norm):
    """y = W . (RMSNorm(x) * w) from the F32 vector in one launch (the prologue quantises x * w, the epilogue applies the
    1/rms factor) against the oracle's rmsnorm -> gemv; also the plain F32 -> GEMV form and the residual epilogue."""
    rng = np.random.default_rng(inn + out + int(dt))
    eps = 1e-5
    W = random_blocks_np(dt, out, inn, rng)
    x = (rng.standard_normal(inn) * 3.0).astype(np.float32)
    w = (1.0 + 0.1 * rng.standard_normal(inn)).astype(np.float32)
    base = rng.standard_normal(out).astype(np.float32)
    y, yb = torch.zeros(out, device=DEV), dev(base)
    Wd, xd, wd = dev(W), dev(x), dev(w) if norm else None
    try:
        K.gemv_fused_f32([y], [Wd], [out], [dt], inn, xd, epilogue=0, norm_w=wd, eps=eps)
    except ValueError:
        pytest.skip("shape not on the TMA/dp4a path")
    K.gemv_fused_f32([yb], [Wd], [out], [dt], inn, xd, epilogue=1, norm_w=wd, eps=eps)
    sync()
    want = O.gemv(W, O.rmsnorm(x, w, eps) if norm else x, out, inn, int(dt))
    assert rel_err(y.cpu().numpy(), want) <= 3e-5
    assert rel_err(yb.cpu().numpy(), base + want) <= 3e-5
    assert np.array_equal(xd.cpu().numpy(), x)


def test_fused_qkv_and_swiglu_from_the_f32_vector():
    rng = np.random.default_rng(77)
    inn, eps = 8192, 1e-5
    outs, dts = [8192, 1024, 1024], [DType.Q4_K_M, DType.Q4_K_M, DType.Q6_K]
    raws = [random_blocks_np(d, o, inn, rng) for o, d in zip(outs, dts)]
    x = rng.standard_normal(inn).astype(np.float32)
    w = (1.0 + 0.1 * rng.standard_normal(inn)).astype(np.float32)
    ys = [torch.zeros(o, device=DEV) for o in outs]
    K.gemv_fused_f32(ys, [dev(r) for r in raws], outs, dts, inn, dev(x), epilogue=0, norm_w=dev(w), eps=eps)
    sync()
    xn = O.rmsnorm(x, w, eps)
    for y, r, o, d in zip(ys, raws, outs, dts):
        assert rel_err(y.cpu().numpy(), O.gemv(r, xn, o, inn, int(d))) <= 3e-5
    inter = 3584
    g, u = random_blocks_np(DType.Q4_K_M, inter, inn, rng, std=0.05), random_blocks_np(DType.Q4_K_M, inter, inn, rng, std=0.05)
    act, dummy = torch.zeros(inter, device=DEV), torch.zeros(inter, device=DEV)
    K.gemv_fused_f32([act, dummy], [dev(g), dev(u)], [inter, inter], [DType.Q4_K_M] * 2, inn, dev(x), epilogue=2, norm_w=dev(w), eps=eps)
    sync()
    want = O.silu_mul(O.gemv(g, xn, inter, inn, O.Q4_K), O.gemv(u, xn, inter, inn, O.Q4_K))
    assert rel_err(act.cpu().numpy(), want) <= 5e-5


@pytest.mark.parametrize("mix", ["Q4_K_M", "Q8_0", "Q6_K"])
def test_fused_chain_matches_the_unfused_launch_sequence(mix, monkeypatch):
    """Same synthetic model through the 6-launch layer (default) and the unfused sequence (NT_B200_FUSE=0): logits agree to
    round-off (the RMSNorm factor moves across the quantiser), greedy ids identical, both within 1e-3 of the oracle."""
    from ntransformer_b200.engine import Model
    from ntransformer_b200.model_spec import LlamaConfig

    cfg = LlamaConfig(vocab_size=1024, hidden_size=1024, intermediate_size=2048, n_layers=4, n_heads=8, n_kv_heads=2, head_dim=128,
                      max_seq_len=256, bos_token_id=1, eos_token_id=2)

    def run(fuse):
        monkeypatch.setenv("NT_B200_FUSE", str(fuse))
        m = Model.synthetic(cfg, mix, seed=11)
        host = {n: (v[0].cpu().numpy(), int(v[1])) for n, v in m._keep.items()}
        logits, ids, pos = [], [], 0
        tok = cfg.bos_token_id
        for _ in range(140):                     # crosses the 64- and 128-key slice boundaries of the one-launch attention
            l = m.forward([tok], pos).copy()
            logits.append(l)
            tok, pos = int(np.argmax(l)), pos + 1
            ids.append(tok)
        n = m.launches_per_step() if hasattr(m, "launches_per_step") else None
        m.close()
        return logits, ids, host, n

    la, ia, host, _ = run(3)
    lb, ib, _, _ = run(0)
    assert ia == ib
    assert max(rel_err(a, b) for a, b in zip(la, lb)) <= 2e-4
    for mask in (1, 5):                       # the default chain, and RoPE + KV write folded into the decode kernel (two-launch attention)
        lc, ic, _, _ = run(mask)
        assert ic == ib
        assert max(rel_err(a, b) for a, b in zip(lc, lb)) <= 2e-4
    om = O.Model(cfg.dict(), host)
    pos, tok = 0, cfg.bos_token_id
    for i in range(12):
        want = om.forward([tok], pos)
        assert rel_err(la[i], want) <= 1e-3
        tok, pos = ia[i], pos + 1


@pytest.mark.parametrize("dt", [DType.Q4_K_M, DType.Q6_K, DType.Q8_0, DType.Q5_K, DType.Q4_0])
@pytest.mark.parametrize("out", [8192, 7507, 7106, 3680, 3557])
def test_gemv_tail_round_split(dt, out):
    """Row counts that leave a partly filled last round on 148 SMs (K = 8192: 888 warp slots of 4 rows): the tail is dealt out in
    1- or 2-row stages over all slots (gemv_kquant.cu KqParams::tail_nr); ragged last groups included.  Store and residual."""
    rng = np.random.default_rng(out + int(dt))
    inn = 8192
    W = random_blocks_np(dt, out, inn, rng)
    x = rng.standard_normal(inn).astype(np.float32)
    base = rng.standard_normal(out).astype(np.float32)
    y, yb = torch.zeros(out, device=DEV), dev(base)
    Wd, xd = dev(W), dev(x)
    K.gemv_fused_f32([y], [Wd], [out], [dt], inn, xd, epilogue=0)
    K.gemv_fused_f32([yb], [Wd], [out], [dt], inn, xd, epilogue=1)
    sync()
    want = O.gemv(W, x, out, inn, int(dt))
    assert rel_err(y.cpu().numpy(), want) <= 3e-5
    assert rel_err(yb.cpu().numpy(), base + want) <= 3e-5


def test_swiglu_tail_round_split_70b_shape():
    rng = np.random.default_rng(5)
    inn, inter = 8192, 28672
    g, u = random_blocks_np(DType.Q4_K_M, inter, inn, rng, std=0.05), random_blocks_np(DType.Q4_K_M, inter, inn, rng, std=0.05)
    x = rng.standard_normal(inn).astype(np.float32)
    act, dummy = torch.zeros(inter, device=DEV), torch.zeros(inter, device=DEV)
    K.gemv_fused_f32([act, dummy], [dev(g), dev(u)], [inter, inter], [DType.Q4_K_M] * 2, inn, dev(x), epilogue=2)
    sync()
    want = O.silu_mul(O.gemv(g, x, inter, inn, O.Q4_K), O.gemv(u, x, inter, inn, O.Q4_K))
    assert rel_err(act.cpu().numpy(), want) <= 5e-5
