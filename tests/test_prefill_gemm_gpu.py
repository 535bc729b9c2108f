"""tcgen05/TMEM prefill GEMM (csrc/prefill_gemm.cu) against a plain torch reference of the same op.

The weights are F16 (exact in the reference too); activations are F32 split into two F16 terms, so the only error
left is the F32 accumulation: tolerance 2.5e-5 * max|C| against an F64 matmul for the store / residual GEMMs (K-chunked
accumulation), 1e-4 for the SwiGLU-fused one (single accumulator over K) (north_star allows 1e-3)."""
import numpy as np
import pytest
import torch

from ntransformer_b200 import kernels as K
from ntransformer_b200.dtypes import DType, dtype_row_size
from ntransformer_b200.synth import random_blocks_np
from oracle import oracle as O

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 64), (128, 128, 256), (256, 384, 512), (1, 128, 64), (77, 256, 4096), (300, 1024, 1024),
          (512, 4096, 4096), (130, 14336, 4096), (96, 4096, 14336)]


@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_f16_tc_vs_torch_f64(shape):
    M, N, Kd = shape
    g = torch.Generator(device="cuda").manual_seed(M * 131 + N * 7 + Kd)
    A = torch.randn(M, Kd, device="cuda", generator=g) * 1.7
    W = (torch.randn(N, Kd, device="cuda", generator=g) * 0.05).half()
    Cm = torch.full((M, N), float("nan"), device="cuda")
    ws = torch.empty(K.gemm_f16_tc_workspace_bytes(M, Kd), dtype=torch.uint8, device="cuda")
    K.gemm_f16_tc(Cm, A, W, M, N, Kd, ws)
    torch.cuda.synchronize()
    ref = A.double() @ W.double().T
    err = (Cm.double() - ref).abs().max().item()
    # K-chunked accumulation (4096 elements per fresh TMEM accumulator, partial sums added in F32 RN): round 1's single
    # accumulator measured 4e-5 at K = 14336 (tensor cores truncate once per MMA step); chunked it stays below 2.5e-5
    assert err <= 2.5e-5 * ref.abs().max().item(), (err, ref.abs().max().item())


def test_gemm_f16_tc_rejects_unaligned_shapes():
    A = torch.zeros(4, 96, device="cuda")
    W = torch.zeros(100, 96, device="cuda").half()
    Cm = torch.zeros(4, 100, device="cuda")
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        K.gemm_f16_tc(Cm, A, W, 4, 100, 96, ws)


def _unsplit(ws, M, N):
    """hi + lo (as F64) of a split workspace holding an [M, N] matrix (rows padded to 128)."""
    Mp = (M + 127) // 128 * 128
    h = ws.view(torch.float16)[: 2 * Mp * N].view(2, Mp, N)
    return h[0, :M].double() + h[1, :M].double()


@pytest.mark.parametrize("shape", [(128, 128, 64), (200, 1024, 512), (1000, 14336, 4096), (4096, 3584, 1024)])
def test_gemm_add_epilogue_and_many_tiles_per_cta(shape):
    """Residual epilogue (C += A.W^T); the larger shapes give each persistent CTA several tiles (both TMEM accumulators)."""
    M, N, Kd = shape
    g = torch.Generator(device="cuda").manual_seed(N + Kd)
    A = torch.randn(M, Kd, device="cuda", generator=g)
    W = (torch.randn(N, Kd, device="cuda", generator=g) * 0.05).half()
    C0 = torch.randn(M, N, device="cuda", generator=g)
    Cm = C0.clone()
    ws = torch.empty(K.gemm_f16_tc_workspace_bytes(M, Kd), dtype=torch.uint8, device="cuda")
    K.split_activations(ws, A, M, Kd)
    K.gemm_f16_tc_ws(Cm, ws, W, M, N, Kd, add=True)
    torch.cuda.synchronize()
    ref = C0.double() + A.double() @ W.double().T
    assert (Cm.double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


@pytest.mark.parametrize("shape", [(128, 128, 64), (77, 1024, 512), (600, 14336, 4096)])
def test_gemm_swiglu_epilogue_vs_torch_f64(shape):
    M, N, Kd = shape
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, Kd, device="cuda", generator=g)
    Wg = (torch.randn(N, Kd, device="cuda", generator=g) * 0.03).half()
    Wu = (torch.randn(N, Kd, device="cuda", generator=g) * 0.03).half()
    ws_in = torch.empty(K.gemm_f16_tc_workspace_bytes(M, Kd), dtype=torch.uint8, device="cuda")
    ws_out = torch.empty(K.gemm_f16_tc_workspace_bytes(M, N), dtype=torch.uint8, device="cuda")
    K.split_activations(ws_in, A, M, Kd)
    K.gemm_f16_tc_swiglu_ws(ws_out, ws_in, Wg, Wu, M, N, Kd)
    torch.cuda.synchronize()
    gate, up = A.double() @ Wg.double().T, A.double() @ Wu.double().T
    ref = gate / (1.0 + torch.exp(-gate)) * up
    got = _unsplit(ws_out, M, N)
    assert (got - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


def test_rmsnorm_split_vs_torch():
    rows, hidden = 300, 4096
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(rows, hidden, device="cuda", generator=g) * 3
    w = 1 + 0.1 * torch.randn(hidden, device="cuda", generator=g)
    ws = torch.empty(K.gemm_f16_tc_workspace_bytes(rows, hidden), dtype=torch.uint8, device="cuda")
    K.rmsnorm_split(ws, x, w, rows, hidden, 1e-5)
    torch.cuda.synchronize()
    xd = x.double()
    ref = xd * torch.rsqrt((xd * xd).mean(dim=1, keepdim=True) + 1e-5) * w.double()
    assert (_unsplit(ws, rows, hidden) - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()


@pytest.mark.parametrize("dt", [DType.Q4_K_M, DType.Q5_K, DType.Q6_K, DType.Q8_0, DType.Q4_0, DType.F16, DType.F32])
def test_dequant_split_vs_oracle_and_gemm_over_it(dt):
    """GGUF blocks -> F16 hi/lo pair: hi + lo reproduces the oracle's dequantised row to F32 rounding, and the two-GEMM
    product over the pair matches an F64 matmul with the oracle's weights (the quantised-model prefill path)."""
    rows, cols, M = 256, 1024, 70
    rng = np.random.default_rng(int(dt) + 40)
    raw = random_blocks_np(dt, rows, cols, rng)
    want = O.dequant_rows(int(dt), raw, rows, cols)
    hi = torch.empty(rows, cols, dtype=torch.float16, device="cuda")
    lo = torch.empty(rows, cols, dtype=torch.float16, device="cuda")
    K.dequant_split(hi, lo, torch.from_numpy(raw).cuda(), dt, rows, cols)
    torch.cuda.synchronize()
    got = hi.double().cpu().numpy() + lo.double().cpu().numpy()
    assert np.abs(got - want).max() <= 4e-7 * np.abs(want).max()
    A = torch.randn(M, cols, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    Cm = torch.empty(M, rows, device="cuda")
    ws = torch.empty(K.gemm_f16_tc_workspace_bytes(M, cols), dtype=torch.uint8, device="cuda")
    K.split_activations(ws, A, M, cols)
    K.gemm_f16_tc_ws(Cm, ws, hi, M, rows, cols)
    K.gemm_f16_tc_ws(Cm, ws, lo, M, rows, cols, add=True)
    torch.cuda.synchronize()
    ref = A.double().cpu().numpy() @ want.astype(np.float64).T
    assert np.abs(Cm.double().cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
