"""tcgen05/TMEM prefill GEMM (csrc/prefill_gemm.cu) against a plain torch reference of the same op.

The weights are F16 (exact in the reference too); activations are F32 split into two F16 terms, so the only error
left is the F32 accumulation order: tolerance 1e-4 * max|C| against an F64 matmul (measured 4e-5 at K=14336) (north_star allows 1e-3)."""
import pytest
import torch

from ntransformer_b200 import kernels as K

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
    assert err <= 1e-4 * ref.abs().max().item(), (err, ref.abs().max().item())


def test_gemm_f16_tc_rejects_unaligned_shapes():
    A = torch.zeros(4, 96, device="cuda")
    W = torch.zeros(100, 96, device="cuda").half()
    Cm = torch.zeros(4, 100, device="cuda")
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        K.gemm_f16_tc(Cm, A, W, 4, 100, 96, ws)
