"""Time the tcgen05 prefill GEMM: python tools/prof_gemm.py M N K [reps]."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ntransformer_b200 import kernels as K

M, N, Kd = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
A = torch.randn(M, Kd, device="cuda")
W = (torch.randn(N, Kd, device="cuda") * 0.05).half()
Cm = torch.empty(M, N, device="cuda")
ws = torch.empty(K.gemm_f16_tc_workspace_bytes(M, Kd), dtype=torch.uint8, device="cuda")
for _ in range(3):
    K.gemm_f16_tc(Cm, A, W, M, N, Kd, ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    K.gemm_f16_tc(Cm, A, W, M, N, Kd, ws)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"gemm_f16_tc M={M} N={N} K={Kd}: {ms * 1e3:.1f} us  {2.0 * M * N * Kd / ms / 1e9:.1f} TFLOP/s useful "
      f"({4.0 * M * N * Kd / ms / 1e9:.1f} issued, hi+lo split)")
err = (Cm.double() - A.double() @ W.double().T).abs().max().item()
print("max abs err", err)
