"""Time the tcgen05 prefill GEMM: python tools/prof_gemm.py M N K [reps] [store|swiglu]."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ntransformer_b200 import kernels as K

M, N, Kd = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
mode = sys.argv[5] if len(sys.argv) > 5 else "store"
A = torch.randn(M, Kd, device="cuda")
W = (torch.randn(N, Kd, device="cuda") * 0.05).half()
W2 = (torch.randn(N, Kd, device="cuda") * 0.05).half()
Cm = torch.empty(M, N, device="cuda")
ws = torch.empty(K.gemm_f16_tc_workspace_bytes(M, Kd), dtype=torch.uint8, device="cuda")
ws2 = torch.empty(K.gemm_f16_tc_workspace_bytes(M, N), dtype=torch.uint8, device="cuda")
K.split_activations(ws, A, M, Kd)


def run():
    if mode == "swiglu":
        K.gemm_f16_tc_swiglu_ws(ws2, ws, W, W2, M, N, Kd)
    else:
        K.gemm_f16_tc_ws(Cm, ws, W, M, N, Kd)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = (4.0 if mode == "swiglu" else 2.0) * M * N * Kd
print(f"gemm_f16_tc[{mode}] M={M} N={N} K={Kd}: {ms * 1e3:.1f} us  {flops / ms / 1e9:.1f} TFLOP/s algorithmic "
      f"({2 * flops / ms / 1e9:.1f} issued, hi+lo split)")
if mode == "store":
    print("max abs err", (Cm.double() - A.double() @ W.double().T).abs().max().item())
