#!/usr/bin/env python3
"""Launches the fused K-quant GEMV a few times on synthetic weights (for ncu captures).
   python tools/prof_gemv.py [rows] [K] [dtype] [epilogue] [reps]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ntransformer_b200 import kernels as K
from ntransformer_b200.dtypes import DType
from ntransformer_b200.synth import random_blocks_cuda

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 28672
Kin = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
dt = DType[sys.argv[3]] if len(sys.argv) > 3 else DType.Q4_K_M
ep = int(sys.argv[4]) if len(sys.argv) > 4 else 2
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 6
nmat = 2 if ep == 2 else 1
Ws = [[random_blocks_cuda(dt, rows, Kin, 100 + 7 * r + m) for m in range(nmat)] for r in range(3)]
x = torch.randn(Kin, device="cuda")
xq = torch.zeros(K.xq_bytes(Kin), device="cuda", dtype=torch.uint8)
K.quantize_x(x, xq, Kin)
ys = [torch.zeros(rows, device="cuda") for _ in range(nmat)]
torch.cuda.synchronize()
evs = []
for r in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    K.gemv_fused(ys, Ws[r % 3], [rows] * nmat, [dt] * nmat, Kin, xq, epilogue=ep)
    b.record()
    evs.append((a, b))
torch.cuda.synchronize()
nbytes = sum(w.numel() for w in Ws[0])
for a, b in evs:
    ms = a.elapsed_time(b)
    print(f"{rows}x{Kin} {dt.name} ep={ep}: {ms * 1e3:.1f} us  {nbytes / ms / 1e6:.0f} GB/s")
