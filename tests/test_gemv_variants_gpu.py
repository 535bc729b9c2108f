"""Opt-in GEMV kernel variants against the oracle (their switches are read once per process, hence the subprocesses):
  NT_B200_GEMV_QB=1      the quarter-block kernel (csrc/gemv_kquant_q.cu: lane <-> 64 weights, 16 warps) — measured slower than the
                         default half-block kernel (profiles/r02_qb_layer_ncu_summary.txt), kept as an A/B aid;
  NT_B200_TAIL_SPLIT=0   whole 4-row stages in a partly filled last round (the pre-round-2 schedule)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = [pytest.mark.gpu]
ROOT = Path(__file__).resolve().parent.parent

SCRIPT = r'''
import numpy as np, torch, sys
sys.path.insert(0, %r)
from ntransformer_b200 import kernels as K
from ntransformer_b200.dtypes import DType
from ntransformer_b200.synth import random_blocks_np
from oracle import oracle as O
rng = np.random.default_rng(1)
for dt in (DType.Q4_K_M, DType.Q5_K, DType.Q6_K, DType.Q8_0, DType.Q4_0):
    for out, inn in ((256, 4096), (8192, 8192), (7507, 8192), (640, 2048), (512, 28672), (130, 1024)):
        raw = random_blocks_np(dt, out, inn, rng)
        x = rng.standard_normal(inn).astype(np.float32)
        w = (1.0 + 0.1 * rng.standard_normal(inn)).astype(np.float32)
        y = torch.zeros(out, device="cuda")
        try:
            K.gemv_fused_f32([y], [torch.from_numpy(raw).cuda()], [out], [dt], inn, torch.from_numpy(x).cuda(), norm_w=torch.from_numpy(w).cuda(), eps=1e-5)
        except ValueError:
            continue
        torch.cuda.synchronize()
        ref = O.gemv(raw, O.rmsnorm(x, w, 1e-5), out, inn, int(dt))
        err = float(np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max())
        assert err < 3e-5, (dt, out, inn, err)
# mixed q/k (Q4_K) + v (Q6_K) launch and the SwiGLU epilogue
inn = 8192
outs, dts = [8192, 1024, 1024], [DType.Q4_K_M, DType.Q4_K_M, DType.Q6_K]
raws = [random_blocks_np(d, o, inn, rng) for o, d in zip(outs, dts)]
x = rng.standard_normal(inn).astype(np.float32)
ys = [torch.zeros(o, device="cuda") for o in outs]
K.gemv_fused_f32(ys, [torch.from_numpy(r).cuda() for r in raws], outs, dts, inn, torch.from_numpy(x).cuda())
torch.cuda.synchronize()
for y, r, o, d in zip(ys, raws, outs, dts):
    ref = O.gemv(r, x, o, inn, int(d))
    assert float(np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max()) < 3e-5
g, u = random_blocks_np(DType.Q4_K_M, 3584, inn, rng, std=0.05), random_blocks_np(DType.Q4_K_M, 3584, inn, rng, std=0.05)
act, dummy = torch.zeros(3584, device="cuda"), torch.zeros(3584, device="cuda")
K.gemv_fused_f32([act, dummy], [torch.from_numpy(g).cuda(), torch.from_numpy(u).cuda()], [3584, 3584], [DType.Q4_K_M] * 2, inn, torch.from_numpy(x).cuda(), epilogue=2)
torch.cuda.synchronize()
want = O.silu_mul(O.gemv(g, x, 3584, inn, O.Q4_K), O.gemv(u, x, 3584, inn, O.Q4_K))
assert float(np.abs(act.cpu().numpy() - want).max() / np.abs(want).max()) < 5e-5
print("ok")
''' % str(ROOT)


@pytest.mark.parametrize("env", [{"NT_B200_GEMV_QB": "1"}, {"NT_B200_TAIL_SPLIT": "0"}, {"NT_B200_GEMV_QB": "1", "NT_B200_GEMV_QB_SPLIT": "0"}])
def test_gemv_variant_matches_the_oracle(env):
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=dict(os.environ, **env), timeout=900)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-3000:]
