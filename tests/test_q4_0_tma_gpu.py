"""Q4_0 on the TMA/dp4a GEMV path (NT_B200_Q4_0_TMA=1; process_stage<4> in csrc/gemv_kq_device.cuh) against the oracle.

GATED (NT_B200_TEST_UNVERIFIED=1 or NT_B200_TEST_MEGA=1): written after round 1's GPU budget was spent; its arithmetic is
verified through the CPU emulation of the persistent kernel (tests/test_mega_sim.py::test_q4_0_blocks_on_the_dp4a_path).  The
switch is read once per process, hence the subprocess."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("NT_B200_TEST_MEGA") != "1" and os.environ.get("NT_B200_TEST_UNVERIFIED") != "1",
                                 reason="Q4_0 TMA path: opt-in until verified on hardware (NT_B200_TEST_UNVERIFIED=1)")]
ROOT = Path(__file__).resolve().parent.parent

SCRIPT = r'''
import numpy as np, torch, sys
sys.path.insert(0, %r)
from ntransformer_b200 import kernels as K
from ntransformer_b200.dtypes import DType
from ntransformer_b200.synth import random_blocks_np
from oracle import oracle as O
rng = np.random.default_rng(0)
before = K.launch_count()
for out, inn in ((256, 4096), (1024, 8192), (130, 2048), (4096, 14336)):
    raw = random_blocks_np(DType.Q4_0, out, inn, rng)
    x = rng.standard_normal(inn).astype(np.float32)
    y = torch.zeros(out, device="cuda")
    K.launch_gemv(y, torch.from_numpy(raw).cuda(), torch.from_numpy(x).cuda(), out, inn, DType.Q4_0)
    torch.cuda.synchronize()
    ref = O.gemv(raw, x, out, inn, int(DType.Q4_0))
    err = float(np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max())
    assert err < 2e-5, (out, inn, err)
print("ok")
''' % str(ROOT)


def test_q4_0_gemv_on_the_tma_path_matches_the_oracle():
    env = dict(os.environ, NT_B200_Q4_0_TMA="1")
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-3000:]
