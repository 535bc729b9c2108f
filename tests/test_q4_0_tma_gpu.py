"""Q4_0 on the TMA/dp4a GEMV path (the default since round 2; process_stage<4> in csrc/gemv_kq_device.cuh) against the oracle,
and the generic kernel (NT_B200_Q4_0_TMA=0) as the A/B twin.  The switch is read once per process, hence the subprocesses."""
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


@pytest.mark.parametrize("tma", ["1", "0"])
def test_q4_0_gemv_matches_the_oracle(tma):
    env = dict(os.environ, NT_B200_Q4_0_TMA=tma)
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-3000:]
