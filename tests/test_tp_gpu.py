"""GPU tensor-parallel parity (needs >= 2 GPUs, run with `gpurun --gpus 2|4|8`): the TP engine must reproduce the single-GPU
engine's logits (<= 1e-3 relative; only the summation grouping of the o-proj / down-proj partials differs) and its greedy
ids, with the partial sums exchanged through NVLink peer memory (default, engine/peer_xchg.h) and through ncclAllReduce
(NT_B200_TP_NCCL=1).  8 KV heads so that 8-way sharding (one KV head per GPU, BASELINE configs[3]) is covered."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from ntransformer_b200.engine import Model
from ntransformer_b200.gguf_write import synthetic_tensors_np, write_gguf
from ntransformer_b200.model_spec import LlamaConfig

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
CFG = LlamaConfig(vocab_size=2048, hidden_size=2048, intermediate_size=4096, n_layers=3, n_heads=16, n_kv_heads=8, head_dim=128,
                  max_seq_len=128, bos_token_id=1, eos_token_id=2)


@pytest.mark.parametrize("mix,world,nccl", [("Q4_K_M", 2, 0), ("Q4_K_M", 2, 1), ("Q6_K", 2, 0), ("Q8_0", 2, 0), ("Q4_K_M", 4, 0), ("Q6_K", 8, 0),
                                            ("Q4_K_M", 8, 0), ("Q6_K", 8, 1)])
def test_tp_matches_single_gpu(tmp_path, mix, world, nccl):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    path = tmp_path / f"{mix}.gguf"
    write_gguf(path, CFG, synthetic_tensors_np(CFG, mix, seed=21))
    m = Model.load(path, CFG.max_seq_len)
    prompt = [1, 17, 300, 5, 44, 9]
    want = [m.forward(prompt, 0).copy()]
    ids, tok, pos = [], int(np.argmax(want[0])), len(prompt)
    for _ in range(12):
        ids.append(tok)
        l = m.forward([tok], pos).copy()
        want.append(l)
        tok, pos = int(np.argmax(l)), pos + 1
    m.close()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = tmp_path / "tp.npz"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "tp_worker.py"), str(path), str(out),
                        str(CFG.max_seq_len)], capture_output=True, text=True, errors="replace", timeout=240,
                       env=dict(os.environ, NT_B200_TP_NCCL=str(nccl), NT_B200_XCHG_TIMEOUT_MS="5000"))
    assert r.returncode == 0, r.stderr[-3000:]
    got = np.load(out)
    assert list(got["ids"]) == ids
    for a, b in zip(got["logits"], want):
        assert np.abs(a - b).max() / np.abs(b).max() <= 1e-3
