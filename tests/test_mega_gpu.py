"""GPU parity of the persistent decode kernel (csrc/engine/decode_mega.h) against the verified graph path and the CPU oracle.

First hardware run: round 2 (all single-GPU cases green on a B200).  The persistent kernel stays an opt-in decode path
(NT_B200_MEGAKERNEL=1 / nt_model_use_megakernel) because it is correct but slower than the graph of fused launches
(profiles/r02_mega_*); these tests keep it correct.

What they assert:
  * compat split rule (NT_B200_MEGA_SPLIT_COMPAT=1): every phase is a transplant of a graph-path kernel with the same
    arithmetic, so logits must agree with the graph path to float round-off (<= 1e-5 relative; bit-equality is reported);
  * adaptive split rule (the default): <= 1e-3 relative against the oracle's forward and the same greedy ids as the graph
    path, through contexts that cross split boundaries;
  * the abort word stays clear (no barrier time-out)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]

from ntransformer_b200.engine import Model  # noqa: E402
from ntransformer_b200.gguf_write import synthetic_tensors_np, write_gguf  # noqa: E402
from ntransformer_b200.model_spec import TINY, LlamaConfig  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent
MID = LlamaConfig(vocab_size=1024, hidden_size=2048, intermediate_size=4096, n_layers=3, n_heads=16, n_kv_heads=4, head_dim=128,
                  max_seq_len=256, bos_token_id=1, eos_token_id=2)
MID_G8 = LlamaConfig(vocab_size=1024, hidden_size=2048, intermediate_size=4096, n_layers=2, n_heads=16, n_kv_heads=2, head_dim=128,
                     max_seq_len=256, bos_token_id=1, eos_token_id=2)
CASES = [(TINY, "Q4_K"), (MID, "Q4_K_M"), (MID_G8, "Q4_K_M"), (MID, "Q8_0"), (MID, "Q6_K")]


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def run(model, prompt, steps):
    """per-token prompt replay (the persistent kernel is the decode step) + greedy continuation; returns logits, ids"""
    logits, pos = [], 0
    for t in prompt:
        l = model.forward([t], pos).copy()
        pos += 1
    logits.append(l)
    ids, tok = [], int(np.argmax(l))
    for _ in range(steps):
        ids.append(tok)
        l = model.forward([tok], pos).copy()
        logits.append(l)
        tok, pos = int(np.argmax(l)), pos + 1
    return logits, ids


@pytest.fixture(scope="module", params=range(len(CASES)))
def case(request, tmp_path_factory):
    cfg, mix = CASES[request.param]
    tensors = synthetic_tensors_np(cfg, mix, seed=31)
    path = tmp_path_factory.mktemp("mega") / f"{mix}.gguf"
    write_gguf(path, cfg, tensors)
    host = {n: (np.ascontiguousarray(a), int(dt)) for n, (a, dt, r, c) in tensors.items()}
    return cfg, mix, path, host


def test_compat_mode_matches_graph_path(case, monkeypatch):
    cfg, mix, path, _ = case
    prompt = [cfg.bos_token_id, 17, 300, 5, 44, 9, 12, 400]
    # the persistent kernel's phases are transplants of the UNFUSED launch sequence (rmsnorm_xq -> GEMV over xq ...): compare with that
    # form of the graph path (NT_B200_FUSE=0), not with the round-2 chain that moves the norm factor across the GEMV
    monkeypatch.setenv("NT_B200_FUSE", "0")
    g = Model.load(path, cfg.max_seq_len)
    want, ids_g = run(g, prompt, 40)
    g.close()
    monkeypatch.setenv("NT_B200_MEGA_SPLIT_COMPAT", "1")
    m = Model.load(path, cfg.max_seq_len)
    m.use_megakernel(True)
    got, ids_m = run(m, prompt, 40)
    assert m.megakernel_active, "the persistent kernel rejected a shape it is meant to cover"
    kinds = m.megakernel_plan()
    assert len(kinds) == 9 * cfg.n_layers + 2
    m.close()
    worst = max(rel(a, b) for a, b in zip(got, want))
    exact = all(np.array_equal(a, b) for a, b in zip(got, want))
    print(f"{mix}: max rel diff vs graph path {worst:.3e}, bit-identical: {exact}")
    assert worst <= 1e-4          # same slices, same per-slice arithmetic; the graph path's merge kernel adds the slices in another order since round 2
    assert ids_m == ids_g


def test_adaptive_split_vs_oracle_and_graph_ids(case):
    from oracle import oracle as O

    cfg, mix, path, host = case
    prompt = [cfg.bos_token_id, 21, 7, 350, 90]
    steps = min(100, cfg.max_seq_len - len(prompt) - 1)           # crosses the 64-key split boundary of the adaptive rule
    g = Model.load(path, cfg.max_seq_len)
    want, ids_g = run(g, prompt, steps)
    g.close()
    m = Model.load(path, cfg.max_seq_len)
    m.use_megakernel(True)
    got, ids_m = run(m, prompt, steps)
    assert m.megakernel_active
    hid = m.debug_read("hid0")
    assert hid.shape == (cfg.hidden_size,) and np.isfinite(hid).all()
    m.close()
    assert max(rel(a, b) for a, b in zip(got, want)) <= 1e-3
    assert ids_m == ids_g
    om = O.Model(cfg.dict(), host)
    pos = 0
    for t in prompt:
        lo = om.forward([t], pos)
        pos += 1
    assert rel(got[0], lo) <= 1e-3
    for i in range(8):
        lo = om.forward([ids_m[i]], pos)
        pos += 1
        assert rel(got[i + 1], lo) <= 1e-3


@pytest.mark.parametrize("fuse", [3, 7, 31, 63])
def test_producer_side_fusions_keep_the_greedy_ids(case, monkeypatch, fuse):
    """NT_B200_MEGA_FUSE: 1 activation quantiser, 2 split combine (same arithmetic as the separate phases), 4 residual add +
    next norm at one rank (1/rms applied by the consumer: equal up to round-off, not bit-equal)."""
    cfg, mix, path, _ = case
    prompt = [cfg.bos_token_id, 33, 8, 120]
    g = Model.load(path, cfg.max_seq_len)
    want, ids_g = run(g, prompt, 70)
    g.close()
    monkeypatch.setenv("NT_B200_MEGA_FUSE", str(fuse))
    m = Model.load(path, cfg.max_seq_len)
    m.use_megakernel(True)
    got, ids_m = run(m, prompt, 70)
    assert m.megakernel_active
    per_layer = {3: 7, 7: 5, 31: 5, 63: 5}[fuse]
    assert len(m.megakernel_plan()) == per_layer * cfg.n_layers + 2      # + first/final norm and the LM head
    m.close()
    assert max(rel(a, b) for a, b in zip(got, want)) <= 1e-3
    assert ids_m == ids_g


def test_uncovered_model_keeps_the_graph_path(tmp_path):
    tensors = synthetic_tensors_np(TINY, "F16", seed=3)
    path = tmp_path / "f16.gguf"
    write_gguf(path, TINY, tensors)
    m = Model.load(path, TINY.max_seq_len)
    m.use_megakernel(True)
    l = m.forward([1], 0).copy()
    assert not m.megakernel_active and np.isfinite(l).all()
    m.close()


@pytest.mark.parametrize("world", [2, 4])
def test_tensor_parallel_peer_exchange_matches_single_gpu(tmp_path, world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cfg = LlamaConfig(vocab_size=2048, hidden_size=2048, intermediate_size=4096, n_layers=3, n_heads=16, n_kv_heads=4, head_dim=128,
                      max_seq_len=128, bos_token_id=1, eos_token_id=2)
    if (cfg.n_heads // world) // (cfg.n_kv_heads // world) != 4:
        pytest.skip("shape")
    path = tmp_path / "tp.gguf"
    write_gguf(path, cfg, synthetic_tensors_np(cfg, "Q4_K_M", seed=21))
    m = Model.load(path, cfg.max_seq_len)
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
    env = dict(os.environ, NT_B200_MEGAKERNEL="1", NT_B200_NO_BATCHED_PREFILL="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "tp_worker.py"), str(path), str(out),
                        str(cfg.max_seq_len)], capture_output=True, text=True, errors="replace", timeout=240, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    got = np.load(out)
    assert list(got["ids"]) == ids
    for a, b in zip(got["logits"], want):
        assert np.abs(a - b).max() / np.abs(b).max() <= 1e-3
