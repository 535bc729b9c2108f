"""CPU test of the tensor-parallel plan with world_size 2 over gloo: every rank computes one transformer block on
its shards with the oracle's kernels, the two partial sums are all-reduced, and the result must equal the
unsharded oracle (same products, different summation grouping -> 1e-5 relative)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ntransformer_b200.dtypes import DType
from ntransformer_b200.model_spec import LlamaConfig, tensor_table
from ntransformer_b200.synth import random_blocks_np
from ntransformer_b200.tp import shard_tensor, split_kind
from oracle import oracle as O

CFG = LlamaConfig(vocab_size=256, hidden_size=512, intermediate_size=1024, n_layers=1, n_heads=8, n_kv_heads=2, head_dim=64,
                  max_seq_len=32, bos_token_id=1, eos_token_id=2)


def full_weights(mix):
    out = {}
    for idx, (name, dt, rows, cols) in enumerate(tensor_table(CFG, mix)):
        rng = np.random.default_rng(100 + idx)
        if name.endswith("norm.weight"):
            out[name] = ((1 + 0.1 * rng.standard_normal(cols)).astype(np.float32), dt, rows, cols)
        else:
            out[name] = (random_blocks_np(dt, rows, cols, rng), dt, rows, cols)
    return out


def block_forward(w, x, rank, size, allreduce):
    """One decoder block at position 0 (attention over a single key == V) on rank's shard."""
    hd, nh, nkv = CFG.head_dim, CFG.n_heads // size, CFG.n_kv_heads // size

    def W(n):
        arr, dt, rows, cols = w[f"blk.0.{n}.weight"]
        return shard_tensor(arr, dt, rows, cols, n, rank, size) + (int(dt),)

    h = x.copy()
    xn = O.rmsnorm(h, w["blk.0.attn_norm.weight"][0], CFG.norm_eps)
    q = O.gemv(*W("attn_q")[:1], xn, W("attn_q")[1], W("attn_q")[2], W("attn_q")[3])
    v = O.gemv(*W("attn_v")[:1], xn, W("attn_v")[1], W("attn_v")[2], W("attn_v")[3])
    assert q.shape[0] == nh * hd
    attn = np.repeat(v.reshape(nkv, hd), nh // nkv, axis=0).reshape(-1)      # softmax over one key is 1
    wo = W("attn_output")
    part = O.gemv(wo[0], attn.astype(np.float32), wo[1], wo[2], wo[3])
    h = h + allreduce(part)
    xn = O.rmsnorm(h, w["blk.0.ffn_norm.weight"][0], CFG.norm_eps)
    g, u, d = W("ffn_gate"), W("ffn_up"), W("ffn_down")
    act = O.silu_mul(O.gemv(g[0], xn, g[1], g[2], g[3]), O.gemv(u[0], xn, u[1], u[2], u[3]))
    part = O.gemv(d[0], act, d[1], d[2], d[3])
    return h + allreduce(part)


def _worker(rank, size, port, mix, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    w = full_weights(mix)
    x = np.random.default_rng(7).standard_normal(CFG.hidden_size).astype(np.float32)

    def allreduce(p):
        t = torch.from_numpy(p.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    out = block_forward(w, x, rank, size, allreduce)
    gathered = [torch.zeros(CFG.hidden_size) for _ in range(size)]
    dist.all_gather(gathered, torch.from_numpy(out))
    if rank == 0:
        q.put([g.numpy() for g in gathered])
    dist.destroy_process_group()


@pytest.mark.parametrize("mix", ["Q4_K_M", "Q8_0"])
def test_two_rank_block_matches_unsharded(mix):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mix, q)) for r in range(2)]
    [p.start() for p in procs]
    outs = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    w = full_weights(mix)
    x = np.random.default_rng(7).standard_normal(CFG.hidden_size).astype(np.float32)
    want = block_forward(w, x, 0, 1, lambda p: p)
    np.testing.assert_array_equal(outs[0], outs[1])                 # ranks stay in lock-step (same reduction everywhere)
    assert np.abs(outs[0] - want).max() / np.abs(want).max() < 1e-5


def test_shards_tile_the_tensor_at_block_boundaries():
    rng = np.random.default_rng(0)
    for dt, rows, cols in [(DType.Q4_K_M, 8, 2048), (DType.Q6_K, 6, 1024), (DType.Q8_0, 5, 256), (DType.F16, 4, 64)]:
        raw = random_blocks_np(dt, rows, cols, rng)
        full = O.dequant_rows(int(dt), raw, rows, cols)
        for size in (2, 4):
            parts = [shard_tensor(raw, dt, rows, cols, "blk.0.ffn_down.weight", r, size) for r in range(size)]
            deq = np.concatenate([O.dequant_rows(int(dt), p[0], p[1], p[2]) for p in parts], axis=1)
            np.testing.assert_array_equal(deq, full)               # column slices are pure byte slices: no requantisation
            rparts = [shard_tensor(raw, dt, rows, cols, "blk.0.attn_q.weight", r, size) for r in range(size)]
            assert sum(p[1] for p in rparts) == rows
    assert split_kind("blk.3.attn_norm.weight") == "replicate" and split_kind("token_embd.weight") == "replicate"
    assert split_kind("output.weight") == "rows" and split_kind("blk.0.attn_output.weight") == "cols"
