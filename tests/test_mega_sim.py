"""The persistent decode kernel's OWN source (csrc/engine/decode_megakernel.cu), compiled with g++ and executed on the CPU
emulator tests/cusim (fibers for threads, OS threads for CTAs, emulated mbarrier/TMA, host memory as global and peer memory),
against the CPU oracle on the same weights.  This is how the kernel's control flow — phase interpreter, ring priming across
barriers, mbarrier parities, grid barriers, split attention with in-phase RoPE/KV write, combine, slot/residual ping-pong and
the tensor-parallel flag exchange — was debugged while no GPU was available.  It says nothing about the memory model, the
hardware TMA path or speed; tests/test_mega_gpu.py covers the real thing."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from ntransformer_b200.gguf_write import synthetic_tensors_np, write_gguf
from ntransformer_b200.model_spec import TINY, LlamaConfig
from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent
SIM_DIR = ROOT / "tests" / "cusim"

SMALL128 = LlamaConfig(vocab_size=512, hidden_size=2048, intermediate_size=2048, n_layers=2, n_heads=16, n_kv_heads=4, head_dim=128,
                       max_seq_len=160, bos_token_id=1, eos_token_id=2)
SMALL128_G8 = LlamaConfig(vocab_size=512, hidden_size=2048, intermediate_size=2048, n_layers=2, n_heads=16, n_kv_heads=2, head_dim=128,
                          max_seq_len=160, bos_token_id=1, eos_token_id=2)


@pytest.fixture(scope="session")
def sim():
    r = subprocess.run(["make", "-C", str(SIM_DIR)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.fail("tests/cusim does not build:\n" + r.stderr[-3000:])
    lib = C.CDLL(str(SIM_DIR / "_sim" / "libmega_sim.so"))
    lib.mega_sim_create.restype = C.c_void_p
    lib.mega_sim_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    lib.mega_sim_free.argtypes = [C.c_void_p]
    lib.mega_sim_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.mega_sim_read.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int]
    return lib


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


class SimModel:
    def __init__(self, lib, path, cfg, host, tp=1, grid=4, split_fixed=0, copy_delay=0, fuse=0):
        self.lib, self.cfg = lib, cfg
        msg = C.create_string_buffer(512)
        self.h = lib.mega_sim_create(str(path).encode(), cfg.max_seq_len, tp, grid, split_fixed, fuse, copy_delay, msg, 512)
        assert self.h, msg.value.decode()
        emb, dt = host["token_embd.weight"]
        self.table = O.dequant_rows(dt, emb, cfg.vocab_size, cfg.hidden_size)

    def step(self, token, pos, with_head=True):
        row = np.ascontiguousarray(self.table[token], dtype=np.float32)
        out = np.zeros(self.cfg.vocab_size, dtype=np.float32)
        rc = self.lib.mega_sim_step(self.h, row.ctypes.data_as(C.c_void_p), int(token), int(pos), int(with_head),
                                    out.ctypes.data_as(C.c_void_p))
        assert rc == 0, {1: "emulator deadlock / watchdog", 2: "a barrier inside the kernel timed out"}.get(rc, rc)
        return out

    def close(self):
        self.lib.mega_sim_free(self.h)


def make_case(tmp_path, cfg, mix, seed=31):
    tensors = synthetic_tensors_np(cfg, mix, seed=seed)
    path = tmp_path / f"{mix}.gguf"
    write_gguf(path, cfg, tensors)
    host = {n: (np.ascontiguousarray(a), int(dt)) for n, (a, dt, r, c) in tensors.items()}
    return path, host


def check_against_oracle(lib, tmp_path, cfg, mix, steps, tol=2e-4, **kw):
    path, host = make_case(tmp_path, cfg, mix)
    om = O.Model(cfg.dict(), host)
    m = SimModel(lib, path, cfg, host, **kw)
    prompt = [cfg.bos_token_id, 17, 300, 5]
    pos = 0
    for t in prompt[:-1]:                                   # prompt tokens: body only (no LM head), like Model::forward_async
        m.step(t, pos, with_head=False)
        om.forward([t], pos)
        pos += 1
    got = m.step(prompt[-1], pos)
    want = om.forward([prompt[-1]], pos)
    pos += 1
    worst = rel(got, want)
    tok = int(np.argmax(want))
    for _ in range(steps):
        assert int(np.argmax(got)) == tok
        got = m.step(tok, pos)
        want = om.forward([tok], pos)
        worst = max(worst, rel(got, want))
        tok, pos = int(np.argmax(want)), pos + 1
    m.close()
    assert worst <= tol, worst
    return worst


def test_tiny_hd64_single_rank(sim, tmp_path):
    check_against_oracle(sim, tmp_path, TINY, "Q4_K", steps=6, grid=3)


def test_hd128_q4_k_m_mix_with_async_copies(sim, tmp_path):
    # Q4_K / Q6_K mix in one launch (half-filled chunks: 8 super-blocks per row), TMA copies completing out of order
    check_against_oracle(sim, tmp_path, SMALL128, "Q4_K_M", steps=3, grid=8, copy_delay=7)


def test_hd128_eight_heads_per_kv_and_q8_0(sim, tmp_path):
    check_against_oracle(sim, tmp_path, SMALL128_G8, "Q8_0", steps=2, grid=8)


def test_context_crossing_split_boundaries(sim, tmp_path):
    # adaptive rule: 64 keys per split -> contexts of 65+ use two splits and the combine phase merges them
    cfg = LlamaConfig(**{**TINY.dict(), "n_layers": 1, "max_seq_len": 160})
    path, host = make_case(tmp_path, cfg, "Q4_K")
    om = O.Model(cfg.dict(), host)
    m = SimModel(sim, path, cfg, host, grid=4)
    rng = np.random.default_rng(0)
    toks = [int(t) for t in rng.integers(3, cfg.vocab_size, size=140)]
    for pos, t in enumerate(toks):
        head = pos in (0, 62, 63, 64, 65, 127, 128, 129, 139)
        got = m.step(t, pos, with_head=head)
        want = om.forward([t], pos)
        if head:
            assert rel(got, want) <= 2e-4, pos
    m.close()


def test_compat_split_rule(sim, tmp_path):
    # the graph path's rule (context cut into a fixed number of splits): many tiny splits, partly empty
    check_against_oracle(sim, tmp_path, TINY, "Q4_K", steps=10, grid=3, split_fixed=4)


TP_CFG = LlamaConfig(vocab_size=512, hidden_size=1024, intermediate_size=1024, n_layers=2, n_heads=16, n_kv_heads=4, head_dim=64,
                     max_seq_len=128, bos_token_id=1, eos_token_id=2)


@pytest.mark.parametrize("tp,cfg,grid,delay", [(2, TINY, 3, 0), (4, TP_CFG, 4, 5), (2, SMALL128, 8, 0)])
def test_tensor_parallel_exchange(sim, tmp_path, tp, cfg, grid, delay):
    # tp emulated GPUs in one process: partial o-proj / down-proj rows pushed into every peer's slots, flag round trip
    # between the master CTAs, identical residual stream on every rank, LM-head shards gathered by the harness
    check_against_oracle(sim, tmp_path, cfg, "Q4_K", steps=4 if cfg is not SMALL128 else 1, tp=tp, grid=grid, copy_delay=delay, tol=5e-4)


def test_ragged_rows_odd_grid_and_uneven_vocab_shards(sim, tmp_path):
    # vocab 510: the LM head's last row-group is ragged (rows % 4 != 0); under TP-2 the shards are 255 rows each;
    # grid 5: row-groups do not divide evenly over the CTAs, some CTAs idle in the last round
    cfg = LlamaConfig(**{**TINY.dict(), "vocab_size": 510, "n_layers": 2})
    check_against_oracle(sim, tmp_path, cfg, "Q4_K", steps=3, grid=5)
    check_against_oracle(sim, tmp_path, cfg, "Q4_K", steps=2, tp=2, grid=3, copy_delay=3, tol=5e-4)


@pytest.mark.parametrize("fuse", [1, 2, 3])
def test_producer_side_fusions(sim, tmp_path, fuse):
    """MEGA_FUSE_QUANT (1): the SwiGLU epilogue's last arriver quantises each 32-row block; MEGA_FUSE_COMBINE (2): the last
    split unit of a head group merges the splits.  Same arithmetic, two barriers fewer per layer."""
    check_against_oracle(sim, tmp_path, SMALL128, "Q4_K_M", steps=2, grid=8, copy_delay=4, fuse=fuse)
    check_against_oracle(sim, tmp_path, TINY, "Q4_K", steps=4, tp=2, grid=3, fuse=fuse, tol=5e-4)


def test_fused_combine_across_split_boundaries(sim, tmp_path):
    cfg = LlamaConfig(**{**TINY.dict(), "n_layers": 1, "max_seq_len": 160})
    path, host = make_case(tmp_path, cfg, "Q4_K")
    om = O.Model(cfg.dict(), host)
    m = SimModel(sim, path, cfg, host, grid=4, fuse=3)
    rng = np.random.default_rng(1)
    toks = [int(t) for t in rng.integers(3, cfg.vocab_size, size=132)]
    for pos, t in enumerate(toks):
        head = pos in (0, 63, 64, 65, 128, 131)
        got = m.step(t, pos, with_head=head)
        want = om.forward([t], pos)
        if head:
            assert rel(got, want) <= 2e-4, pos
    m.close()


@pytest.mark.parametrize("fuse", [4, 7])
def test_fused_residual_and_norm_single_rank(sim, tmp_path, fuse):
    """MEGA_FUSE_NORM (4): at one rank the o-proj / down-proj epilogue's last arriver adds the block to the residual stream,
    stores its sum of squares and quantises h * w for the next norm; the consuming GEMV applies 1/rms to its results.
    5 phases per layer with all three fusions.  Mathematically equal, not bit-equal, to the separate norm phase."""
    check_against_oracle(sim, tmp_path, SMALL128, "Q4_K_M", steps=3, grid=8, copy_delay=4, fuse=fuse)
    check_against_oracle(sim, tmp_path, TINY, "Q4_K", steps=8, grid=3, fuse=fuse)
    # under tensor parallelism the flag is ignored (the norm needs every rank's partial rows): still correct
    check_against_oracle(sim, tmp_path, TINY, "Q4_K", steps=2, tp=2, grid=3, fuse=fuse, tol=5e-4)


def test_multi_chunk_rows_with_ragged_last_chunk(sim, tmp_path):
    """The 70B-like row geometries: hidden 4608 = 18 super-blocks -> two chunks per row, the second only 2 super-blocks wide
    (idle lanes); intermediate 12544 = 49 super-blocks -> four chunks for the down projection (8 of 12 warps busy, last chunk
    1 super-block).  36 query heads / 9 KV heads, one layer, grid 18 (the norm phase needs hidden / 256 CTAs)."""
    cfg = LlamaConfig(vocab_size=512, hidden_size=4608, intermediate_size=12544, n_layers=1, n_heads=36, n_kv_heads=9, head_dim=128,
                      max_seq_len=64, bos_token_id=1, eos_token_id=2)
    check_against_oracle(sim, tmp_path, cfg, "Q4_K", steps=1, grid=18, copy_delay=3, fuse=3)


def test_eight_way_tensor_parallel(sim, tmp_path):
    """The benchmark's widest configuration in miniature: 8 emulated ranks, one KV head and a 256-column o-proj / down-proj
    shard each (one super-block per row: 2 of 32 lanes busy), every exchange an 8-way flag round trip."""
    cfg = LlamaConfig(vocab_size=512, hidden_size=2048, intermediate_size=2048, n_layers=1, n_heads=32, n_kv_heads=8, head_dim=64,
                      max_seq_len=64, bos_token_id=1, eos_token_id=2)
    check_against_oracle(sim, tmp_path, cfg, "Q4_K", steps=1, tp=8, grid=8, fuse=3, tol=5e-4)


@pytest.mark.parametrize("fuse", [8, 11])
def test_distributed_reduce_phase_single_rank(sim, tmp_path, fuse):
    """MEGA_DEFER_RMS (8): the norm phases become one-warp-per-block reduce phases (always on under tensor parallelism, where
    every test above with tp > 1 already runs them); here at one rank, alone and with the quantiser/combine fusions."""
    check_against_oracle(sim, tmp_path, SMALL128, "Q4_K_M", steps=3, grid=5, copy_delay=4, fuse=fuse)
    check_against_oracle(sim, tmp_path, TINY, "Q4_K", steps=6, grid=1, fuse=fuse)      # a single CTA is a valid grid now


def test_q4_0_blocks_on_the_dp4a_path(sim, tmp_path):
    """Q4_0 (18-byte blocks, reference K1 gemm.cu:32-90) through process_stage<4>: the format the stand-alone GEMV still sends
    to the generic kernel unless NT_B200_Q4_0_TMA=1."""
    check_against_oracle(sim, tmp_path, TINY, "Q4_0", steps=4, grid=3, copy_delay=2)
    check_against_oracle(sim, tmp_path, SMALL128, "Q4_0", steps=1, grid=8, fuse=3)


@pytest.mark.parametrize("fuse", [18, 31])
def test_weights_stream_through_the_attention_phase(sim, tmp_path, fuse):
    """MEGA_OVERLAP_ATTN (16, with the fused combine): the attention scratch sits above the o-projection's smaller rings, which
    are primed at the end of the q/k/v phase — TMA copies land in shared memory while the attention phase computes next to them."""
    check_against_oracle(sim, tmp_path, SMALL128, "Q4_K_M", steps=3, grid=8, copy_delay=9, fuse=fuse)
    check_against_oracle(sim, tmp_path, TINY, "Q4_K", steps=4, tp=2, grid=3, copy_delay=5, fuse=fuse, tol=5e-4)
    cfg = LlamaConfig(**{**TINY.dict(), "n_layers": 1, "max_seq_len": 160})      # contexts of 2-3 splits
    path, host = make_case(tmp_path, cfg, "Q4_K")
    om = O.Model(cfg.dict(), host)
    m = SimModel(sim, path, cfg, host, grid=4, fuse=fuse, copy_delay=6)
    rng = np.random.default_rng(2)
    for pos, t in enumerate(int(t) for t in rng.integers(3, cfg.vocab_size, size=131)):
        head = pos in (0, 64, 65, 130)
        got = m.step(t, pos, with_head=head)
        want = om.forward([t], pos)
        if head:
            assert rel(got, want) <= 2e-4, pos
    m.close()


def test_direct_exchange_protocol(sim, tmp_path):
    """MEGA_XCHG_DIRECT (128): every CTA adds itself to every rank's never-reset arrival counter and waits for
    sequence * tp * grid arrivals on its own — no master CTA, no second hop."""
    check_against_oracle(sim, tmp_path, TINY, "Q4_K", steps=6, tp=2, grid=3, copy_delay=3, fuse=128 | 3, tol=5e-4)
    check_against_oracle(sim, tmp_path, TP_CFG, "Q4_K", steps=3, tp=4, grid=4, fuse=128 | 51, tol=5e-4)
    cfg = LlamaConfig(vocab_size=512, hidden_size=2048, intermediate_size=2048, n_layers=1, n_heads=32, n_kv_heads=8, head_dim=64,
                      max_seq_len=64, bos_token_id=1, eos_token_id=2)
    check_against_oracle(sim, tmp_path, cfg, "Q4_K", steps=2, tp=8, grid=8, fuse=128, tol=5e-4)


def test_l2_prefetch_flag_is_functionally_neutral(sim, tmp_path):
    # MEGA_L2_PREFETCH (64) only adds prefetch hints (no-ops in the emulator): the plan must stay consistent and results equal
    check_against_oracle(sim, tmp_path, TINY, "Q4_K", steps=2, tp=2, grid=3, fuse=64 | 3, tol=5e-4)


@pytest.mark.parametrize("fuse", [32, 63])
def test_split_tail_rounds(sim, tmp_path, fuse):
    """MEGA_SPLIT_TAIL (32): the partly filled last round of a GEMV phase is dealt out 1 or 2 rows at a time over all warp slots
    (process_stage<FMT, 1>), tickets of the last-arriver fusions count rows.  Different grids give different tail shapes."""
    for grid in (3, 4, 5, 7):
        check_against_oracle(sim, tmp_path, TINY, "Q4_K", steps=2, grid=grid, copy_delay=3, fuse=fuse)
    check_against_oracle(sim, tmp_path, SMALL128, "Q4_K_M", steps=2, grid=8, copy_delay=4, fuse=fuse)
    check_against_oracle(sim, tmp_path, SMALL128_G8, "Q8_0", steps=1, grid=11, fuse=fuse)
    cfg = LlamaConfig(**{**TINY.dict(), "vocab_size": 510, "n_layers": 2})     # ragged LM head + uneven TP shards
    check_against_oracle(sim, tmp_path, cfg, "Q4_K", steps=2, tp=2, grid=3, copy_delay=3, fuse=fuse, tol=5e-4)
    check_against_oracle(sim, tmp_path, cfg, "Q4_0", steps=2, grid=5, fuse=fuse)


def test_a_missing_peer_times_out_instead_of_hanging(sim, tmp_path, monkeypatch):
    """Rank 1 never launches: rank 0's master CTA gives up on the flag after the time-out, raises the abort word, every later
    barrier falls through and the launch ends (the host then reports 'decode megakernel aborted')."""
    path, host = make_case(tmp_path, TINY, "Q4_K")
    m = SimModel(sim, path, TINY, host, tp=2, grid=3)
    monkeypatch.setenv("CUSIM_MEGA_SKIP_RANK", "1")
    monkeypatch.setenv("CUSIM_MEGA_TIMEOUT_MS", "200")
    row = np.ascontiguousarray(m.table[1], dtype=np.float32)
    out = np.zeros(TINY.vocab_size, dtype=np.float32)
    rc = sim.mega_sim_step(m.h, row.ctypes.data_as(C.c_void_p), 1, 0, 1, out.ctypes.data_as(C.c_void_p))
    assert rc == 2                                       # the abort word, not the emulator's watchdog (1) and not success
    m.close()


@pytest.mark.skipif(os.environ.get("CUSIM_SHUFFLE") is not None, reason="already running under a shuffled schedule")
def test_results_do_not_depend_on_the_thread_schedule(sim):
    """CUSIM_SHUFFLE: the emulator visits the threads of a CTA in a fresh pseudo-random order every scheduling pass, so a
    missing __syncthreads / __syncwarp that in-order execution would hide shows up as a wrong result.  (The switch is read
    once per process: run a slice of this file in a child.)"""
    env = dict(os.environ, CUSIM_SHUFFLE="7")
    r = subprocess.run([sys.executable, "-m", "pytest", str(Path(__file__)), "-x", "-q", "-k",
                        "tiny_hd64 or producer_side_fusions or eight_way or fused_residual"], capture_output=True, text=True, env=env,
                       cwd=str(ROOT), timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:]


def _mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


@pytest.mark.skipif(_mem_available_gb() < 24 and os.environ.get("NT_B200_SIM_FULLSIZE") != "1",
                    reason="one Llama-3 70B layer on 148 emulated CTAs needs ~8 GB of fiber stacks (NT_B200_SIM_FULLSIZE=1 forces it)")
def test_one_70b_layer_on_148_ctas(sim, tmp_path):
    """Production geometry: hidden 8192, 64 query / 8 KV heads, intermediate 28672 (7 chunks per down-projection row), Q4_K_M
    mix, 148 CTAs x 384 threads — one layer, small vocabulary."""
    cfg = LlamaConfig(vocab_size=1024, hidden_size=8192, intermediate_size=28672, n_layers=1, n_heads=64, n_kv_heads=8, head_dim=128,
                      max_seq_len=64, bos_token_id=1, eos_token_id=2)
    for fuse in (0, 63):
        check_against_oracle(sim, tmp_path, cfg, "Q6_K" if fuse == 0 else "Q4_K", steps=1, grid=148, copy_delay=3, fuse=fuse)
