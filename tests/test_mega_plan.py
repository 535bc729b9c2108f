"""CPU tests of the persistent decode kernel's host logic (csrc/engine/decode_mega.h): the per-token program built for
every benchmark shape and tensor-parallel width, and a CPU replay of each GEMV phase's TMA schedule with the same cursor
functions the kernel runs (producer sequence == consumer sequence, every row-group/chunk fetched exactly once, copies
inside their rows, rings inside shared memory).  The kernel itself needs a GPU (tests/test_mega_gpu.py)."""
import ctypes as C

import numpy as np
import pytest

from ntransformer_b200._engine_sigs import ModelConfigC
from ntransformer_b200._lib import lib
from ntransformer_b200.dtypes import DType
from ntransformer_b200.model_spec import LLAMA3_8B, LLAMA3_70B, TINY, LlamaConfig, tensor_dtype

ORDER = ("attn_q", "attn_k", "attn_v", "attn_output", "ffn_gate", "ffn_up", "ffn_down")


def selftest(cfg: LlamaConfig, mix: str, tp_rank=0, tp_size=1, grid=148, split_fixed=0, fuse=0):
    big = cfg.n_layers >= 64
    dts = np.array([[int(tensor_dtype(mix, n, l, cfg.n_layers, big)) for n in ORDER] for l in range(cfg.n_layers)], dtype=np.int32)
    head = int(tensor_dtype(mix, "output", 0, cfg.n_layers, big))
    c = ModelConfigC(**cfg.dict())
    info = (C.c_int * 8)()
    msg = C.create_string_buffer(512)
    rc = lib().nt_mega_plan_selftest(C.byref(c), tp_rank, tp_size, dts.ctypes.data_as(C.c_void_p), head, grid, split_fixed, fuse, info, msg, 512)
    return rc, list(info), msg.value.decode()


@pytest.mark.parametrize("cfg,mix", [(LLAMA3_70B, "Q4_K_M"), (LLAMA3_70B, "Q6_K"), (LLAMA3_8B, "Q4_K_M"), (LLAMA3_8B, "Q8_0"),
                                     (LLAMA3_8B, "Q5_K")])
@pytest.mark.parametrize("tp", [1, 2, 4, 8])
def test_plan_and_schedule_for_benchmark_shapes(cfg, mix, tp):
    for rank in {0, tp - 1}:
        rc, info, msg = selftest(cfg, mix, rank, tp)
        assert rc == 0, (cfg.hidden_size, mix, tp, rank, msg)
        n_phases, n_body, n_gemv, min_warps, min_stages, n_xchg, n_splits, max_split = info
        L = cfg.n_layers
        assert n_body == 9 * L and n_phases == n_body + 2            # 9 phases per layer + final norm + LM head
        assert n_gemv == 4 * L + 1 and n_xchg == 2 * L               # o-proj and down-proj end in an exchange
        assert min_warps >= 4 and min_stages >= 2                     # every GEMV phase keeps a double-buffered ring
        assert 64 <= max_split <= 256 and n_splits >= 1


MID = LlamaConfig(vocab_size=1024, hidden_size=2048, intermediate_size=4096, n_layers=3, n_heads=16, n_kv_heads=4, head_dim=128,
                  max_seq_len=256, bos_token_id=1, eos_token_id=2)
MID_G8 = LlamaConfig(vocab_size=1024, hidden_size=2048, intermediate_size=4096, n_layers=2, n_heads=16, n_kv_heads=2, head_dim=128,
                     max_seq_len=256, bos_token_id=1, eos_token_id=2)


def test_small_test_models_and_compat_split_rule():
    """The shapes tests/test_mega_gpu.py runs: TINY (head_dim 64) with pure Q4_K rows (its Q6_K rows of 420 bytes are not
    TMA-aligned, the graph path sends them to the generic GEMV), and two head_dim-128 models with the Q4_K_M mix."""
    rc, _, msg = selftest(TINY, "Q4_K_M")
    assert rc == 1 and "K-quant" in msg
    for cfg, mix in ((TINY, "Q4_K"), (MID, "Q4_K_M"), (MID_G8, "Q4_K_M"), (MID, "Q8_0")):
        rc, info, msg = selftest(cfg, mix)
        assert rc == 0, (mix, msg)
        rc, info, msg = selftest(cfg, mix, split_fixed=4)             # the graph path's fixed split count
        assert rc == 0 and info[6] == 4, msg
        rc, info, msg = selftest(cfg, mix, 1, 2)
        assert rc == 0, msg


def test_small_grids_and_odd_grids_keep_the_schedule_consistent():
    for grid in (32, 33, 100, 132, 148, 160):
        rc, info, msg = selftest(LLAMA3_8B, "Q4_K_M", grid=grid)
        assert rc == 0, (grid, msg)


def test_uncovered_shapes_are_rejected_not_mangled():
    rc, _, msg = selftest(LLAMA3_8B, "F16")                           # F16 weights are not on the K-quant TMA path
    assert rc == 1 and "K-quant" in msg
    rc, _, msg = selftest(LLAMA3_8B, "Q4_0")                          # Q4_0 is on the dp4a path of the persistent kernel
    assert rc == 0, msg
    rc, _, msg = selftest(LLAMA3_8B, "F32")
    assert rc == 1
    odd = LlamaConfig(vocab_size=512, hidden_size=512, intermediate_size=1024, n_layers=2, n_heads=6, n_kv_heads=2, head_dim=64, max_seq_len=128)
    rc, _, msg = selftest(odd, "Q4_K_M")                              # 3 query heads per KV head: attention not instantiated
    assert rc == 1 and "attention" in msg
    rc, _, msg = selftest(LLAMA3_8B, "Q4_K_M", grid=8)                # hidden / 256 CTAs are needed by the norm phase
    assert rc == 1


def test_fusion_flags_shorten_the_program():
    for fuse, per_layer in ((0, 9), (1, 8), (2, 8), (3, 7), (7, 7)):          # bit 4 is ignored under tensor parallelism
        rc, info, msg = selftest(LLAMA3_70B, "Q4_K_M", 0, 8, fuse=fuse)
        assert rc == 0, msg
        assert info[1] == per_layer * LLAMA3_70B.n_layers and info[2] == 4 * LLAMA3_70B.n_layers + 1
    L = LLAMA3_70B.n_layers
    rc, info, msg = selftest(LLAMA3_70B, "Q4_K_M", fuse=7)                    # single rank: 5 phases per layer + the first norm
    assert rc == 0, msg
    assert info[1] == 5 * L + 1 and info[0] == info[1] + 1 and info[2] == 4 * L + 1
    rc, info, msg = selftest(LLAMA3_8B, "Q4_K_M", fuse=4)
    assert rc == 0 and info[1] == 7 * LLAMA3_8B.n_layers + 1, msg


def test_overlap_flag_keeps_the_plan_consistent_at_benchmark_shapes():
    for cfg in (LLAMA3_70B, LLAMA3_8B):
        for tp in (1, 8):
            rc, info, msg = selftest(cfg, "Q4_K_M", 0, tp, fuse=63)
            assert rc == 0, (cfg.hidden_size, tp, msg)
            assert info[3] >= 4 and info[4] >= 2           # the o-projection keeps >= 4 warps and a double-buffered ring


def test_split_tail_schedules_cover_every_row_once():
    """fuse bit 32 on the benchmark shapes and a few awkward grids: the CPU replay inside nt_mega_plan_selftest checks that the
    1- and 2-row tail stages fetch every row of every matrix exactly once and never straddle a 32-row block."""
    for cfg, mix in ((LLAMA3_70B, "Q4_K_M"), (LLAMA3_70B, "Q6_K"), (LLAMA3_8B, "Q4_K_M"), (LLAMA3_8B, "Q8_0")):
        for tp in (1, 2, 8):
            rc, info, msg = selftest(cfg, mix, tp - 1, tp, fuse=32)
            assert rc == 0, (mix, tp, msg)
    for grid in (33, 100, 132, 160):
        rc, info, msg = selftest(LLAMA3_8B, "Q4_K_M", grid=grid, fuse=63)
        assert rc == 0, (grid, msg)
