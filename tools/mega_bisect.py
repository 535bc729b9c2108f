"""Bring-up aid for the persistent decode kernel: find the first phase whose result on the GPU differs from the CPU emulation
of the very same program (tests/cusim runs the kernel's own source).

    python tools/mega_bisect.py [--model tiny|mid] [--mix Q4_K] [--fuse 0] [--tol 1e-5]

Runs on a GPU box.  For k = 1 .. n_phases it decodes ONE token (position 0, fresh KV cache) with the program truncated after
k phases (NT_B200_MEGA_MAX_PHASES) on the GPU and in the emulator and compares the working vectors (hid0, hid1, q, k, v, attn,
act).  Truncated runs leave later buffers untouched, so the first k at which a buffer differs names the guilty phase."""
import argparse
import ctypes as C
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
KINDS = {0: "norm+quantise", 1: "quantise", 2: "GEMV", 3: "attention", 4: "combine", 5: "reduce+quantise"}
BUFS = ("hid0", "hid1", "q", "k", "v", "attn", "act")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="mid", choices=["tiny", "mid"])
    ap.add_argument("--mix", default="Q4_K")
    ap.add_argument("--fuse", type=int, default=0)
    ap.add_argument("--tol", type=float, default=1e-4)      # fast-math exp / division on the GPU vs libm in the emulator
    ap.add_argument("--token", type=int, default=17)
    args = ap.parse_args()
    os.environ["NT_B200_MEGA_FUSE"] = str(args.fuse)
    from ntransformer_b200.engine import Model
    from ntransformer_b200.gguf_write import synthetic_tensors_np, write_gguf
    from ntransformer_b200.model_spec import TINY, LlamaConfig
    from oracle import oracle as O

    cfg = TINY if args.model == "tiny" else LlamaConfig(vocab_size=1024, hidden_size=2048, intermediate_size=4096, n_layers=2, n_heads=16,
                                                        n_kv_heads=4, head_dim=128, max_seq_len=128, bos_token_id=1, eos_token_id=2)
    subprocess.run(["make", "-C", str(ROOT / "tests" / "cusim")], check=True, capture_output=True)
    sim = C.CDLL(str(ROOT / "tests" / "cusim" / "_sim" / "libmega_sim.so"))
    sim.mega_sim_create.restype = C.c_void_p
    sim.mega_sim_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    sim.mega_sim_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    sim.mega_sim_read.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int]
    sim.mega_sim_free.argtypes = [C.c_void_p]
    tmp = Path(tempfile.mkdtemp())
    tensors = synthetic_tensors_np(cfg, args.mix, seed=31)
    path = tmp / "m.gguf"
    write_gguf(path, cfg, tensors)
    emb, dt = tensors["token_embd.weight"][:2]
    row = np.ascontiguousarray(O.dequant_rows(int(dt), np.ascontiguousarray(emb), cfg.vocab_size, cfg.hidden_size)[args.token], dtype=np.float32)

    import torch
    grid = torch.cuda.get_device_properties(0).multi_processor_count
    sim_grid = min(grid, 8 if cfg.hidden_size <= 2048 else 32)          # the program does not depend on the grid size, the schedule does
    probe = Model.load(path, cfg.max_seq_len)
    probe.use_megakernel(True)
    probe.forward([args.token], 0)
    assert probe.megakernel_active, "persistent kernel not active for this model"
    kinds = probe.megakernel_plan()
    probe.close()
    print(f"{len(kinds)} phases; GPU grid {grid}, emulator grid {sim_grid}")
    for k in range(1, len(kinds) + 1):
        os.environ["NT_B200_MEGA_MAX_PHASES"] = str(k)
        g = Model.load(path, cfg.max_seq_len)
        g.use_megakernel(True)
        g.forward([args.token], 0)
        msg = C.create_string_buffer(256)
        h = sim.mega_sim_create(str(path).encode(), cfg.max_seq_len, 1, sim_grid, 0, args.fuse, 0, msg, 256)
        out = np.zeros(cfg.vocab_size, dtype=np.float32)
        assert sim.mega_sim_step(h, row.ctypes.data_as(C.c_void_p), args.token, 0, 1, out.ctypes.data_as(C.c_void_p)) == 0
        worst, where = 0.0, None
        for name in BUFS:
            a = g.debug_read(name)
            b = np.zeros(len(a), dtype=np.float32)
            sim.mega_sim_read(h, 0, name.encode(), b.ctypes.data_as(C.c_void_p), len(b))
            scale = max(float(np.abs(b).max()), 1e-30)
            err = float(np.abs(a - b).max()) / scale if np.isfinite(a).all() else float("inf")
            if err > worst:
                worst, where = err, name
        sim.mega_sim_free(h)
        g.close()
        flag = "OK " if worst <= args.tol else "BAD"
        print(f"{flag} after phase {k - 1:3d} ({KINDS[kinds[k - 1]]:>16}): worst relative difference {worst:.3e} in {where}")
        if worst > args.tol:
            print("first divergence: phase", k - 1, KINDS[kinds[k - 1]])
            return 1
    print("GPU and emulation agree after every phase")
    return 0


if __name__ == "__main__":
    sys.exit(main())
