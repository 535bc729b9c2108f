"""Short decode run for ncu: a Llama-3 shaped model with few layers (the kernels of one layer are the kernels of all layers).

    python tools/prof_decode.py --model 70b --layers 8 --mix Q4_K_M [--mega 1] [--ctx 2048] [--tokens 6]

--ctx N starts decoding at position N-1 (the KV cache rows below are zeros: same traffic, same instruction stream).
Prints nothing but a one-line summary; meant to be wrapped in `ncu ... python tools/prof_decode.py ...` under gpurun."""
import argparse
import dataclasses
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="70b", choices=["70b", "8b"])
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--mix", default="Q4_K_M")
    ap.add_argument("--mega", type=int, default=0)
    ap.add_argument("--fuse", type=int, default=0)
    ap.add_argument("--ctx", type=int, default=64)
    ap.add_argument("--tokens", type=int, default=6)
    ap.add_argument("--max-seq", type=int, default=4096)
    ap.add_argument("--big-mix", type=int, default=1, help="use the >=64-layer Q4_K_M recipe (Q5_K attn_v) even with few layers")
    ap.add_argument("--shard-of", type=int, default=1,
                    help="g > 1: a single-GPU model with the per-rank shapes of g-way tensor parallelism (heads/g, kv/g, inter/g, vocab/g; hidden "
                         "unchanged) — what one rank computes per token, without any exchange: the no-communication bound of TP-g")
    args = ap.parse_args()
    os.environ["NT_B200_MEGA_FUSE"] = str(args.fuse)
    if args.shard_of > 1:
        os.environ["NT_B200_SYNTH_PAD"] = "1"          # narrow rows padded to a 16-byte pitch, like the real tensor-parallel shards
    import torch

    from ntransformer_b200 import model_spec
    from ntransformer_b200.engine import Model
    from ntransformer_b200.model_spec import LLAMA3_8B, LLAMA3_70B

    base = {"70b": LLAMA3_70B, "8b": LLAMA3_8B}[args.model]
    cfg = dataclasses.replace(base, n_layers=args.layers, max_seq_len=args.max_seq)
    if args.shard_of > 1:
        g = args.shard_of
        cfg = dataclasses.replace(cfg, n_heads=base.n_heads // g, n_kv_heads=base.n_kv_heads // g, intermediate_size=base.intermediate_size // g,
                                  vocab_size=-(-base.vocab_size // g))
    m = Model.synthetic(cfg, args.mix, seed=1)
    if args.mega:
        m.use_megakernel(True)
    pos = args.ctx - 1
    t0 = time.time()
    for i in range(args.tokens):
        m.forward([(i * 7919 + 5) % cfg.vocab_size], pos + i)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / args.tokens
    # device-timed steady state (graph replays, no host sync between steps)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.ExternalStream(m.stream)
    n = 64
    ev0.record(stream)
    for i in range(n):
        m.forward_async([(i * 7919 + 5) % cfg.vocab_size], pos + args.tokens + i)
    ev1.record(stream)
    m.sync()
    ms = ev0.elapsed_time(ev1) / n
    print(json.dumps({"model": args.model, "layers": args.layers, "mix": args.mix, "mega": bool(args.mega and m.megakernel_active),
                      "shard_of": args.shard_of, "ctx": args.ctx, "ms_per_token_wall": round(dt * 1e3, 3), "ms_per_token_device": round(ms, 4),
                      "us_per_layer": round(ms * 1e3 / args.layers, 2)}))
    m.close()


if __name__ == "__main__":
    main()
