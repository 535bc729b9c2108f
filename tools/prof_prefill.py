"""Short batched-prefill run for ncu: a Llama-3 8B shaped model with few layers, one prompt through the tensor-core path.

    python tools/prof_prefill.py --layers 2 --mix F16 --tokens 2048"""
import argparse
import dataclasses
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--mix", default="F16")
    ap.add_argument("--tokens", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch

    from ntransformer_b200.engine import Model
    from ntransformer_b200.model_spec import LLAMA3_8B

    cfg = dataclasses.replace(LLAMA3_8B, n_layers=args.layers, max_seq_len=args.tokens + 64)
    m = Model.synthetic(cfg, args.mix, seed=1)
    prompt = [(i * 7919 + 11) % cfg.vocab_size for i in range(args.tokens)]
    t0 = time.time()
    for _ in range(args.reps):
        m.forward(prompt, 0)
    torch.cuda.synchronize()
    print(json.dumps({"layers": args.layers, "mix": args.mix, "tokens": args.tokens, "ms_per_prompt_wall": round((time.time() - t0) / args.reps * 1e3, 2)}))
    m.close()


if __name__ == "__main__":
    main()
