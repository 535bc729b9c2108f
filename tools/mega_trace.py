"""Phase timeline of the persistent decode kernel: where a token's time goes, per phase kind.

    python tools/mega_trace.py [--workload llama3-70b-q4_k_m-decode] [--fuse 0] [--tokens 24]

Runs on a GPU box (gpurun).  Prints, for the last decoded token, the mean time per phase kind split into work (phase start ->
work done on the slowest traced CTA) and barrier wait (work done -> barrier passed), per layer and per token."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
KINDS = {0: "norm+quantise", 1: "quantise", 2: "GEMV", 3: "attention", 4: "combine"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="70b", choices=["70b", "8b", "tiny"])
    ap.add_argument("--mix", default="Q4_K_M")
    ap.add_argument("--fuse", type=int, default=0)
    ap.add_argument("--tokens", type=int, default=24)
    args = ap.parse_args()
    os.environ["NT_B200_MEGA_FUSE"] = str(args.fuse)
    from ntransformer_b200.engine import Model
    from ntransformer_b200.model_spec import LLAMA3_8B, LLAMA3_70B, LlamaConfig

    cfg = {"70b": LLAMA3_70B, "8b": LLAMA3_8B,
           "tiny": LlamaConfig(vocab_size=1024, hidden_size=2048, intermediate_size=4096, n_layers=3, n_heads=16, n_kv_heads=4, head_dim=128,
                               max_seq_len=256)}[args.model]
    m = Model.synthetic(cfg, args.mix, seed=1)
    m.use_megakernel(True)
    m.forward([1], 0)
    assert m.megakernel_active, "persistent kernel not active for this model"
    m.megakernel_trace(True)
    for pos in range(1, args.tokens):
        m.forward([(pos * 7919) % cfg.vocab_size], pos)
    ticks, ns_per_tick = m.megakernel_trace_read()
    t = ticks.astype(np.float64) * ns_per_tick[:, None, None]   # [4, phases, 3] in ns, per-CTA clock (durations only)
    kinds = m.megakernel_plan()
    m.close()
    work = (t[:, :, 1] - t[:, :, 0]).max(axis=0) / 1e3          # us, slowest traced CTA
    wait = (t[:, :, 2] - t[:, :, 1]).min(axis=0) / 1e3          # us, the CTA that arrived last waits least
    total = (t[:, -1, 2] - t[:, 0, 0]).max() / 1e3
    out = {"token_us": round(float(total), 1), "phases": len(kinds), "per_kind": {}}
    for k, name in KINDS.items():
        sel = [i for i, kk in enumerate(kinds) if kk == k]
        if sel:
            out["per_kind"][name] = {"count": len(sel), "work_us_total": round(float(work[sel].sum()), 1),
                                     "barrier_wait_us_total": round(float(wait[sel].sum()), 1),
                                     "work_us_mean": round(float(work[sel].mean()), 2), "wait_us_mean": round(float(wait[sel].mean()), 2)}
    # streaming rate of each GEMV phase type: algorithmic weight bytes of the phase / its work time (slowest traced CTA)
    from ntransformer_b200.dtypes import dtype_row_size
    from ntransformer_b200.model_spec import tensor_table

    nbytes = {name: rows * dtype_row_size(dt, cols) for name, dt, rows, cols in tensor_table(cfg, args.mix)}
    gemv_idx = [i for i, kk in enumerate(kinds) if kk == 2]
    groups = {"q/k/v": ("attn_q", "attn_k", "attn_v"), "o": ("attn_output",), "gate/up": ("ffn_gate", "ffn_up"), "down": ("ffn_down",)}
    rates = {k: [] for k in groups}
    for layer in range(cfg.n_layers):
        for j, (gname, tensors) in enumerate(groups.items()):
            i = gemv_idx[4 * layer + j]
            b = sum(nbytes[f"blk.{layer}.{t}.weight"] for t in tensors)
            rates[gname].append(b / (work[i] * 1e-6) / 1e9)
    out["gemv_GBps"] = {k: {"mean": round(float(np.mean(v)), 1), "min": round(float(np.min(v)), 1), "max": round(float(np.max(v)), 1)} for k, v in rates.items()}
    head_i = gemv_idx[-1]
    out["gemv_GBps"]["lm_head"] = round(nbytes["output.weight"] / (work[head_i] * 1e-6) / 1e9, 1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
