#!/usr/bin/env bash
# round 2, GPU call 12 (8 GPUs): TP-8 with the activation quantiser in the down GEMV's prologue; one-launch attention on/off
mkdir -p gpurun_out
for f in 1 3; do
NT_B200_FUSE=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$f bench.py --gpus 8 --steps 64 --warmup 8 --no-configs3 > gpurun_out/c12_bench_tp8_fuse$f.json 2> gpurun_out/c12_bench_tp8_fuse$f.err; echo "bench8 fuse=$f rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c12_bench_tp8_fuse$f.json").read().strip().splitlines()[-1]); print("tp8 fuse=$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["path"]["launches_per_step"])
except Exception as e: print("tp8 fuse=$f failed", e); print(open("gpurun_out/c12_bench_tp8_fuse$f.err").read()[-2000:])
PY
done
