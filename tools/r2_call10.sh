#!/usr/bin/env bash
# round 2, GPU call 10 (8 GPUs): TP-8 bench (Q4_K_M + configs3 Q6_K riding along), TP-4 bench, TP-8/TP-4 parity tests
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/c10_ngpus.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 64 --warmup 8 > gpurun_out/c10_bench_tp8.json 2> gpurun_out/c10_bench_tp8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c10_bench_tp8.json").read().strip().splitlines()[-1]); print("tp8", d["value"], d["ms_per_step"], d["e2e"]["value"], d["tp"], d["path"]); c=d.get("configs3"); print("configs3", c and (c["value"], c["ms_per_step"], c["e2e"]["value"], c["roofline"]["step_frac"]))
except Exception as e: print("tp8 failed", e); print(open("gpurun_out/c10_bench_tp8.err").read()[-2500:])
PY
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 64 --warmup 8 > gpurun_out/c10_bench_tp4.json 2> gpurun_out/c10_bench_tp4.err; echo "bench4 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c10_bench_tp4.json").read().strip().splitlines()[-1]); print("tp4", d["value"], d["ms_per_step"], d["e2e"]["value"], d["tp"])
except Exception as e: print("tp4 failed", e); print(open("gpurun_out/c10_bench_tp4.err").read()[-2500:])
PY
timeout 400 python -m pytest tests/test_tp_gpu.py -q -x -k "8-0 or 4-0" > gpurun_out/c10_tp.log 2>&1; echo "tp tests rc=$?"; tail -8 gpurun_out/c10_tp.log
