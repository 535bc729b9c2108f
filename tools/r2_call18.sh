#!/usr/bin/env bash
# round 2, GPU call 18 (8 GPUs): the final tensor-parallel series with the last build: TP-8 (+ configs3 Q6_K), TP-4, TP-2
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 64 --warmup 8 > gpurun_out/c18_bench_tp8.json 2> gpurun_out/c18_bench_tp8.err; echo "bench8 rc=$?"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 64 --warmup 8 > gpurun_out/c18_bench_tp4.json 2> gpurun_out/c18_bench_tp4.err; echo "bench4 rc=$?"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/c18_bench_tp2.json 2> gpurun_out/c18_bench_tp2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for n in (8,4,2):
    try:
        d=json.loads(open(f"gpurun_out/c18_bench_tp{n}.json").read().strip().splitlines()[-1]); print("tp",n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["tp"]["exchange"][:20]); c=d.get("configs3"); 
        if c: print("  configs3", c["value"], c["ms_per_step"], c["e2e"]["value"], c["roofline"]["step_frac"])
    except Exception as e: print("tp",n,"failed", e); print(open(f"gpurun_out/c18_bench_tp{n}.err").read()[-1500:])
PY
