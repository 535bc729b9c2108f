#!/usr/bin/env bash
# round 2, GPU call 26 (2 GPUs): sanity of the tensor-parallel path on the final library (bench asserts bit-identical ranks)
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/c26_bench_tp2.json 2> gpurun_out/c26_bench_tp2.err; echo "bench2 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/c26_bench_tp2.json').read().strip().splitlines()[-1]); print('tp2', d['value'], d['ms_per_step'], d['e2e']['value'], d['tp'])" || tail -c 1500 gpurun_out/c26_bench_tp2.err
timeout 200 python -m pytest tests/test_tp_gpu.py -q -x -k "Q4_K_M-2-0 or Q6_K-2-0" 2>&1 | tail -3
