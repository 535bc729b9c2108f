#!/usr/bin/env bash
# round 2, GPU call 8 (2 GPUs): tensor-parallel parity (NVLink peer exchange vs ncclAllReduce vs single GPU) and TP-2 bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c8_gpus.txt
timeout 900 python -m pytest tests/test_tp_gpu.py -q -x > gpurun_out/c8_tp.log 2>&1; echo "tp rc=$?"; tail -15 gpurun_out/c8_tp.log
for nccl in 0 1; do
NT_B200_TP_NCCL=$nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/c8_bench_tp2_nccl$nccl.json 2> gpurun_out/c8_bench_tp2_nccl$nccl.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c8_bench_tp2_nccl$nccl.json").read().strip().splitlines()[-1]); print("nccl=$nccl", d["value"], d["ms_per_step"], d["e2e"]["value"], d["tp"], d["path"])
except Exception as e: print("nccl=$nccl failed", e); print(open("gpurun_out/c8_bench_tp2_nccl$nccl.err").read()[-2000:])
PY
done
