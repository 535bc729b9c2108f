#!/usr/bin/env bash
# round 2, GPU call 3: first run of the fused launch chain (GEMV epilogue fusions + one-launch attention)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_gpu.py -x -q > gpurun_out/c3_chain.log 2>&1; echo "chain rc=$?"; tail -15 gpurun_out/c3_chain.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/c3_model.log 2>&1; echo "model+kernels rc=$?"; tail -5 gpurun_out/c3_model.log
for f in 3 0 1 2; do NT_B200_FUSE=$f timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/c3_bench_fuse$f.json 2> gpurun_out/c3_bench_fuse$f.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c3_bench_fuse$f.json").read().strip().splitlines()[-1]); print("fuse=$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["step_frac"], d.get("gpu_launches"))
except Exception as e: print("fuse=$f failed", e)
PY
done
