#!/usr/bin/env bash
# round 2, GPU call 4: the short launch chain (norm + quantiser in the GEMV prologues, one-launch attention)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_gpu.py -q > gpurun_out/c4_chain.log 2>&1; echo "chain rc=$?"; tail -12 gpurun_out/c4_chain.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/c4_model.log 2>&1; echo "model+kernels rc=$?"; tail -5 gpurun_out/c4_model.log
for f in 3 0 1 2; do NT_B200_FUSE=$f timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/c4_bench_fuse$f.json 2> gpurun_out/c4_bench_fuse$f.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c4_bench_fuse$f.json").read().strip().splitlines()[-1]); print("fuse=$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["step_frac"], d.get("gpu_launches"), d["roofline"]["avg_launch_us"])
except Exception as e: print("fuse=$f failed", e)
PY
done
NCU="ncu --set full --clock-control none"
timeout 600 $NCU -k regex:"gemv_kq|decode_fused|quantize_x|embed" -s 30 -c 14 -f -o /tmp/r02_chain_layer python tools/prof_decode.py --layers 8 > gpurun_out/c4_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02_chain_layer.ncu-rep --page raw --csv > gpurun_out/r02_chain_layer_raw.csv 2>> gpurun_out/c4_ncu.log
