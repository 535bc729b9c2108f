#!/usr/bin/env bash
# round 2, GPU call 11: quarter-block GEMV kernel (16 warps) vs half-block: tests + bench A/B + ncu of one layer
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_chain_gpu.py -q -x > gpurun_out/c11_kernels.log 2>&1; echo "kernels rc=$?"; tail -6 gpurun_out/c11_kernels.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_prefill_gemm_gpu.py -q -x -k "not parity_at" > gpurun_out/c11_model.log 2>&1; echo "model rc=$?"; tail -4 gpurun_out/c11_model.log
for v in 1 0; do NT_B200_GEMV_QB=$v timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/c11_bench_qb$v.json 2> gpurun_out/c11_bench_qb$v.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c11_bench_qb$v.json").read().strip().splitlines()[-1]); print("qb=$v", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["step_frac"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"])
except Exception as e: print("qb=$v failed", e); print(open("gpurun_out/c11_bench_qb$v.err").read()[-1500:])
PY
done
NCU="ncu --set full --clock-control none"
timeout 600 $NCU -k regex:"gemv_kq" -s 30 -c 10 -f -o /tmp/r02_qb_layer python tools/prof_decode.py --layers 8 > gpurun_out/c11_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02_qb_layer.ncu-rep --page raw --csv > gpurun_out/r02_qb_layer_raw.csv 2>> gpurun_out/c11_ncu.log
