#!/usr/bin/env bash
# round 2, GPU call 15: the driver's round-end sequence (pytest -m gpu, smoke, bench) + the other BASELINE configs + remaining ncu captures
mkdir -p gpurun_out
for g in 8 1; do timeout 300 python tools/prof_decode.py --layers 80 --shard-of $g > gpurun_out/c15_shard_$g.json 2>/dev/null; cat gpurun_out/c15_shard_$g.json; done
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c15_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -6 gpurun_out/c15_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c15_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c15_smoke.log
timeout 900 python bench.py > gpurun_out/c15_bench_70b.json 2> gpurun_out/c15_bench_70b.err; echo "bench rc=$?"; tail -c 900 gpurun_out/c15_bench_70b.json
for w in llama3-8b-q8_0-decode llama3-8b-q4_k_m-decode-ctx2048 llama3-8b-q4_k_m-decode llama3-8b-f16-prefill-4096 llama3-8b-q4_k_m-prefill-4096; do
  timeout 600 python bench.py --workload $w --steps 64 --warmup 4 --no-cpu-baseline > gpurun_out/c15_bench_$w.json 2> gpurun_out/c15_bench_$w.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c15_bench_$w.json").read().strip().splitlines()[-1]); print("$w", d["metric"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"].get("step_frac"), d["roofline"].get("frac"))
except Exception as e: print("$w failed", e); print(open("gpurun_out/c15_bench_$w.err").read()[-1200:])
PY
done
for w in llama3-8b-q8_0-decode llama3-8b-q4_k_m-decode; do timeout 600 python bench.py --impl reference --workload $w --steps 32 --warmup 4 > gpurun_out/c15_ref_$w.json 2>/dev/null; tail -c 400 gpurun_out/c15_ref_$w.json; echo; done
NCU="ncu --set full --clock-control none"
timeout 600 $NCU -k regex:"decode_kernel|decode_combine|rope_kv|rmsnorm_xq|embed" -s 6 -c 6 -f -o /tmp/r02_attn python tools/prof_decode.py --model 8b --layers 4 --ctx 2048 > gpurun_out/c15_ncu_attn.log 2>&1
ncu -i /tmp/r02_attn.ncu-rep --page raw --csv > gpurun_out/r02_attn_ctx2048_raw.csv 2>> gpurun_out/c15_ncu_attn.log
timeout 600 $NCU -k regex:"gemm_f16_tc|prefill_mma|rmsnorm_split|split_f32|dequant_split" -s 12 -c 12 -f -o /tmp/r02_prefill python tools/prof_prefill.py --layers 2 --mix F16 --tokens 2048 > gpurun_out/c15_ncu_prefill.log 2>&1
ncu -i /tmp/r02_prefill.ncu-rep --page raw --csv > gpurun_out/r02_prefill_f16_raw.csv 2>> gpurun_out/c15_ncu_prefill.log
timeout 600 $NCU -k regex:"gemm_f16_tc|dequant_split" -s 14 -c 8 -f -o /tmp/r02_prefill_q python tools/prof_prefill.py --layers 2 --mix Q4_K_M --tokens 2048 > gpurun_out/c15_ncu_prefill_q.log 2>&1
ncu -i /tmp/r02_prefill_q.ncu-rep --page raw --csv > gpurun_out/r02_prefill_q4km_raw.csv 2>> gpurun_out/c15_ncu_prefill_q.log
ls -la gpurun_out | tail -5
