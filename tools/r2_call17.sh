#!/usr/bin/env bash
# round 2, GPU call 17: full GPU suite + smoke after the fixes, narrow-shard GEMV variants, prefill / long-context benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c17_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -6 gpurun_out/c17_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c17_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c17_smoke.log
bash tools/r2_call16.sh
for w in llama3-8b-q4_k_m-decode-ctx2048 llama3-8b-f16-prefill-4096 llama3-8b-q4_k_m-prefill-4096; do
  timeout 600 python bench.py --workload $w --steps 64 --warmup 4 --no-cpu-baseline > gpurun_out/c17_bench_$w.json 2> gpurun_out/c17_bench_$w.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c17_bench_$w.json").read().strip().splitlines()[-1]); print("$w", d["metric"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"].get("step_frac"), d["roofline"].get("frac"))
except Exception as e: print("$w failed", e); print(open("gpurun_out/c17_bench_$w.err").read()[-1200:])
PY
done
timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/c17_bench_70b.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/c17_bench_70b.json').read().strip().splitlines()[-1]); print('70b', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_frac'])"
