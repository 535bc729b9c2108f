#!/usr/bin/env bash
# round 2, GPU call 20: the driver's round-end sequence on the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c20_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -5 gpurun_out/c20_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c20_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c20_smoke.log
timeout 900 python bench.py > gpurun_out/c20_bench_70b.json 2> gpurun_out/c20_bench_70b.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/c20_bench_70b.json').read().strip().splitlines()[-1]); print('70b', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_frac'], d['roofline']['frac'], d.get('cpu_baseline'))"
