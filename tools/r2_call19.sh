#!/usr/bin/env bash
# round 2, GPU call 19: two-launch attention A/B (70B, 8B, TP-8 shard shapes) + the full GPU suite and smoke on the final build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chain_gpu.py -x -q -k "fused_chain or fused_decode" > gpurun_out/c19_chain.log 2>&1; echo "chain rc=$?"; tail -3 gpurun_out/c19_chain.log
for f in 1 5; do
  echo "FUSE=$f shard-of 8: $(NT_B200_FUSE=$f timeout 300 python tools/prof_decode.py --layers 80 --shard-of 8 2>/dev/null)"
  echo "FUSE=$f 8b: $(NT_B200_FUSE=$f timeout 300 python tools/prof_decode.py --model 8b --layers 32 2>/dev/null)"
  NT_B200_FUSE=$f timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/c19_bench_fuse$f.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/c19_bench_fuse$f.json').read().strip().splitlines()[-1]); print('FUSE=$f 70b', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_frac'], d['path']['launches_per_step'])"
done
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c19_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -5 gpurun_out/c19_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c19_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c19_smoke.log
