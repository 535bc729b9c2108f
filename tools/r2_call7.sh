#!/usr/bin/env bash
# round 2, GPU call 7: where our error comes from at real sizes (per-token vs batched prompt, F64 oracle beside), tail split, drop-in, loader
mkdir -p gpurun_out
for L in 4 32; do
  timeout 900 python bench.py --check --workload llama3-8b-q4_k_m-decode --layers $L --steps 12 --oracle-steps 4 --per-token-prompt > gpurun_out/c7_check_8b_${L}l_pt.json 2> gpurun_out/c7_check_${L}.err; cat gpurun_out/c7_check_8b_${L}l_pt.json
done
NT_B200_FUSE=0 timeout 900 python bench.py --check --workload llama3-8b-q4_k_m-decode --layers 4 --steps 12 --oracle-steps 4 --per-token-prompt > gpurun_out/c7_check_8b_4l_pt_unfused.json 2>/dev/null; cat gpurun_out/c7_check_8b_4l_pt_unfused.json
timeout 900 python bench.py --check --workload llama3-8b-q8_0-decode --layers 4 --steps 12 --oracle-steps 4 --per-token-prompt > gpurun_out/c7_check_8b_q8_4l_pt.json 2>/dev/null; cat gpurun_out/c7_check_8b_q8_4l_pt.json
timeout 900 python -m pytest tests/test_chain_gpu.py -q -x -k "tail or prologue" > gpurun_out/c7_tail.log 2>&1; echo "tail rc=$?"; tail -3 gpurun_out/c7_tail.log
timeout 600 python -m pytest tests/test_dropin_gpu.py -q > gpurun_out/c7_dropin.log 2>&1; echo "dropin rc=$?"; tail -3 gpurun_out/c7_dropin.log
for t in 8 16; do timeout 900 python tools/bench_load.py --workload llama3-8b-q4_k_m-decode --threads $t > gpurun_out/c7_load_8b_t$t.json 2> gpurun_out/c7_load.err; cat gpurun_out/c7_load_8b_t$t.json; done
for v in 1 0; do NT_B200_TAIL_SPLIT=$v timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/c7_bench_tail$v.json 2> gpurun_out/c7_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/c7_bench_tail$v.json").read().strip().splitlines()[-1]); print("tail_split=$v", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["step_frac"], d["roofline"]["avg_launch_us"])
PY
done
