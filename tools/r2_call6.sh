#!/usr/bin/env bash
# round 2, GPU call 6: parity at the 8B bench config with conditioning diagnostics, drop-in link test, loader timing
mkdir -p gpurun_out
timeout 900 python bench.py --check --workload llama3-8b-q4_k_m-decode --steps 24 --oracle-steps 3 > gpurun_out/c6_check_8b_q4km.json 2> gpurun_out/c6_check_8b_q4km.err; echo "check rc=$?"; cat gpurun_out/c6_check_8b_q4km.json
timeout 900 python bench.py --check --workload llama3-8b-q4_k_m-decode --steps 24 --per-token-prompt > gpurun_out/c6_check_8b_q4km_pt.json 2> gpurun_out/c6_check_8b_q4km_pt.err; echo "check pt rc=$?"; cat gpurun_out/c6_check_8b_q4km_pt.json
timeout 900 python bench.py --check --workload llama3-8b-q4_k_m-decode --layers 4 --steps 24 --oracle-steps 3 > gpurun_out/c6_check_8b_4l.json 2> gpurun_out/c6_check_8b_4l.err; echo "check 4l rc=$?"; cat gpurun_out/c6_check_8b_4l.json
timeout 600 python -m pytest tests/test_dropin_gpu.py -q > gpurun_out/c6_dropin.log 2>&1; echo "dropin rc=$?"; tail -5 gpurun_out/c6_dropin.log
timeout 900 python tools/bench_load.py --workload llama3-8b-q4_k_m-decode > gpurun_out/c6_load_8b.json 2> gpurun_out/c6_load_8b.err; echo "load rc=$?"; cat gpurun_out/c6_load_8b.json
timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; tail -c 1500 gpurun_out/c6_bench.json
