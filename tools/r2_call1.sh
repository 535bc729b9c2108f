#!/usr/bin/env bash
# round 2, GPU call 1: first hardware run of everything that was gated in round 1
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi -L > gpurun_out/c1_gpus.txt 2>&1
NT_B200_TEST_MEGA=1 timeout 600 python -m pytest tests/test_mega_gpu.py -x -q -k "compat or uncovered" -s > gpurun_out/c1_mega_compat.log 2>&1; echo "compat rc=$?"
NT_B200_TEST_MEGA=1 timeout 900 python -m pytest tests/test_mega_gpu.py -q -k "not tensor_parallel" > gpurun_out/c1_mega_all.log 2>&1; echo "mega all rc=$?"
NT_B200_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_sample_gpu.py tests/test_q4_0_tma_gpu.py -q > gpurun_out/c1_unverified.log 2>&1; echo "unverified rc=$?"
for v in 0 1; do NT_B200_MEGAKERNEL=$v timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/c1_bench_mega_$v.json 2> gpurun_out/c1_bench_mega_$v.err; tail -1 gpurun_out/c1_bench_mega_$v.json; done
for f in 3 7 31 63; do NT_B200_MEGAKERNEL=1 NT_B200_MEGA_FUSE=$f timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/c1_bench_mega_fuse$f.json 2> gpurun_out/c1_bench_mega_fuse$f.err; tail -1 gpurun_out/c1_bench_mega_fuse$f.json; done
for f in 0 63; do timeout 300 python tools/mega_trace.py --model 70b --fuse $f > gpurun_out/c1_mega_trace_fuse$f.json 2> gpurun_out/c1_mega_trace_fuse$f.err; done
tail -5 gpurun_out/c1_mega_compat.log gpurun_out/c1_mega_all.log gpurun_out/c1_unverified.log
