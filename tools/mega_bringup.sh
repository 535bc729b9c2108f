#!/usr/bin/env bash
# Round-2 bring-up of the persistent decode kernel on a real B200: run from the repo root, one step per
# gpurun call so a failure costs little.  Usage: tools/mega_bringup.sh <step>   (steps 1..6)
set -euo pipefail
G=/usr/local/graft/bin/gpurun
case "${1:-}" in
  1) $G --timeout 900 -- 'NT_B200_TEST_MEGA=1 timeout 800 python -m pytest tests/test_mega_gpu.py -x -q -k "compat or uncovered" -s 2>&1 | tail -40' ;;
  2) $G --timeout 1200 -- 'NT_B200_TEST_MEGA=1 timeout 1100 python -m pytest tests/test_mega_gpu.py -x -q -k "not tensor_parallel" -s 2>&1 | tail -40' ;;
  3) $G --timeout 1200 -- 'mkdir -p gpurun_out; for v in 0 1; do NT_B200_MEGAKERNEL=$v python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench_mega_$v.json 2> gpurun_out/bench_mega_$v.err || true; tail -1 gpurun_out/bench_mega_$v.json; done; for f in 7 31 63; do NT_B200_MEGAKERNEL=1 NT_B200_MEGA_FUSE=$f python bench.py --steps 64 --warmup 8 --no-cpu-baseline | tee gpurun_out/bench_mega_fuse$f.json; done' ;;
  4) $G --timeout 1500 -- 'mkdir -p gpurun_out; NT_B200_MEGAKERNEL=1 timeout 1400 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -c 1 -o gpurun_out/mega_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1; ls -la gpurun_out' ;;
  5) $G --gpus 2 --timeout 1200 -- 'NT_B200_TEST_MEGA=1 timeout 1100 python -m pytest tests/test_mega_gpu.py -x -q -k tensor_parallel -s 2>&1 | tail -40' ;;
  6) $G --gpus 8 --timeout 1500 -- 'mkdir -p gpurun_out; for v in 0 1; do NT_B200_MEGAKERNEL=$v NT_B200_MEGA_FUSE=${FUSE:-51} python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 48 --warmup 8 > gpurun_out/bench_tp8_mega_$v.json 2> gpurun_out/bench_tp8_mega_$v.err || true; tail -1 gpurun_out/bench_tp8_mega_$v.json; done' ;;
  7) $G --timeout 900 -- 'mkdir -p gpurun_out; for f in 0 3 7 31 63; do python tools/mega_trace.py --model 70b --fuse $f > gpurun_out/mega_trace_fuse$f.json 2> gpurun_out/mega_trace_fuse$f.err || true; cat gpurun_out/mega_trace_fuse$f.json; done' ;;
  8) $G --timeout 900 -- 'NT_B200_TEST_UNVERIFIED=1 timeout 800 python -m pytest tests/test_sample_gpu.py -x -q -s 2>&1 | tail -20' ;;
  9) $G --timeout 1200 -- 'timeout 1100 python tools/mega_bisect.py --model mid --mix Q4_K_M --fuse ${FUSE:-0} 2>&1 | tail -40' ;;
  *) echo "usage: $0 <1..9>  (1 compat parity, 2 all single-GPU mega tests, 3 bench graph vs mega, 4 ncu of the kernel, 5 TP-2 exchange test, 6 TP-8 bench, 7 phase timeline, 8 GPU sampler, 9 per-phase GPU-vs-emulator bisect)"; exit 2 ;;
esac
