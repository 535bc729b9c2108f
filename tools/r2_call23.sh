#!/usr/bin/env bash
# round 2, GPU call 23: launch list (device time per launch, cold-cache, serialised) of the final default decode step, 70B Q4_K_M
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv_|rope_kv|decode_|quantize_x|rmsnorm|embed|argmax|set_step" -s 1500 -c 700 --csv --log-file gpurun_out/r02_launches_70b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c23_bench_under_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_70b.csv | head -24
