#!/usr/bin/env bash
# round 2, GPU call 13: what one tensor-parallel rank computes per token without any exchange (single GPU, per-rank shapes)
mkdir -p gpurun_out
for g in 8 4 2 1; do timeout 300 python tools/prof_decode.py --layers 80 --shard-of $g > gpurun_out/c13_shard_$g.json 2> gpurun_out/c13_shard_$g.err; cat gpurun_out/c13_shard_$g.json; done
timeout 300 python tools/prof_decode.py --layers 80 --shard-of 8 --mix Q6_K > gpurun_out/c13_shard_8_q6k.json 2>/dev/null; cat gpurun_out/c13_shard_8_q6k.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv_kq|rope_kv|decode_|quantize_x|rmsnorm|embed" -s 400 -c 60 --csv --log-file gpurun_out/c13_launches_shard8.csv python tools/prof_decode.py --layers 16 --shard-of 8 > /dev/null 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/c13_launches_shard8.csv 2>/dev/null | head -20
