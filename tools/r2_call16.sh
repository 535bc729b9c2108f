#!/usr/bin/env bash
# round 2, GPU call 16: narrow-shard GEMVs as two co-resident 6-warp CTAs per SM (the next launch's weight prefetch overlaps the running one)
mkdir -p gpurun_out
for cfg in "default" "NT_B200_GEMV_WARPS=6 NT_B200_GEMV_STAGES=2" "NT_B200_GEMV_WARPS=6" "NT_B200_GEMV_STAGES=2"; do
  for g in 8 1; do
    if [ "$cfg" = "default" ]; then out=$(timeout 300 python tools/prof_decode.py --layers 80 --shard-of $g 2>/dev/null); else out=$(env $cfg timeout 300 python tools/prof_decode.py --layers 80 --shard-of $g 2>/dev/null); fi
    echo "$cfg shard-of $g: $out" | tee -a gpurun_out/c16_shard_variants.txt
  done
done
