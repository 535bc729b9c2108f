#!/usr/bin/env bash
# round 2, GPU call 25: two-launch attention on the 8B shapes at short and long context (is it worth a per-shape default?)
mkdir -p gpurun_out
for f in 1 5; do for w in llama3-8b-q4_k_m-decode-ctx2048 llama3-8b-q4_k_m-decode llama3-8b-q8_0-decode; do
  NT_B200_FUSE=$f timeout 300 python bench.py --workload $w --steps 96 --warmup 6 --no-cpu-baseline > gpurun_out/c25_${w}_fuse$f.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/c25_${w}_fuse$f.json').read().strip().splitlines()[-1]); print('FUSE=$f $w', d['value'], d['ms_per_step'], d['e2e']['value'])"
done; done
