#!/usr/bin/env bash
# round 2, GPU call 24: the reference arm on the metric's config (70B Q4_K_M), same build of bench.py
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 32 --warmup 4 > gpurun_out/c24_ref_70b.json 2> gpurun_out/c24_ref_70b.err; echo "rc=$?"; tail -c 1200 gpurun_out/c24_ref_70b.json
