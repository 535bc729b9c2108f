#!/usr/bin/env bash
# round 2, GPU call 2: ncu --set full of the persistent kernel and of one layer's worth of graph-path kernels (70B shapes, 8 layers)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:decode_step_kernel -s 3 -c 1 -f -o gpurun_out/r02_mega_full python tools/prof_decode.py --layers 8 --mega 1 > gpurun_out/c2_ncu_mega.log 2>&1; echo "mega ncu rc=$?"
timeout 900 $NCU -s 175 -c 30 -f -o gpurun_out/r02_graph_layer python tools/prof_decode.py --layers 8 > gpurun_out/c2_ncu_graph.log 2>&1; echo "graph ncu rc=$?"
timeout 600 $NCU -k regex:"decode_kernel|decode_combine" -s 8 -c 2 -f -o gpurun_out/r02_attn_ctx2048 python tools/prof_decode.py --model 8b --layers 4 --ctx 2048 > gpurun_out/c2_ncu_attn.log 2>&1; echo "attn ncu rc=$?"
tail -3 gpurun_out/c2_ncu_mega.log gpurun_out/c2_ncu_graph.log gpurun_out/c2_ncu_attn.log
ls -la gpurun_out/*.ncu-rep
