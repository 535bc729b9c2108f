#!/usr/bin/env bash
# round 2, GPU call 2b: ncu --set full of the persistent kernel and of one layer's worth of graph-path kernels (70B shapes, 8 layers).
# gpurun_out/ is capped at 64 MiB: raw pages are exported to CSV on the box, only the small reports travel.
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none"
timeout 900 $NCU --import-source on -k regex:decode_step_kernel -s 3 -c 1 -f -o gpurun_out/r02_mega_full python tools/prof_decode.py --layers 8 --mega 1 > gpurun_out/c2_ncu_mega.log 2>&1; echo "mega ncu rc=$?"
timeout 900 $NCU -s 175 -c 26 -f -o /tmp/ncu/r02_graph_layer python tools/prof_decode.py --layers 8 > gpurun_out/c2_ncu_graph.log 2>&1; echo "graph ncu rc=$?"
ncu -i /tmp/ncu/r02_graph_layer.ncu-rep --page raw --csv > gpurun_out/r02_graph_layer_raw.csv 2> gpurun_out/c2_export.log
ncu -i /tmp/ncu/r02_graph_layer.ncu-rep --page details --csv > gpurun_out/r02_graph_layer_details.csv 2>> gpurun_out/c2_export.log
timeout 600 $NCU --import-source on -k regex:"decode_kernel|decode_combine" -s 8 -c 2 -f -o gpurun_out/r02_attn_ctx2048 python tools/prof_decode.py --model 8b --layers 4 --ctx 2048 > gpurun_out/c2_ncu_attn.log 2>&1; echo "attn ncu rc=$?"
# the fused gate+up and the Q4_K down launch with sources, for the per-instruction stall view
timeout 600 $NCU --import-source on -k regex:gemv_kq_kernel -s 40 -c 5 -f -o gpurun_out/r02_gemv5 python tools/prof_decode.py --layers 8 > gpurun_out/c2_ncu_gemv5.log 2>&1; echo "gemv5 ncu rc=$?"
du -sh gpurun_out; ls -la gpurun_out
