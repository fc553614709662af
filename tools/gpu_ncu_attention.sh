#!/bin/bash
# ncu --set full of the attention kernels at transformer-base shapes (3rd launch of each = non-causal, warm)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:attn_bwd -s 2 -c 1 -o gpurun_out/prof_attn_bwd_final -f python tools/attn_bench.py quick > gpurun_out/prof_attn_bwd_final.log 2>&1
timeout 300 $NCU -k regex:attn_fwd -s 2 -c 1 -o gpurun_out/prof_attn_fwd_final -f python tools/attn_bench.py quick > gpurun_out/prof_attn_fwd_final.log 2>&1
ls -la gpurun_out/prof_attn_*_final.ncu-rep
