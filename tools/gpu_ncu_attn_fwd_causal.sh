#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:attn_fwd -s 5 -c 1 -o gpurun_out/prof_attn_fwd_causal2 -f python tools/attn_bench.py quick > gpurun_out/prof_attn_fwd_causal.log 2>&1
timeout 600 $NCU -k regex:attn_fwd -s 2 -c 1 -o gpurun_out/prof_attn_fwd_full2 -f python tools/attn_bench.py quick > gpurun_out/prof_attn_fwd_full.log 2>&1
