#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:residual_ln_fwd -s 3 -c 1 -o gpurun_out/prof_ln_fwd_bf16 -f python tools/ew_bench.py > gpurun_out/prof_ln1.log 2>&1
timeout 600 $NCU -k regex:residual_ln_bwd -s 3 -c 1 -o gpurun_out/prof_ln_bwd_bf16 -f python tools/ew_bench.py > gpurun_out/prof_ln2.log 2>&1
timeout 600 $NCU -k regex:xent_kernel -s 2 -c 1 -o gpurun_out/prof_xent2 -f python tools/ew_bench.py > gpurun_out/prof_ln3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
