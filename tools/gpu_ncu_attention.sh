#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:attn_fwd -s 2 -c 1 -o gpurun_out/prof_attn_fwd2 -f python tools/attn_bench.py quick > gpurun_out/prof_attn_fwd2.log 2>&1
timeout 600 $NCU -k regex:attn_bwd -s 2 -c 1 -o gpurun_out/prof_attn_bwd2 -f python tools/attn_bench.py quick > gpurun_out/prof_attn_bwd2.log 2>&1
timeout 600 $NCU -k regex:residual_ln_bwd -s 2 -c 1 -o gpurun_out/prof_ln_bwd -f python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "layernorm" > gpurun_out/prof_ln.log 2>&1
ls -la gpurun_out/*2.ncu-rep gpurun_out/prof_ln_bwd.ncu-rep
