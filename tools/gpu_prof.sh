#!/bin/bash
# ncu --set full captures of the attention kernels and the FFN GEMM (one GPU, short commands)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:attn_fwd -s 2 -c 2 -o gpurun_out/prof_attn_fwd -f python tools/attn_bench.py quick > gpurun_out/prof_attn_fwd.log 2>&1
timeout 600 $NCU -k regex:attn_bwd -s 2 -c 2 -o gpurun_out/prof_attn_bwd -f python tools/attn_bench.py quick > gpurun_out/prof_attn_bwd.log 2>&1
timeout 600 $NCU -k regex:gemm_kernel -s 3 -c 2 -o gpurun_out/prof_gemm -f python tools/gemm_bench.py > gpurun_out/prof_gemm.log 2>&1
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1
cat gpurun_out/attn_bench.log; ls -la gpurun_out/*.ncu-rep
