#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/prof_gemm_pair -f python tools/gemm_one.py 2 8192 8192 8192 > gpurun_out/prof_gemm_pair.log 2>&1
timeout 300 $NCU -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/prof_gemm_single -f python tools/gemm_one.py 1 8192 8192 8192 > gpurun_out/prof_gemm_single.log 2>&1
ls -la gpurun_out/prof_gemm_*.ncu-rep
