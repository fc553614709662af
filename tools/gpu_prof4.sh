#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
export B200XF_CUDA_GRAPH=0
timeout 600 $NCU -k regex:residual_ln_bwd -s 40 -c 1 -o gpurun_out/prof_ln_bwd -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_ln.log 2>&1
timeout 600 $NCU -k regex:xent -s 2 -c 1 -o gpurun_out/prof_xent -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_xent.log 2>&1
timeout 600 $NCU -k regex:gemm_kernel -s 300 -c 4 -o gpurun_out/prof_gemm_instep -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_gemm_instep.log 2>&1
ls -la gpurun_out/prof_ln_bwd.ncu-rep gpurun_out/prof_xent.ncu-rep gpurun_out/prof_gemm_instep.ncu-rep
