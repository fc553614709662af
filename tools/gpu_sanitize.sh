#!/bin/bash
# compute-sanitizer memcheck over the kernels that changed late in round 2 (small shapes of the parity tests)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "layernorm_fused or output_column_sums or xent or residual_ln or ln_bwd" > gpurun_out/r02_sanitizer_memcheck.txt 2>&1
echo "memcheck rc $?"; tail -6 gpurun_out/r02_sanitizer_memcheck.txt
