#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 -k gemm > gpurun_out/pytest_gemm.log 2>&1
tail -15 gpurun_out/pytest_gemm.log
echo "--- cta1"; timeout 300 python tools/gemm_bench.py 1 2>&1 | tail -12
echo "--- ctapair"; timeout 300 python tools/gemm_bench.py 2 2>&1 | tail -12
