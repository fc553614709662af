#!/bin/bash
# first GPU session: kernel unit tests + GEMM micro-benchmark; everything logged under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -40 > gpurun_out/kernels_pytest.log
echo "pytest exit: $?" >> gpurun_out/kernels_pytest.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 2>&1 | tail -60 > gpurun_out/kernels_pytest_all.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1
tail -30 gpurun_out/kernels_pytest_all.log
cat gpurun_out/gemm_bench.log
