#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 > gpurun_out/pytest_gpu_full.log 2>&1
tail -30 gpurun_out/pytest_gpu_full.log > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err >> gpurun_out/bench.log
grep -A 12 "gradient rel-L2" gpurun_out/pytest_gpu_full.log | head -40
grep "worst rel" gpurun_out/pytest_gpu_full.log
cat gpurun_out/pytest_gpu.log | tail -12; tail -2 gpurun_out/smoke.log; cat gpurun_out/gemm_bench.log; cat gpurun_out/bench.log
