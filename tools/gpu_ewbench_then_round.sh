#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ew_bench.py > gpurun_out/ew_bench.log 2>&1; cat gpurun_out/ew_bench.log | tail -8
bash tools/gpu_tests_smoke_bench.sh
