#!/bin/bash
# GPU session 2: all gpu tests (kernels + model vs oracle), smoke, first bench, ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --timeout 600 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err >> gpurun_out/bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -60 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.log
