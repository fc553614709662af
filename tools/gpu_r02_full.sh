#!/bin/bash
# full GPU test suite + bench on the current tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_full_pytest.txt 2>&1
tail -4 gpurun_out/r02_full_pytest.txt
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_full_bench.json 2> gpurun_out/r02_full_bench.err
cat gpurun_out/r02_full_bench.json | cut -c1-1500
B200XF_LN_V1=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_full_bench_v1.json 2>> gpurun_out/r02_full_bench.err
cat gpurun_out/r02_full_bench_v1.json | cut -c1-400
