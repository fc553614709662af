#!/bin/bash
# round 2, call D: O_lo residual + dk=128 + lse snap: all GPU tests, BASELINE-shape parity, bench
python -m pytest tests -q -m gpu -x --deselect tests/test_baseline_shapes_gpu.py 2>&1 | tail -15 > gpurun_out/r02_d_pytest_gpu.txt
python -m pytest tests/test_baseline_shapes_gpu.py -q -m gpu -s 2>&1 | grep -v "^$" | grep "logits:\|^loss\|cosine\|^   0\.\|Error\|passed\|failed\| | " > gpurun_out/r02_d_baseline.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_d_bench.json 2>gpurun_out/r02_d_bench.err
tail -4 gpurun_out/r02_d_pytest_gpu.txt; tail -3 gpurun_out/r02_d_baseline.txt; head -c 300 gpurun_out/r02_d_bench.json
