#!/bin/bash
# round 2, call B: attention-backward precision experiment, full output of the BASELINE-shape tests, side-stream A/B bench
python tools/attn_precision.py > gpurun_out/r02_attn_precision.txt 2>&1
python -m pytest tests/test_baseline_shapes_gpu.py -q -m gpu -s -k "s512 or s128 or big" > gpurun_out/r02_b_baseline_full.txt 2>&1
B200XF_CONCURRENCY=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_b_bench_noconc.json 2>gpurun_out/r02_b_bench_noconc.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_b_bench_conc.json 2>gpurun_out/r02_b_bench_conc.err
python -m pytest tests -q -m gpu -x --deselect tests/test_baseline_shapes_gpu.py 2>&1 | tail -5
