#!/bin/bash
mkdir -p gpurun_out
B200XF_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r1.csv | head -26
