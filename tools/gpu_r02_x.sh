#!/bin/bash
# 1 GPU: C5 sweep, ncu of the attention backward, launch list of one step
python bench.py --sweep > gpurun_out/r02_x_sweep.jsonl 2> gpurun_out/r02_x_sweep.err; tail -c 600 gpurun_out/r02_x_sweep.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 2 -c 1 -o gpurun_out/r02_x_attn_bwd -f python tools/attn_bench.py quick > gpurun_out/r02_x_ncu_bwd.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_x_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_x_launches_bench.log 2>&1
ls -la gpurun_out/r02_x_*
