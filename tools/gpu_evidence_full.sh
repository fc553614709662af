#!/bin/bash
# round-end evidence: full GPU suite, smoke, bench (graph), ncu launch list of one eager step, elementwise micro-bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu_full.log 2>&1; tail -3 gpurun_out/pytest_gpu_full.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench_full.json
bash tools/gpu_launchlist.sh
timeout 300 python tools/ew_bench.py > gpurun_out/ew_bench.log 2>&1
timeout 300 python tools/attn_bench.py > gpurun_out/attn_sweep.log 2>&1; tail -4 gpurun_out/attn_sweep.log
