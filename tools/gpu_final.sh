#!/bin/bash
# round-end evidence: full GPU suite, smoke, bench line (graph mode, with cpu_baseline), ncu launch list of one eager step
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu_full.log 2>&1; tail -2 gpurun_out/pytest_gpu_full.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_full.json
bash tools/gpu_launchlist.sh | head -14
