#!/bin/bash
# round 2, closing run on the final tree: full GPU test suite, smoke(), default bench, launch list of one step
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r02_final3_pytest_gpu.txt; tail -1 gpurun_out/r02_final3_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final3_smoke.txt 2>&1; tail -1 gpurun_out/r02_final3_smoke.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_final3_bench.json 2> gpurun_out/r02_final3_bench.err; head -c 330 gpurun_out/r02_final3_bench.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_final3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final3_launches_bench.log 2>&1
