#!/bin/bash
# round 2, final evidence run on one B200: full GPU test suite, smoke(), default bench, ncu captures of the dominant kernels, launch list
set -o pipefail
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_final_pytest_gpu.txt; tail -2 gpurun_out/r02_final_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.txt 2>&1; tail -1 gpurun_out/r02_final_smoke.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; head -c 400 gpurun_out/r02_final_bench.json; echo
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_kernel -s 3 -c 1 -o gpurun_out/r02_final_gemm_ffn_wgrad -f python tools/gemm_ffn_wgrad.py > gpurun_out/r02_final_ncu1.log 2>&1
timeout 300 $NCU -k regex:attn_fwd -s 2 -c 1 -o gpurun_out/r02_final_attn_fwd -f python tools/attn_bench.py quick > gpurun_out/r02_final_ncu2.log 2>&1
timeout 300 $NCU -k regex:attn_bwd -s 2 -c 1 -o gpurun_out/r02_final_attn_bwd -f python tools/attn_bench.py quick > gpurun_out/r02_final_ncu3.log 2>&1
timeout 300 $NCU -k regex:gemm_ln -s 3 -c 1 -o gpurun_out/r02_final_gemm_ln -f python tools/ln_bench.py > gpurun_out/r02_final_ncu4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_launches_bench.log 2>&1
python tools/attn_bench.py > gpurun_out/r02_final_attn_bench.txt 2>&1
ls gpurun_out/r02_final_* | wc -l
