#!/bin/bash
# round 2, second (late) evidence run on one B200: full GPU test suite, smoke(), default bench, ncu captures of the kernels that changed
# since r02_final_* (gemm_ln v2, GEMM with the second epilogue warpgroup, LayerNorm backward), launch list of one step
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_final2_pytest_gpu.txt; tail -2 gpurun_out/r02_final2_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final2_smoke.txt 2>&1; tail -1 gpurun_out/r02_final2_smoke.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_final2_bench.json 2> gpurun_out/r02_final2_bench.err; head -c 400 gpurun_out/r02_final2_bench.json; echo
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_ln2 -s 3 -c 1 -o gpurun_out/r02_final2_gemm_ln2 -f python tools/ln_bench.py > gpurun_out/r02_final2_ncu1.log 2>&1
timeout 300 $NCU -k regex:gemm_kernel -s 3 -c 1 -o gpurun_out/r02_final2_gemm_ffn_dense0 -f python tools/gemm_bench.py > gpurun_out/r02_final2_ncu2.log 2>&1
timeout 300 $NCU -k regex:residual_ln_bwd -s 2 -c 1 -o gpurun_out/r02_final2_ln_bwd -f python tools/lnbwd_bench.py > gpurun_out/r02_final2_ncu3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_final2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final2_launches_bench.log 2>&1
python tools/ln_bench.py > gpurun_out/r02_final2_ln_bench.txt 2>&1
python tools/gemm_bench.py > gpurun_out/r02_final2_gemm_bench.txt 2>&1
python tools/lnbwd_bench.py > gpurun_out/r02_final2_lnbwd_bench.txt 2>&1
ls gpurun_out/r02_final2_* | wc -l
