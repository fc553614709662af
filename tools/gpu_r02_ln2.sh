#!/bin/bash
# gemm_ln v2 against v1 and the GEMM + LayerNorm pair; then the per-warp clock trace (needs transformer_b200/libb200xf_trace.so:
# gemm_ln.cu compiled with -DBX_LN2_TRACE and linked with the other objects of csrc/build, see tools/ln_trace.py)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "layernorm_fused" > gpurun_out/r02_ln2_test.txt 2>&1
tail -3 gpurun_out/r02_ln2_test.txt
echo "== v2" > gpurun_out/r02_ln2_bench.txt
B200XF_LN_DEBUG=1 timeout 300 python tools/ln_bench.py >> gpurun_out/r02_ln2_bench.txt 2>&1
echo "== v1" >> gpurun_out/r02_ln2_bench.txt
B200XF_LN_V1=1 timeout 300 python tools/ln_bench.py >> gpurun_out/r02_ln2_bench.txt 2>&1
cat gpurun_out/r02_ln2_bench.txt
rm -f gpurun_out/r02_ln2_trace.txt
for cfg in "512 0.0" "512 0.1" "2048 0.0"; do
  B200XF_LIB_PATH=$PWD/transformer_b200/libb200xf_trace.so timeout 120 python tools/ln_trace.py $cfg >> gpurun_out/r02_ln2_trace.txt 2>&1
done
