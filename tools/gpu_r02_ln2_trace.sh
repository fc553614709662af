#!/bin/bash
mkdir -p gpurun_out
for cfg in "512 0.0" "512 0.1" "2048 0.0"; do
  B200XF_LIB_PATH=$PWD/transformer_b200/libb200xf_trace.so timeout 120 python tools/ln_trace.py $cfg >> gpurun_out/r02_ln2_trace.txt 2>&1
done
