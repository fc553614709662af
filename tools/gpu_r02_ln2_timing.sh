#!/bin/bash
mkdir -p gpurun_out
for cfg in "512 0.0" "512 0.1" "2048 0.0"; do
  echo "== $cfg" >> gpurun_out/r02_ln2_timing3.txt
  B200XF_LIB_PATH=$PWD/transformer_b200/libb200xf_timing.so timeout 120 python tools/ln_one.py $cfg >> gpurun_out/r02_ln2_timing3.txt 2>&1
done
