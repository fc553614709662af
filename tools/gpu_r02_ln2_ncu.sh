#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_ln2 -s 1 -c 1 -o gpurun_out/r02_ln2 -f python tools/ln_one.py 512 0.1 > gpurun_out/r02_ln2_ncu.log 2>&1
tail -3 gpurun_out/r02_ln2_ncu.log; ls -la gpurun_out/r02_ln2.ncu-rep
