#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 -k attention > gpurun_out/pytest_attn.log 2>&1
tail -12 gpurun_out/pytest_attn.log
timeout 300 python tools/attn_timeline.py 2>&1 | tail -24
timeout 300 python tools/attn_bench.py 2>&1 | tail -10
