#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 300 -k gemm > gpurun_out/pytest_gemm.log 2>&1
tail -5 gpurun_out/pytest_gemm.log
python - <<'PY'
import torch, ctypes
print("SMs", torch.cuda.get_device_properties(0).multi_processor_count)
PY
echo "--- ctapair"; timeout 300 python tools/gemm_bench.py 2 2>&1 | tail -12
