#!/bin/bash
python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "xent" 2>&1 | tail -4
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_y_bench.json 2> gpurun_out/r02_y_bench.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02_y_bench.json").read().strip().splitlines()[-1])
print(round(j["ms_per_step"],3), "ms", round(j["value"]), "tok/s; launches/step", j["gpu_launches"]/j["steps"], "gemm frac", round(j["roofline"]["frac"],3), "clk", j["clocks"])
PY
