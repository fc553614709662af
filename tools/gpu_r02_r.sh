#!/bin/bash
python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/r02_r_pytest_gpu.txt; tail -2 gpurun_out/r02_r_pytest_gpu.txt
B200XF_FUSE_LN=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_r_bench_nofuse.json 2>gpurun_out/r02_r_bench_nofuse.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_r_bench_fuse.json 2>gpurun_out/r02_r_bench_fuse.err
B200XF_FUSE_LN=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_r_bench_nofuse2.json 2>gpurun_out/r02_r_bench_nofuse2.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_r_bench_fuse2.json 2>gpurun_out/r02_r_bench_fuse2.err
python - <<'PY'
import json
for n in ("nofuse","fuse","nofuse2","fuse2"):
    try:
        j=json.loads(open(f"gpurun_out/r02_r_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["ms_per_step"],3), "ms", round(j["value"]), "tok/s; gemm frac", round(j["roofline"]["frac"],3), "ffn", round(j["roofline"]["ffn_gemm"]["frac"],3), "clk", j["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/r02_r_bench_{n}.err").read()[-800:])
PY
