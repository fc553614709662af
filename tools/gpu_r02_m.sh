#!/bin/bash
python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "fused" 2>&1 | tail -15 > gpurun_out/r02_m_ln_tests.txt
tail -3 gpurun_out/r02_m_ln_tests.txt
python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r02_m_pytest_gpu.txt
tail -3 gpurun_out/r02_m_pytest_gpu.txt
B200XF_FUSE_LN=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_m_bench_nofuse.json 2>gpurun_out/r02_m_bench_nofuse.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_m_bench_fuse.json 2>gpurun_out/r02_m_bench_fuse.err
B200XF_CONCURRENCY=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_m_bench_fuse_noconc.json 2>gpurun_out/r02_m_bench_fuse_noconc.err
python - <<'PY'
import json
for n in ("nofuse","fuse","fuse_noconc"):
    try:
        j=json.loads(open(f"gpurun_out/r02_m_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["ms_per_step"],3), "ms", round(j["value"]), "tok/s; gemm frac", round(j["roofline"]["frac"],3), "ffn", round(j["roofline"]["ffn_gemm"]["frac"],3), "loss/tok", round(j["loss_per_token"],4))
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/r02_m_bench_{n}.err").read()[-800:])
PY
