#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu_full.log 2>&1
tail -6 gpurun_out/pytest_gpu_full.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err; cat gpurun_out/bench.log | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l[:300]); continue
    print({k:j.get(k) for k in ('value','ms_per_step','gpu_launches','launch_mode','clocks','e2e')}); r=j['roofline']; print({k:r[k] for k in ('achieved','frac','share_of_step','launches_per_step','measured_in')})
"
B200XF_CUDA_GRAPH=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l); print('eager:', {k:j.get(k) for k in ('value','ms_per_step','launch_mode')})
    except Exception: print(l[:200])
"
