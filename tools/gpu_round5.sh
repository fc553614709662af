#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 > gpurun_out/pytest_gpu_full.log 2>&1
tail -8 gpurun_out/pytest_gpu_full.log
grep "worst rel\|weight drift\|cosine" gpurun_out/pytest_gpu_full.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: print(l[:300]); continue
    print({k:j[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); r=j['roofline']; print({k:r[k] for k in ('achieved','frac','share_of_step','launches_per_step')}, r['attention_fwd'], r['attention_bwd'])
"


