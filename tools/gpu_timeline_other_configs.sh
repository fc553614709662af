#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/attn_timeline.py 0 2>&1 | grep -A4 "FWD causal=1"
timeout 100 python tools/attn_bench.py quick | cut -c1-170
for c in base_s128_b256 big_s256_b16 tiny_s40_b32; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$c.err > gpurun_out/bench_$c.json
  python -c "
import json,sys
try:
    j=json.loads(open('gpurun_out/bench_$c.json').read().strip().splitlines()[-1]); print('$c', {k:j.get(k) for k in ('value','ms_per_step','gpu_launches')}, j['roofline']['frac'], j['e2e']['value'])
except Exception as e: print('$c FAILED', e); print(open('gpurun_out/bench_$c.err').read()[-600:])
"
done
