#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err
echo "bench exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_2gpu.log'):
    try: j=json.loads(l)
    except Exception: continue
    print({k:j.get(k) for k in ('value','ms_per_step','n_gpus','launch_mode')}, j['e2e'])
PY
timeout 150 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 140 2>&1 | tail -3
