#!/bin/bash
# N-GPU data-parallel bench exactly as the driver launches it (N from $1)
N=${1:-4}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
echo "bench exit $?"
python - <<PY
import json
for l in open('gpurun_out/bench_${N}gpu.log'):
    try: j=json.loads(l)
    except Exception: continue
    print({k:j.get(k) for k in ('value','ms_per_step','n_gpus','launch_mode')}, j['e2e'])
PY
tail -3 gpurun_out/bench_${N}gpu.err
