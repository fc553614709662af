#!/bin/bash
# 2 GPUs: bucket granularity x NCCL CTA budget
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_u_$tag.json 2> gpurun_out/r02_u_$tag.err; }
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_u_n1.json 2> gpurun_out/r02_u_n1.err
run coarse FOO=1
run coarse_cta8 NCCL_MAX_CTAS=8
run coarse_cta16 NCCL_MAX_CTAS=16
run layer_cta8 B200XF_DP_BUCKETS=layer NCCL_MAX_CTAS=8
run layer_cta4 B200XF_DP_BUCKETS=layer NCCL_MAX_CTAS=4
run coarse_b FOO=1
python - <<'PY' | tee gpurun_out/r02_u_summary.txt
import json
for n in ("n1","coarse","coarse_cta8","coarse_cta16","layer_cta8","layer_cta4","coarse_b"):
    try:
        j=json.loads(open(f"gpurun_out/r02_u_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["ms_per_step"],3), "ms", round(j["value"]), "tok/s")
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/r02_u_{n}.err").read()[-600:])
PY
