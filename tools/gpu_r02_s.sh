#!/bin/bash
# 2 GPUs: where does the data-parallel step lose its 5-6 %?  NCCL CTA footprint sweep + DP parity test + CLI entry point
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_s_bench2_$tag.json 2> gpurun_out/r02_s_bench2_$tag.err; }
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_s_bench1.json 2> gpurun_out/r02_s_bench1.err
run default FOO=1
run cta4 NCCL_MAX_CTAS=4
run cta8 NCCL_MAX_CTAS=8
run cta16 NCCL_MAX_CTAS=16
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -i "channel\|nvls\|CTA\|Connected\|algo" | head -30 > gpurun_out/r02_s_nccl_info.txt
python -m pytest tests/test_multigpu.py -q -m gpu -s 2>&1 | tail -5 > gpurun_out/r02_s_dp_parity.txt
(cd /tmp && PYTHONPATH=$GRAFT_REPO_ROOT timeout 600 python -m transformer_b200.distributed_train --num_gpu=2 --epochs=1 --dataset_path=/tmp/data/ --ckpt_path=/tmp/ck2 2>&1 | tail -12) > gpurun_out/r02_s_dist_cli.txt
python - <<'PY'
import json
for n in ("bench1","bench2_default","bench2_cta4","bench2_cta8","bench2_cta16"):
    try:
        j=json.loads(open(f"gpurun_out/r02_s_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["ms_per_step"],3), "ms", round(j["value"]), "tok/s", j.get("dp_check"))
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/r02_s_{n}.err").read()[-600:])
PY
tail -3 gpurun_out/r02_s_dp_parity.txt; tail -6 gpurun_out/r02_s_dist_cli.txt
