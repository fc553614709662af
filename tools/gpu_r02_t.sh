#!/bin/bash
# 2 GPUs: per-layer gradient buckets + split Adam
python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/r02_t_pytest_gpu_2gpu_box.txt; tail -2 gpurun_out/r02_t_pytest_gpu_2gpu_box.txt
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_t_bench1.json 2> gpurun_out/r02_t_bench1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_t_bench2.json 2> gpurun_out/r02_t_bench2.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_t_bench1b.json 2> gpurun_out/r02_t_bench1b.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_t_bench2b.json 2> gpurun_out/r02_t_bench2b.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --scaling strong > gpurun_out/r02_t_bench2_strong.json 2> gpurun_out/r02_t_bench2_strong.err
python - <<'PY'
import json
for n in ("bench1","bench2","bench1b","bench2b","bench2_strong"):
    try:
        j=json.loads(open(f"gpurun_out/r02_t_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["ms_per_step"],3), "ms", round(j["value"]), "tok/s", j["scaling"], j["config"]["global_batch"], j.get("dp_check"))
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/r02_t_{n}.err").read()[-900:])
PY
