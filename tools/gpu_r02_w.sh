#!/bin/bash
# 8 GPUs of one box: BASELINE C3 (weak, 64 sequences per GPU), C4 = transformer-big global batch 128 (the reference's strong-scaling semantics), C3 strong
run8() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r02_w_$tag.json 2> gpurun_out/r02_w_$tag.err; }
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_w_base_n1.json 2> gpurun_out/r02_w_base_n1.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --config big_s256_b16 > gpurun_out/r02_w_big_b16_n1.json 2> gpurun_out/r02_w_big_b16_n1.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --config big_s256_b128 > gpurun_out/r02_w_big_b128_n1.json 2> gpurun_out/r02_w_big_b128_n1.err
run8 base_weak_n8
run8 big_c4_strong_n8 --config big_s256_b128 --scaling strong
run8 base_strong_n8 --scaling strong
python - <<'PY' | tee gpurun_out/r02_w_summary.txt
import json
for n in ("base_n1","big_b16_n1","big_b128_n1","base_weak_n8","big_c4_strong_n8","base_strong_n8"):
    try:
        j=json.loads(open(f"gpurun_out/r02_w_{n}.json").read().strip().splitlines()[-1])
        print(f"{n:18s} {j['ms_per_step']:.3f} ms/step {j['value']:.0f} tok/s  n_gpus {j['n_gpus']} scaling {j['scaling']} global_batch {j['config']['global_batch']} gemm_frac {j['roofline']['frac']:.3f} dp_check {j.get('dp_check')}")
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/r02_w_{n}.err").read()[-600:])
PY
