#!/bin/bash
# 2 GPUs: chunked GEMM schedule during the data-parallel backward
python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "chunked or gemm" 2>&1 | tail -3
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_v_$tag.json 2> gpurun_out/r02_v_$tag.err; }
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_v_n1.json 2> gpurun_out/r02_v_n1.err
run chunks0 B200XF_DP_GEMM_CHUNKS=0
run chunks4 B200XF_DP_GEMM_CHUNKS=4
run chunks2 B200XF_DP_GEMM_CHUNKS=2
run chunks8 B200XF_DP_GEMM_CHUNKS=8
run chunks4_layer B200XF_DP_GEMM_CHUNKS=4 B200XF_DP_BUCKETS=layer
python -m pytest tests/test_multigpu.py -q -m gpu 2>&1 | tail -2
python - <<'PY' | tee gpurun_out/r02_v_summary.txt
import json
for n in ("n1","chunks0","chunks4","chunks2","chunks8","chunks4_layer"):
    try:
        j=json.loads(open(f"gpurun_out/r02_v_{n}.json").read().strip().splitlines()[-1])
        print(n, round(j["ms_per_step"],3), "ms", round(j["value"]), "tok/s", (j.get("dp_check") or {}).get("weights_rel_diff_vs_1gpu"))
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/r02_v_{n}.err").read()[-600:])
PY
