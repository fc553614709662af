#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu > gpurun_out/r02_colsum_pytest.txt 2>&1
tail -4 gpurun_out/r02_colsum_pytest.txt
for i in 1 2; do
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused', d['ms_per_step'], d['loss_per_token'])"
  B200XF_FUSE_COLSUM=0 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('separate', d['ms_per_step'], d['loss_per_token'])"
done
