#!/bin/bash
# same-box A/B of two library builds on the full step: transformer_b200/libb200xf_base.so vs libb200xf.so
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu > gpurun_out/r02_ab_pytest.txt 2>&1
tail -3 gpurun_out/r02_ab_pytest.txt
for i in 1 2; do
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new ', d['ms_per_step'], d['loss_per_token'])"
  B200XF_LIB_PATH=$PWD/transformer_b200/libb200xf_base.so timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('base', d['ms_per_step'], d['loss_per_token'])"
done
