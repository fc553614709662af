#!/bin/bash
# round 2, call C: head depth 128 kernels + precision experiment with a common component in V
python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -15 > gpurun_out/r02_c_attn_tests.txt
python tools/attn_precision.py > gpurun_out/r02_c_attn_precision.txt 2>&1
python -m pytest tests/test_baseline_shapes_gpu.py -q -m gpu -s -k "default" 2>&1 | tail -40 > gpurun_out/r02_c_default_flags.txt
(cd /tmp && PYTHONPATH=$GRAFT_REPO_ROOT timeout 600 python -m transformer_b200.train --epochs=1 --dataset_path=/tmp/data/ --ckpt_path=/tmp/ck 2>&1 | tail -8) > gpurun_out/r02_c_train_cli_default_flags.txt
cat gpurun_out/r02_c_attn_tests.txt | tail -5; tail -3 gpurun_out/r02_c_default_flags.txt; cat gpurun_out/r02_c_train_cli_default_flags.txt | tail -4
