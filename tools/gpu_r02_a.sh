#!/bin/bash
# round 2, call A: full GPU test suite (incl. the new BASELINE-shape parity tests), bench, CLI entry points on synthetic text
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02_a_gpu.txt
python -m pytest tests -q -m gpu -x --deselect tests/test_baseline_shapes_gpu.py 2>&1 | tail -15 > gpurun_out/r02_a_pytest_gpu.txt
python -m pytest tests/test_baseline_shapes_gpu.py -q -m gpu -s 2>&1 | tail -120 > gpurun_out/r02_a_pytest_baseline_shapes.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_a_bench.json 2> gpurun_out/r02_a_bench.err
(cd /tmp && PYTHONPATH=$GRAFT_REPO_ROOT timeout 600 python -m transformer_b200.train --num_heads=8 --epochs=1 --dataset_path=/tmp/data/ --ckpt_path=/tmp/ck 2>&1 | tail -15) > gpurun_out/r02_a_train_cli.txt
tail -3 gpurun_out/r02_a_pytest_gpu.txt; tail -5 gpurun_out/r02_a_pytest_baseline_shapes.txt; cat gpurun_out/r02_a_bench.json | head -c 1500; tail -5 gpurun_out/r02_a_train_cli.txt
