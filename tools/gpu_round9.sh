#!/bin/bash
mkdir -p gpurun_out



bash tools/gpu_round8.sh
