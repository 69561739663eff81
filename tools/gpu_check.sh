#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, output kept under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
