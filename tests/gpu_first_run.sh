#!/bin/bash
# first-contact script for a gpurun call: quick sanity + timings into gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
