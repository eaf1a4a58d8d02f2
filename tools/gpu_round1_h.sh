#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h.log
tail -30 gpurun_out/pytest_gpu_h.log
