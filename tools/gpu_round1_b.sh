#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nmpc.py -m gpu -x -q > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_b.log
tail -40 gpurun_out/pytest_gpu_b.log


