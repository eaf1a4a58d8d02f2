#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_linear.py -m gpu -x -q > gpurun_out/pytest_gpu_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_l.log
tail -25 gpurun_out/pytest_gpu_l.log
timeout 600 python tools/dev_bench.py --set hsweep > gpurun_out/hsweep_r01c.jsonl 2> gpurun_out/hsweep_err.log
cat gpurun_out/hsweep_r01c.jsonl; tail -3 gpurun_out/hsweep_err.log
