#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_e.log
tail -15 gpurun_out/pytest_gpu_e.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_e.json 2> gpurun_out/bench_e_err.log
cat gpurun_out/bench_r01_e.json; tail -5 gpurun_out/bench_e_err.log
