#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linear.py -m gpu -x -q -k "streamed or state_constraint or lti" > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_o.log
tail -15 gpurun_out/pytest_gpu_o.log
timeout 300 python tools/dev_bench.py --set hsweep > gpurun_out/hsweep_r01d.jsonl 2> gpurun_out/hsweep_err.log
timeout 300 python tools/dev_bench.py --set lti1 > gpurun_out/lti_r01d.jsonl 2> gpurun_out/lti_err.log
cat gpurun_out/hsweep_r01d.jsonl gpurun_out/lti_r01d.jsonl; tail -3 gpurun_out/lti_err.log gpurun_out/hsweep_err.log
