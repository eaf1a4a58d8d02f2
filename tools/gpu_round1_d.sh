#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py -m gpu -x -q -k "streamed or lti" > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log
tail -15 gpurun_out/pytest_gpu_d.log
timeout 600 python tools/dev_bench.py --set hsweep > gpurun_out/hsweep_r01b.jsonl 2> gpurun_out/hsweep_err.log
timeout 600 python tools/dev_bench.py --set lti1 > gpurun_out/lti_r01b.jsonl 2> gpurun_out/lti_err.log
timeout 600 python tools/dev_bench.py --set nmpc > gpurun_out/nmpc_r01b.jsonl 2> gpurun_out/nmpc_err.log
cat gpurun_out/hsweep_r01b.jsonl gpurun_out/lti_r01b.jsonl gpurun_out/nmpc_r01b.jsonl
