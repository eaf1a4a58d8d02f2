#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py -m gpu -x -q > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_g.log
tail -15 gpurun_out/pytest_gpu_g.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r01_g.json 2> gpurun_out/bench_g_err.log
cat gpurun_out/bench_r01_g.json; tail -5 gpurun_out/bench_g_err.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:recover -s 3 -c 1 -o gpurun_out/prof_recover_r01f -f python tools/dev_bench.py --set onefull > gpurun_out/ncu_e3.log 2>&1
