#!/bin/bash
# GPU session A: parity tests, bench, ncu launch list + full captures of the two QT kernels, configs 3/4 dev numbers.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_c.json 2> gpurun_out/bench_c_err.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_ref.json 2>> gpurun_out/bench_c_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:admm_onchip -s 3 -c 1 -o gpurun_out/prof_onchip_r01d -f python tools/dev_bench.py --set onefull > gpurun_out/ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:recover -s 3 -c 1 -o gpurun_out/prof_recover_r01d -f python tools/dev_bench.py --set onefull > gpurun_out/ncu_e.log 2>&1
timeout 600 python tools/dev_bench.py --set hsweep > gpurun_out/hsweep_r01.jsonl 2> gpurun_out/hsweep_err.log
timeout 900 python tools/dev_bench.py --set lti > gpurun_out/lti_r01.jsonl 2> gpurun_out/lti_err.log
tail -3 gpurun_out/pytest_gpu_a.log; cat gpurun_out/hsweep_r01.jsonl gpurun_out/lti_r01.jsonl; tail -5 gpurun_out/lti_err.log
