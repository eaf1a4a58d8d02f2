#!/bin/bash
# Round-1 closing measurements (final code): tests, bench (ours + reference arm), ncu launch list of the bench command, full
# captures of the dominant kernels, config 5 / re-linearised numbers.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_q.log
tail -4 gpurun_out/pytest_gpu_q.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_q_ref.json 2> gpurun_out/bench_q_err.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_q_n1.json 2>> gpurun_out/bench_q_err.log
cut -c1-400 gpurun_out/bench_q_n1.json; tail -3 gpurun_out/bench_q_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch_q.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_onchip -s 3 -c 1 -o gpurun_out/prof_onchip_q -f python tools/dev_bench.py --set onefull > gpurun_out/ncu_q1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_sqp -s 1 -c 1 -o gpurun_out/prof_nmpc_q -f python tools/dev_bench.py --set nmpc1 > gpurun_out/ncu_q2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_sqp -s 1 -c 1 -o gpurun_out/prof_relin_q -f python tools/dev_bench.py --set relin1 > gpurun_out/ncu_q3.log 2>&1
timeout 300 python tools/dev_bench.py --set nmpc > gpurun_out/nmpc_config5_q.jsonl 2>&1
timeout 300 python tools/dev_bench.py --set relin > gpurun_out/relin_q.jsonl 2>&1
timeout 300 python tools/dev_bench.py --set rows > gpurun_out/rows_q.jsonl 2>&1
timeout 300 python tools/dev_bench.py --set smemk > gpurun_out/smemk_q.jsonl 2>&1
timeout 300 python tools/dev_bench.py --set hsweep > gpurun_out/hsweep_q.jsonl 2>&1
ls -la gpurun_out/*_q*
