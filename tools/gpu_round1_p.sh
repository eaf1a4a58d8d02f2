#!/bin/bash
# Round-1 closing measurements: tests, bench (ours + reference arm), ncu launch list of the bench command, full captures of the
# dominant kernels of each regime.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_p.log
tail -5 gpurun_out/pytest_gpu_p.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final2_ref.json 2> gpurun_out/bench_final2_err.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final2_n1.json 2>> gpurun_out/bench_final2_err.log
cat gpurun_out/bench_final2_n1.json; tail -3 gpurun_out/bench_final2_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final2.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_onchip -s 3 -c 1 -o gpurun_out/prof_onchip_final2 -f python tools/dev_bench.py --set onefull > gpurun_out/ncu_m1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_smem -s 1 -c 1 -o gpurun_out/prof_smemk_final2 -f python tools/dev_bench.py --set h50 > gpurun_out/ncu_m2.log 2>&1
ls -la gpurun_out/*final*
