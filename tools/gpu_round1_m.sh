#!/bin/bash
# Round-1 closing measurements: tests, bench (ours + reference arm), ncu launch list of the bench command, full captures of the
# dominant kernels of each regime.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_m.log
tail -5 gpurun_out/pytest_gpu_m.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_err.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final_n1.json 2>> gpurun_out/bench_final_err.log
cat gpurun_out/bench_final_n1.json; tail -3 gpurun_out/bench_final_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_onchip -s 3 -c 1 -o gpurun_out/prof_onchip_final -f python tools/dev_bench.py --set onefull > gpurun_out/ncu_m1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_smem -s 1 -c 1 -o gpurun_out/prof_smemk_final -f python tools/dev_bench.py --set h50 > gpurun_out/ncu_m2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_iter -s 40 -c 1 -o gpurun_out/prof_stream_lti_final -f python tools/dev_bench.py --set lti1 > gpurun_out/ncu_m3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_sqp -s 1 -c 1 -o gpurun_out/prof_nmpc_final -f python tools/dev_bench.py --set nmpc1 > gpurun_out/ncu_m4.log 2>&1
ls -la gpurun_out/*final*
