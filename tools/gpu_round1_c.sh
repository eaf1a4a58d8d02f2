#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/dev_bench.py --set nmpc > gpurun_out/nmpc_r01.jsonl 2> gpurun_out/nmpc_err.log
cat gpurun_out/nmpc_r01.jsonl; tail -3 gpurun_out/nmpc_err.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_iter -s 40 -c 1 -o gpurun_out/prof_stream_lti_r01a -f python tools/dev_bench.py --set lti1 > gpurun_out/ncu_f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_iter -s 40 -c 1 -o gpurun_out/prof_stream_h50_r01a -f python tools/dev_bench.py --set h50 > gpurun_out/ncu_g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nmpc_sqp -s 1 -c 1 -o gpurun_out/prof_nmpc_r01a -f python tools/dev_bench.py --set nmpc1 > gpurun_out/ncu_h.log 2>&1
tail -3 gpurun_out/ncu_f.log gpurun_out/ncu_g.log gpurun_out/ncu_h.log
