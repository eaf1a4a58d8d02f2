#!/bin/bash
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r01_n2b.json 2> gpurun_out/bench_n2b_err.log
cat gpurun_out/bench_r01_n2b.json; tail -3 gpurun_out/bench_n2b_err.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r01_n2ref.json 2>> gpurun_out/bench_n2b_err.log
cat gpurun_out/bench_r01_n2ref.json
timeout 900 python tools/dev_bench.py --set lti > gpurun_out/lti_r01c.jsonl 2> gpurun_out/lti_err.log
cat gpurun_out/lti_r01c.jsonl; tail -3 gpurun_out/lti_err.log
