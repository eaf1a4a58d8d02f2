#!/bin/bash
# Scaling check: bench.py at N = 1, 2, 4 on one box (launched as the driver does).
set -x
mkdir -p gpurun_out
N=${1:-4}
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_err.log
for n in 2 4 8; do
  [ $n -le $N ] || break
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_n$n.json 2>> gpurun_out/scale_err.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scale_n*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], round(d['value']/1e6,1),'M solves/s', round(d['ms_per_step'],4),'ms/step', 'e2e', round(d['e2e']['value']/1e6,1))
    except Exception as e: print(f,'ERR',e)
PY
tail -5 gpurun_out/scale_err.log
