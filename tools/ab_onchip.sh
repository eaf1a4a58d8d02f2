#!/bin/bash
# A/B of kernel build variants on one box: every gpurun_in/ab/lib_*.so against the in-tree library, interleaved twice.
for rep in 1 2; do
  for lib in "" $(ls gpurun_in/ab/lib_*.so 2>/dev/null); do
    echo "== lib=${lib:-default} rep=$rep"
    for set in ${AB_SETS:-one h50}; do
      MPCB200_LIB=${lib:+$PWD/$lib} timeout 120 python tools/dev_bench.py --set $set 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$set', d['ms'], d.get('frac'), d.get('mean_iters'))"
    done
  done
done
