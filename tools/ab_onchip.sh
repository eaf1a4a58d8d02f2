#!/bin/bash
# A/B of kernel build variants on one box: gpurun_in/ab/lib_*.so against the in-tree library, interleaved twice.
for rep in 1 2; do
  for lib in "" gpurun_in/ab/lib_00.so gpurun_in/ab/lib_10.so gpurun_in/ab/lib_01.so; do
    echo "== lib=${lib:-default} rep=$rep"
    MPCB200_LIB=${lib:+$PWD/$lib} timeout 120 python tools/dev_bench.py --set one 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms'], d['frac'])"
    MPCB200_LIB=${lib:+$PWD/$lib} timeout 120 python tools/dev_bench.py --set h50 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('h50', d['ms'])"
  done
done
