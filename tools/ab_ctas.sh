#!/bin/bash
# steady-state vs real-batch efficiency of the on-chip kernel for 3 / 2 / 1 CTAs per SM, fused and plain loops
for lib in "" gpurun_in/ab/lib_plain.so; do
  for cap in 3 2 1; do
    echo "== lib=${lib:-fused} ctas/SM=$cap"
    for set in one steady1; do
      MPCB_CTAS_PER_SM=$cap MPCB200_LIB=${lib:+$PWD/$lib} timeout 120 python tools/dev_bench.py --set $set 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$set', d['n'], d['ms'], d.get('frac'))"
    done
  done
done
