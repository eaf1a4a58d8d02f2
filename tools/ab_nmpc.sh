#!/bin/bash
# A/B of NMPC kernel build variants (gpurun_in/ab/lib_*.so) against the in-tree library
for lib in "" $(ls gpurun_in/ab/lib_*.so 2>/dev/null); do
  echo "== lib=${lib:-default}"
  MPCB200_LIB=${lib:+$PWD/$lib} timeout 200 python tools/dev_bench.py --set nmpc 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{') and 'cfg' in l:
        d=json.loads(l); print(d['fixture'], d['n'], d['ms'], d['solves_per_s'])"
done
