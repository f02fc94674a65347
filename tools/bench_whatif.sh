#!/bin/bash
# in-bench what-if: how memory-sensitive are the conv launches inside the real frame loop? (results are WRONG with skip bits)
for b in 0 1 2 3; do
  PNP_DEBUG_SKIP=$b python bench.py --frames 20 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('skip=$b fps %.1f  clocks %s  kernels_ms %s' % (d['value'], d['clocks'].get('sm_mhz'), {k: round(v*1e3,1) for k,v in d['kernels_ms'].items()}))"
done
