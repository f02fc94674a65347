#!/bin/bash
# A/B of two builds of libpnpvcve.so on the SAME box (box-to-box variation is larger than most kernel tweaks)
#   tools/ab_bench.sh <other.so> [bench args]
OTHER=$1; shift
ARGS=${@:---steps 2 --warmup 2 --no-cpu-baseline}
show() { python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$1 fps %.1f e2e %.1f clocks %s kernels_us %s' % (d['value'], d['e2e']['value'], d['clocks'].get('sm_mhz'), {k: round(v*1e3,1) for k,v in d['kernels_ms'].items()}))"; }
for i in 1 2; do
  python bench.py $ARGS 2>/dev/null | show "new  "
  PNP_LIB_PATH=$OTHER python bench.py $ARGS 2>/dev/null | show "other"
done
