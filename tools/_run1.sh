timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "resblock" 2>&1 | tail -3
timeout 120 python tools/bench_block.py 2>&1 | grep -A3 "fused pair"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for f in 1 0; do echo "PNP_FUSED_BLOCK=$f"; PNP_FUSED_BLOCK=$f timeout 600 python bench.py --frames 20 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'clocks',d['clocks']['sm_mhz'],d['clocks']['reasons'],'kernels_ms',d['kernels_ms'],'launches',d['gpu_launches'])
"; done
