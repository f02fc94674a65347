timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "resblock" 2>&1 | tail -3
PNP_TRACE=1 timeout 120 python tools/bench_block.py 2>&1 | grep -v "^per-CTA\|^0-" | grep -A15 "fused pair\|role 0 MMA\|role 0 epilogue warp 4"
