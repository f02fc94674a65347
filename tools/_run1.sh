PNP_TRACE=1 timeout 120 python tools/bench_block.py 2>&1 | grep -v "^per-CTA\|^0-" | tail -75
