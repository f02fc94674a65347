timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "resblock" 2>&1 | tail -15
timeout 120 python tools/bench_block.py 2>&1 | tail -8
