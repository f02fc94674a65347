timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02c_pytest.log
timeout 300 python tools/e2e_probe.py C2 100 3 2>&1 | tee gpurun_out/r02c_e2e_c2.log
timeout 300 python tools/e2e_probe.py C4 100 2 2>&1 | tee gpurun_out/r02c_e2e_c4.log
