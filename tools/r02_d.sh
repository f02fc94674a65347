timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02d_pytest.log
timeout 300 python tools/e2e_probe.py C2 100 3 2>&1 | grep -v "^frame" | tail -12 | tee gpurun_out/r02d_e2e_c2.log
timeout 300 python tools/e2e_probe.py C4 100 2 2>&1 | grep -v "^frame" | tail -12 | tee gpurun_out/r02d_e2e_c4.log
python bench.py --no-other-configs > gpurun_out/r02d_bench_c2.json 2> gpurun_out/r02d_bench_c2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02d_bench_c2.json").read().strip().splitlines()[-1])
print("bench", d["value"], d["e2e"]["value"], d["clocks"], d["kernels_ms"], d["roofline"]["block_pair"]["us"])
PY
