timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_metrics.py -m gpu -x -q -s 2>&1 | grep -E "seam|passed|failed|Error|error" | tail -8 | tee gpurun_out/r02e_pytest.log
python bench.py --no-other-configs > gpurun_out/r02e_bench_c2.json 2> gpurun_out/r02e_bench_c2.err || tail -5 gpurun_out/r02e_bench_c2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02e_bench_c2.json").read().strip().splitlines()[-1])
print("bench", d["value"], d["e2e"], d["clocks"], d["frame_windows"])
PY
