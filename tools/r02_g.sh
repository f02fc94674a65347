timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02g_pytest.log
python bench.py > gpurun_out/r02g_bench_c2.json 2> gpurun_out/r02g_bench_c2.err || tail -5 gpurun_out/r02g_bench_c2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02g_bench_c2.json").read().strip().splitlines()[-1])
print("bench", d["value"], d["e2e"]["value"], d["clocks"], d["kernels_ms"], d["roofline"]["frac"], d["roofline"]["block_pair"]["us"], d["roofline"]["whole_path_frac"], d["roofline_warp"]["frac"])
for k,v in d["other_configs"].items(): print(k, v["frames_per_s"], v["max_abs_err"])
PY
ncu --set full --clock-control none --import-source on -k regex:'mv_warp' -c 1 -o gpurun_out/r02g_warp -f python tools/ncu_target.py > gpurun_out/r02g_ncu.log 2>&1; tail -n 1 gpurun_out/r02g_ncu.log
