# Validation of the current HEAD on one B200: GPU tests, smoke, bench (C2 default), phase profile.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r02q_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | tee gpurun_out/r02q_smoke.log
python bench.py > gpurun_out/r02q_bench_c2.json 2> gpurun_out/r02q_bench_c2.err; tail -c 300 gpurun_out/r02q_bench_c2.json
python tools/phase_profile.py 40 > gpurun_out/r02q_phase_profile.log 2>&1; cat gpurun_out/r02q_phase_profile.log
