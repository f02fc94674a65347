mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r02a_pytest_gpu.log
python bench.py > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err; tail -c 600 gpurun_out/r02a_bench_c2.json
PNP_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a_launches_bench_t4.csv python bench.py --frames 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02a_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv3x3_rows|mv_warp' -c 3 -o gpurun_out/r02a_kernels -f python tools/ncu_target.py > gpurun_out/r02a_ncu.log 2>&1; tail -n 2 gpurun_out/r02a_ncu.log
