TAG=${1:-r02zz}
# Round-2 evidence run on one B200 (bash tools/r02_final.sh [tag]): GPU tests, smoke, bench (C2 default), launch list, ncu summaries (single-CTA and pair
# conv forms, warp v3), in-situ DRAM traffic.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | tee gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; tail -c 300 gpurun_out/${TAG}_bench_c2.json
PNP_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_t4.csv python bench.py --frames 4 --steps 1 --warmup 1 --no-cpu-baseline --no-windows --no-sideinfo > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv3x3_rows_kernel|mv_warp_kernel' -c 3 -o gpurun_out/${TAG}_kernels -f python tools/ncu_target.py > gpurun_out/${TAG}_ncu.log 2>&1; tail -n 1 gpurun_out/${TAG}_ncu.log
PNP_PAIR=1 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_rows_kernel' -c 2 -o gpurun_out/${TAG}_pair -f python tools/ncu_target.py > gpurun_out/${TAG}_ncu_pair.log 2>&1; tail -n 1 gpurun_out/${TAG}_ncu_pair.log
PNP_GRAPHS=0 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -s 150 -c 260 --csv --log-file gpurun_out/${TAG}_insitu_traffic.csv python bench.py --frames 6 --steps 1 --warmup 1 --no-cpu-baseline --no-windows --no-sideinfo > gpurun_out/${TAG}_insitu.log 2>&1; tail -n 2 gpurun_out/${TAG}_insitu_traffic.csv | cut -c1-200
python tools/phase_profile.py 40 > gpurun_out/${TAG}_phase_profile.log 2>&1; cat gpurun_out/${TAG}_phase_profile.log
