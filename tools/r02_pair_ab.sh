for P in 0 1; do PNP_SUSTAINED=1 PNP_PAIR=$P timeout 200 python tools/conv_bench.py; done 2>&1 | tee gpurun_out/r02b_conv_bench4.log
for P in 0 1; do PNP_PAIR=$P python bench.py --frames 30 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02b_bench_pair$P.json 2> gpurun_out/r02b_bench_pair$P.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02b_bench_pair$P.json").read().strip().splitlines()[-1])
print("PAIR=$P", d["value"], d["e2e"]["value"], d["clocks"], d["kernels_ms"], d["roofline"]["block_pair"]["us"])
PY
done
