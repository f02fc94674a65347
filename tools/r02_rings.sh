for R in "4,5" "5,4" "6,4" "4,4"; do echo "PNP_RINGS=$R"; PNP_RINGS=$R PNP_SUSTAINED=1 timeout 200 python tools/conv_bench.py 2>&1 | grep -E "plain|launch B"; done | tee gpurun_out/r02j_rings.log
