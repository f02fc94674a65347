timeout 500 ncu --set full --clock-control none --import-source on -k regex:resblock_pair -s 6 -c 1 -f -o gpurun_out/block_prof3 python tools/bench_block.py > gpurun_out/block_prof3.log 2>&1
tail -2 gpurun_out/block_prof3.log
