# compute-sanitizer over the kernels (memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards).
# It also slows the epilogue warps ~100x relative to the tensor pipe, which is how the step-barrier ABA (kStepRing) was
# reproduced: keep the 720p pair-form test in the memcheck list.
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "mv_warp_zero or per_pixel_flow or fetch_pinned or (pair_form and (68 or 72-200 or 720))" > gpurun_out/r02n_memcheck.log 2>&1; grep -vE "^pnp: mbarrier" gpurun_out/r02n_memcheck.log | tail -12
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "mv_warp_zero or per_pixel_flow" > gpurun_out/r02n_racecheck.log 2>&1; tail -4 gpurun_out/r02n_racecheck.log
