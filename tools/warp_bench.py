"""mv_warp_kernel under a COLD L2 (rotating buffer sets, ~1 GB footprint), synthetic block-constant quarter-pel motion
like the bench: TMA-staged tap windows (default) against global gathers only (PNP_WARP_TMA=0, read once per process)."""
import os, sys, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "run":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from pnpvcve_b200 import ops
    dev = torch.device("cuda:0"); h, w = 720, 1280
    NS = 4
    g = torch.Generator(device=dev).manual_seed(0)
    srcs = [torch.randn((1, h, w, 64), generator=g, device=dev).to(torch.bfloat16) for _ in range(NS)]
    dsts = [ops.new_feature(1, h, w, dev) for _ in range(NS)]
    flows = [(torch.randint(-64, 65, (1, 2, h // 8, w // 8), generator=g, device=dev).float() / 4)
             .repeat_interleave(8, 2).repeat_interleave(8, 3).contiguous() for _ in range(NS)]
    if os.environ.get("WARP_FIELD") == "shifted":     # vector blocks at unaligned positions (reversed P-frame vectors)
        flows = [torch.roll(f, shifts=(5, 3), dims=(2, 3)).contiguous() for f in flows]
    def run(iters):
        for i in range(iters):
            ops.mv_warp(srcs[i % NS], flows[i % NS], dsts[i % NS])
    run(8); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(200); e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) / 200 * 1e3
    print(f"PNP_WARP_TMA={os.environ.get('PNP_WARP_TMA', '1')}: {us:6.1f} us per 720p warp = {264 * h * w / us * 1e-3:7.1f} GB/s algorithmic (cold L2)")
else:
    for v, field in (("0", "aligned"), ("1", "aligned"), ("1", "shifted")):
        print(f"vector blocks {field}: ", end="", flush=True)
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, PNP_WARP_TMA=v, WARP_FIELD=field))
