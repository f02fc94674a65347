"""Resident generator rate on the bench's synthetic block-constant motion fields against fields rasterised from
H.264-style per-block records (sideinfo.synthetic_records), and the warp kernel's bracketed time on each."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pnpvcve_b200 as P
from pnpvcve_b200 import sideinfo, synthetic, weights
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 30
net = P.build_backbone(bench.GEN_CFG); net.load_state_dict(weights.random_state_dict(0), strict=True); net = net.to(dev).eval()
clip = bench.make_device_batch(bench.CONFIGS["C2"], T, 1, 2000, 1, dev)
types = [chr(int(v)) for v in clip["slices"][0].flatten()]
tmpl = sideinfo.synthetic_records(720, 1280, "IBBP", seed=77)
per_type = {"I": [tmpl[0]], "B": [tmpl[1], tmpl[2]], "P": [tmpl[3]]}
recs = [per_type[st][f % len(per_type[st])] for f, st in enumerate(types)]
mv, par = sideinfo.rasterize_clip(np.concatenate(recs, 0), np.cumsum([0] + [len(r) for r in recs]), types, 720, 1280, device=dev)
clip2 = dict(clip, mvs=mv[None].contiguous(), partitions=par[None].contiguous())


def rate(c, tag):
    args = synthetic.generator_args(c)
    with torch.no_grad():
        for _ in range(2): net(*args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net(*args); net(*args); e1.record(); torch.cuda.synchronize()
        fps = 2 * T / e0.elapsed_time(e1) * 1e3
        net._engine.prof = {"warp": [], "block_a": []}; net._engine.prof_every = 4
        net(*args); torch.cuda.synchronize()
        pr = net._engine.prof; net._engine.prof = None
    w = [a.elapsed_time(b) * 1e3 for a, b in pr["warp"]]
    a_ = [a.elapsed_time(b) * 1e3 for a, b in pr["block_a"]]
    print(f"{tag}: {fps:.1f} frames/s; warp bracketed mean {sum(w) / len(w):.1f} us (min {min(w):.1f}, max {max(w):.1f}); launch A {sum(a_) / len(a_):.1f} us")


for _ in range(3):          # alternate: under the power cap the rate drifts down over the first seconds of load
    rate(clip, "synthetic 8x8 block-constant fields")
    rate(clip2, "fields rasterised from records   ")
rate(dict(clip, mvs=clip2["mvs"]), "records mvs + synthetic partitions")
rate(dict(clip, partitions=clip2["partitions"]), "synthetic mvs + records partitions")
