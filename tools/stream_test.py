"""ClipStreamer == plain forward (bit for bit) on a small clip, and timing at 720p T=30."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pnpvcve_b200 as P
from pnpvcve_b200 import driver, synthetic, weights
dev = torch.device("cuda:0")
net = P.build_backbone(bench.GEN_CFG); net.load_state_dict(weights.random_state_dict(0), strict=True); net = net.to(dev).eval()
clip = synthetic.make_clip(64, 96, 23, seed=5)
host = {k: v.pin_memory() for k, v in clip.items()}
with torch.no_grad():
    ref = net(*[a.to(dev) for a in synthetic.generator_args(clip)]).cpu()
outs = [torch.empty_like(ref).pin_memory() for _ in range(3)]
driver.stream_clips(net, [host] * 3, outs, dev, chunk=5)
torch.cuda.synchronize()
print("streamed == plain:", [bool((o == ref).all()) for o in outs])
