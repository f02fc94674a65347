"""Throughput of the other BASELINE shapes with the final kernels (diagnostic, not the bench line):
C3 (IPB conditioning, 720p), C4 (320x180, many clips batched), C5 (KITTI 376x1244 pairs), and the vsr=True x4 tail on
C4-shape input (output 720x1280)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pnpvcve_b200 as P
from pnpvcve_b200 import synthetic, weights
import bench
dev = torch.device("cuda:0")
def build(vsr=False):
    net = P.build_backbone(dict(bench.GEN_CFG, vsr=vsr)); net.load_state_dict(weights.random_state_dict(0, vsr=vsr)); return net.to(dev).eval()
def run(net, name, n, t, reps=3):
    clip = synthetic.cat_clips([synthetic.make_config_clip(name, clip_idx=i, t=t, device=dev) for i in range(n)])
    args = synthetic.generator_args(clip)
    with torch.no_grad():
        for _ in range(2): net(*args)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): net(*args)
        e1.record(); torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / reps / 1e3
    h, w = clip["lq"].shape[-2:]
    print(f"{name} {w}x{h} n={n} T={t} vsr={bool(net.vsr)}: {n*t/dt:8.1f} frames/s  ({dt*1e3/(n*t):.3f} ms/frame, "
          f"{n*t*h*w/dt/1e6:7.1f} Mpx/s, {net.gpu_launches} launches)", flush=True)
net = build()
run(net, "C3", 1, 40)
run(net, "C4", 16, 25)
run(net, "C4", 32, 25)
run(net, "C5", 8, 2, reps=10)
run(net, "C5", 16, 2, reps=10)
del net
run(build(vsr=True), "C4", 4, 25)
