"""Throughput at the small BASELINE shapes (C1 128x128, C4 320x180 LR) with 1 / 2 clip lanes (diagnostic)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pnpvcve_b200 as P
from pnpvcve_b200 import synthetic, weights
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import GENERATOR_CFG
dev = torch.device("cuda:0")
net = P.build_backbone(dict(GENERATOR_CFG)); net.load_state_dict(weights.random_state_dict(0)); net = net.to(dev).eval()
for name, n, t in (("C1", 4, 7), ("C4", 4, 25), ("C4", 16, 25)):
    clip = synthetic.cat_clips([synthetic.make_config_clip(name, clip_idx=i, t=t, device=dev) for i in range(n)])
    args = synthetic.generator_args(clip)
    for lanes, batch in ((1, False), (2, False), (1, True)):
        net._engine.max_lanes = lanes
        net._engine.batch_clips = batch
        with torch.no_grad():
            for _ in range(2): net(*args)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3): net(*args)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        print(f"{name} n={n} T={t} lanes={lanes} batched={batch}: {n*t/dt:8.1f} frames/s  ({dt*1e3/(n*t):.3f} ms/frame, {net.gpu_launches} launches)", flush=True)
