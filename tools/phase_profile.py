"""Where a frame's time goes inside the real engine: events at the phase boundaries of each frame step
(im2col + warp + input convs | 8-block stack | head), 7 per frame -- light enough not to disturb the run."""
import os, sys, collections, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pnpvcve_b200 as P
from pnpvcve_b200 import synthetic, weights
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 30
net = P.build_backbone(bench.GEN_CFG); net.load_state_dict(weights.random_state_dict(0), strict=True); net = net.to(dev).eval()
clip = bench.make_device_batch(bench.CONFIGS["C2"], T, 1, 2000, 1, dev)
args = synthetic.generator_args(clip)
with torch.no_grad():
    for _ in range(2): net(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); net(*args); e1.record(); host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    clean = e0.elapsed_time(e1)
    net._engine.prof = {"phases": []}
    e0.record(); net(*args); e1.record(); torch.cuda.synchronize()
    ph = net._engine.prof["phases"]; net._engine.prof = None
print(f"T={T}: clean {clean:.1f} ms = {clean / T * 1e3:.0f} us/frame ({T / clean * 1e3:.1f} fps); host enqueue time {host_ms:.1f} ms; with phase events {e0.elapsed_time(e1):.1f} ms; launches {net.gpu_launches}")
acc = collections.defaultdict(list)
for (n0, a), (n1, b) in zip(ph[:-1], ph[1:]):
    acc[f"{n0} -> {n1}"].append(a.elapsed_time(b) * 1e3)
for k, v in acc.items():
    print(f"  {k:36s} n={len(v):4d}  mean {sum(v) / len(v):8.1f} us   total {sum(v) / 1e3:8.2f} ms")
