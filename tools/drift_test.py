"""Does the resident throughput drift over a long run (thermal / power-cap)?  12 steps of the C2 clip, per-step frames/s
and SM clock; then the same clip through the ClipStreamer (host-resident) for comparison."""
import os, sys, subprocess, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pnpvcve_b200 as P
from pnpvcve_b200 import driver, synthetic, weights
dev = torch.device("cuda:0"); T = 100
net = P.build_backbone(bench.GEN_CFG); net.load_state_dict(weights.random_state_dict(0), strict=True); net = net.to(dev).eval()
clip = bench.make_device_clip(T, 2000, 25, dev)
args = synthetic.generator_args(clip)
def clock():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
with torch.no_grad():
    net(*args); torch.cuda.synchronize()
    for s in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net(*args); e1.record()
        c = clock(); torch.cuda.synchronize()
        print(f"resident step {s}: {T / e0.elapsed_time(e1) * 1e3:6.1f} frames/s   [{c}]", flush=True)
    host = {k: v.cpu().pin_memory() for k, v in clip.items()}
    out_host = torch.empty((1, T, 3, 720, 1280)).pin_memory()
    st = driver.ClipStreamer(net, dev, chunk=10)
    ticket = st.upload(host)
    for s in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nxt = st.upload(host)
        st.run(ticket, out_host); ticket = nxt
        e1.record()
        c = clock(); torch.cuda.synchronize()
        print(f"streamed step {s}: {T / e0.elapsed_time(e1) * 1e3:6.1f} frames/s   [{c}]", flush=True)
    for s in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net(*args); e1.record()
        c = clock(); torch.cuda.synchronize()
        print(f"resident again {s}: {T / e0.elapsed_time(e1) * 1e3:6.1f} frames/s   [{c}]", flush=True)
