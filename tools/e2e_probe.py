"""Where does the end-to-end (host-resident) rate go?  Per-step frames/s of the selected bench config: resident clip,
then through driver.ClipStreamer with both copy directions, upload only, download only, and neither.
   python tools/e2e_probe.py [C2|C3|C4|C5] [frames] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import pnpvcve_b200 as P  # noqa: E402
from pnpvcve_b200 import driver, synthetic, weights  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = dict(bench.CONFIGS[name], name=name)
T = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["t"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
net = P.build_backbone(bench.GEN_CFG)
net.load_state_dict(weights.random_state_dict(0), strict=True)
net = net.to(dev).eval()
clip = bench.make_device_batch(cfg, T, cfg["clips"], 2000, 1, dev)
args = synthetic.generator_args(clip)
n = cfg["clips"]


def fps(ms, k=1):
    return k * n * T / ms * 1e3


with torch.no_grad():
    net(*args)
    torch.cuda.synchronize()
    for s in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        net(*args)
        e1.record()
        torch.cuda.synchronize()
        print(f"resident step {s}: {fps(e0.elapsed_time(e1)):7.1f} frames/s", flush=True)
    host = {k: v.cpu().pin_memory() for k, v in clip.items()}
    out_host = torch.empty((n, T, 3, cfg["h"], cfg["w"])).pin_memory()
    import time
    st = driver.ClipStreamer(net, dev, chunk=max(1, min(10, T)))
    for s in range(3):       # the upload alone (no kernels running): host enqueue time and transfer rate
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st.upload(host)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"upload alone {s}: enqueue {1e3 * (t1 - t0):.1f} ms, done after {1e3 * (t2 - t0):.1f} ms = "
              f"{st.h2d_bytes / (t2 - t0) / 1e9:.1f} GB/s", flush=True)
    del st
    for cin, cout in ((True, True), (True, False), (False, True), (False, False), (True, True)):
        st = driver.ClipStreamer(net, dev, chunk=max(1, min(10, T)))
        st.copy_in, st.copy_out = cin, cout
        ticket = st.upload(host)
        rates = []
        w0 = torch.cuda.Event(enable_timing=True)
        for s in range(steps + 2):
            if s == 2:
                w0.record()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            st.run(ticket, out_host)
            ticket = st.upload(host)
            e1.record()
            rates.append((e0, e1))
        st.finish()
        w1 = torch.cuda.Event(enable_timing=True)
        w1.record()
        torch.cuda.synchronize()
        per = " ".join(f"{fps(a.elapsed_time(b)):7.1f}" for a, b in rates)
        print(f"streamed h2d={int(cin)} d2h={int(cout)}: per step {per}   steady {fps(w0.elapsed_time(w1), steps):7.1f} frames/s",
              flush=True)
