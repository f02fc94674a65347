"""Stress loop for the host-resident path: N streamed steps of the C2 clip (chunked H2D / D2H concurrent with the
kernels), then N resident steps; any pipeline time-out inside a kernel traps the launch.
   PNP_SPIN_LIMIT='(1u<<21)' python -m pnpvcve_b200.build && python tools/stress_streamed.py [steps] [frames]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import pnpvcve_b200 as P  # noqa: E402
from pnpvcve_b200 import driver, synthetic, weights  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mode = sys.argv[3] if len(sys.argv) > 3 else "both"
cfg = dict(bench.CONFIGS["C2"], name="C2")
dev = torch.device("cuda:0")
net = P.build_backbone(bench.GEN_CFG)
net.load_state_dict(weights.random_state_dict(0), strict=True)
net = net.to(dev).eval()
clip = bench.make_device_batch(cfg, T, 1, 2000, 1, dev)
host = {k: v.cpu().pin_memory() for k, v in clip.items()}
out_host = torch.empty((1, T, 3, cfg["h"], cfg["w"])).pin_memory()
with torch.no_grad():
    ref = net(*synthetic.generator_args(clip)).clone()
    torch.cuda.synchronize()
    if mode in ("both", "streamed"):
        st = driver.ClipStreamer(net, dev, chunk=10)
        ticket = st.upload(host)
        for s in range(steps):
            st.run(ticket, out_host)
            ticket = st.upload(host)
            if s % 5 == 4:
                st.finish()
                torch.cuda.synchronize()
                ok = torch.equal(out_host, ref.cpu())
                print(f"streamed step {s}: identical to the resident result: {ok}", flush=True)
        st.finish()
        torch.cuda.synchronize()
    if mode in ("both", "resident"):
        for s in range(steps):
            out = net(*synthetic.generator_args(clip))
            if s % 5 == 4:
                torch.cuda.synchronize()
                print(f"resident step {s}: identical: {torch.equal(out, ref)}", flush=True)
print("stress done")
