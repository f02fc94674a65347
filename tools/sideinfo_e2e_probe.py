"""Per-entry timing of driver.enhance_clips on dense and on compact (records) host entries of the C2 clip: looks for
stalls of the compact feed (device time between entries, host time per call)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pnpvcve_b200 as P
from pnpvcve_b200 import driver, sideinfo, weights
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
net = P.build_backbone(bench.GEN_CFG); net.load_state_dict(weights.random_state_dict(0), strict=True); net = net.to(dev).eval()
clip = bench.make_device_batch(bench.CONFIGS["C2"], T, 1, 2000, 1, dev)
host = {k: v.cpu().pin_memory() for k, v in clip.items()}
out_host = torch.empty((1, T, 3, 720, 1280)).pin_memory()
types = [chr(int(v)) for v in host["slices"][0].flatten()]
tmpl = sideinfo.synthetic_records(720, 1280, "IBBP", seed=77)
per = {"I": [tmpl[0]], "B": [tmpl[1], tmpl[2]], "P": [tmpl[3]]}
recs = [per[s][f % len(per[s])] for f, s in enumerate(types)]
compact = {k: v for k, v in host.items() if k not in ("mvs", "partitions")}
compact["side"] = [sideinfo.pack_side(np.concatenate(recs, 0), np.cumsum([0] + [len(r) for r in recs]), types)]
del clip
torch.cuda.empty_cache()


def leg(entry, tag, steps):
    with torch.no_grad():
        driver.enhance_clips(net, [entry] * 2, device=dev, out_hosts=[out_host] * 2)
        torch.cuda.synchronize()
        for rep in range(2):
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            driver.enhance_clips(net, [entry] * steps, device=dev, out_hosts=[out_host] * steps)
            e1.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(f"{tag} rep {rep}: {steps * T / e0.elapsed_time(e1) * 1e3:.1f} frames/s; host returned after "
                  f"{(t1 - t0) * 1e3:.0f} ms, device done after {(t2 - t0) * 1e3:.0f} ms", flush=True)


leg(host, "dense  ", 3)
leg(compact, "records", 3)
leg(host, "dense  ", 3)
leg(compact, "records", 3)
