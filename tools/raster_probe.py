import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from pnpvcve_b200 import sideinfo
dev = torch.device("cuda:0")
tmpl = sideinfo.synthetic_records(720, 1280, "IBBP", seed=77)
for T in (4, 40, 100):
    types = (["I"] + ["B", "B", "P"] * 40)[:T]
    per = {"I": [tmpl[0]], "B": [tmpl[1], tmpl[2]], "P": [tmpl[3]]}
    recs = [per[s][f % len(per[s])] for f, s in enumerate(types)]
    flat = np.concatenate(recs, 0); offs = np.cumsum([0] + [len(r) for r in recs])
    side = sideinfo.pack_side(flat, offs, types)
    mvs = torch.empty((T, 4, 720, 1280), device=dev); par = torch.empty((T, 3, 720, 1280), device=dev)
    work = torch.empty((3, T, 720, 1280), dtype=torch.int32, device=dev); status = torch.zeros(1, dtype=torch.int32, device=dev)
    rec, meta, nb = sideinfo.upload_side(side, dev)
    for _ in range(2): sideinfo.rasterize_uploaded(rec, meta, T, mvs, par, work, status)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): sideinfo.rasterize_uploaded(rec, meta, T, mvs, par, work, status)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"T={T}: rasterise {ms:.2f} ms = {ms / T * 1e3:.1f} us per frame, records {len(flat)}, status {int(status.item())}")
