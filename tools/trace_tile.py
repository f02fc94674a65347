"""Per-tile clock64 trace of CTA 0 of the conv kernel (diagnostic)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnpvcve_b200 import ops
dev = torch.device("cuda:0"); h, w = 720, 1280
x = torch.randn((1, h, w, 64), device=dev).to(torch.bfloat16)
out = ops.new_feature(1, h, w, dev)
LAYOUT = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if LAYOUT == 0:
    wp9 = ops.new_wpack(9, dev); ops.pack_conv3x3(torch.randn((64, 64, 3, 3), device=dev) * 0.05, wp9)
else:
    wp9 = ops.new_wpack_rowstack(dev); ops.pack_conv3x3_rowstack(torch.randn((64, 64, 3, 3), device=dev) * 0.05, wp9)
trace = torch.zeros(2 * 64 * 8 + 64 * 16 + 3 * 160, dtype=torch.int64, device=dev)
for _ in range(3): ops.conv3x3(x, wp9, out=out, wlayout=LAYOUT)
torch.cuda.synchronize()
os.environ["PNP_TRACE_PTR"] = str(trace.data_ptr())
PAR = len(sys.argv) > 2 and sys.argv[2] == 'par'
IDT = len(sys.argv) > 2 and sys.argv[2] == 'id'
idt = torch.randn((1, h, w, 64), device=dev).to(torch.bfloat16)
if PAR:
    if LAYOUT == 1:
        wp9 = ops.new_wpack_rowstack(dev, with_par=True); ops.pack_conv3x3_rowstack(torch.randn((64, 64, 3, 3), device=dev) * 0.05, wp9)
    else:
        wp9 = ops.new_wpack(12, dev); ops.pack_conv3x3(torch.randn((64, 64, 3, 3), device=dev) * 0.05, wp9, center_chunks=4)
    par = torch.rand((1, 3, h, w), device=dev)
    bias = torch.randn(64, device=dev)
    ops.conv3x3(x, wp9, out=out, par=par, bias=bias, act=2, wlayout=LAYOUT)
elif IDT:
    ops.conv3x3(x, wp9, out=out, idt=idt, wlayout=LAYOUT)
else:
    ops.conv3x3(x, wp9, out=out, wlayout=LAYOUT)
torch.cuda.synchronize()
t = trace[:512].view(64, 8).cpu()
t2 = trace[512:1024].view(64, 8).cpu()
t3 = trace[1024:2048].view(64, 16).cpu()
body = trace[2048:2048 + 148].cpu()
g0 = trace[2208:2208 + 147].cpu(); g1 = trace[2368:2368 + 147].cpu()
t0 = int(t[0, 0])
print("tile  mma_ready  mma_issued | epi_start  acc_full   epi_math   epi_bar   (cycles since first MMA ready; deltas in brackets)")
prev = None
for i in range(49):
    r = [int(v) - t0 for v in t[i, :8]]
    if LAYOUT == 1:
        d = "" if prev is None else f"  [step period {r[0]-prev[0]:5d}: dx0 +{r[6]-r[0]:4d}  dx1 +{r[7]-r[6]:4d}  lookahead waits +{r[2]-r[7]:4d}  dx2.. +{r[1]-r[2]:4d}  gap {r[0]-prev[1]:4d} | epi math {r[4]-r[3]:4d}  bar {r[5]-r[4]:4d}]"
        w = [int(v) for v in t2[i, :3]]
        d += f"  [wait a_full {w[1]-w[0]:4d}  acc_free {w[2]-w[1]:4d}]"
        print(f"{i:3d} {r[0]:9d} {r[1]:9d} | {r[2]:9d} {r[3]:9d} {r[4]:9d} {r[5]:9d}{d}")
        prev = r
        continue
    d = "" if prev is None else f"  [period {r[0]-prev[0]:5d}  issue {r[1]-r[0]:4d} (4 MMAs +{r[6]-r[0]:4d}, 20 MMAs +{r[7]-r[0]:4d})  gap {r[0]-prev[1]:4d}  accwait {r[3]-r[2]:5d}  math {r[4]-r[3]:4d}  bar {r[5]-r[4]:4d}]"
    print(f"{i:3d} {r[0]:9d} {r[1]:9d} | {r[2]:9d} {r[3]:9d} {r[4]:9d} {r[5]:9d}{d}")
    prev = r

if LAYOUT == 1:
    print("per-MMA issue stamps (cycles since step start), steps 8..20; columns = dx*4+k")
    for i in list(range(0, 4)) + list(range(8, 21)):
        base = int(t[i, 0])
        print(f"{i:3d} " + " ".join(f"{int(v) - base:5d}" for v in t3[i, :12]))

if LAYOUT == 1:
    print(f"CTA body cycles: traced CTA0 {int(body[0])}, others min {int(body[1:].min())} median {int(body[1:].median())} max {int(body[1:].max())}")
    span = int(g1.max() - g0.min()); med = int((g1 - g0).median())
    print(f"body ns: median {med}, first start -> last end {span} ns, start skew {int(g0.max() - g0.min())} ns, end skew {int(g1.max() - g1.min())} ns; clock ~ {int(body[1:147].median()) / med:.3f} GHz")
    # steady-state launch-to-launch period for comparison
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    del os.environ["PNP_TRACE_PTR"]
    e0.record()
    for _ in range(50): ops.conv3x3(x, wp9, out=out, idt=idt if IDT else None, wlayout=LAYOUT)
    e1.record(); torch.cuda.synchronize()
    print(f"back-to-back launches: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per launch")
if LAYOUT == 1 and os.environ.get("PNP_SKEW"):
    import numpy as np
    dur = (g1 - g0).numpy(); end = (g1 - g0.min()).numpy(); start = (g0 - g0.min()).numpy()
    tiles_per = 49
    def steps(c):
        t0, t1 = c * tiles_per, min(7200, (c + 1) * tiles_per)
        n = 0; t = t0
        while t < t1:
            col, y = divmod(t, 720)
            ln = min(720 - y, t1 - t)
            n += ln + (1 if y > 0 else 0) + (1 if y + ln < 720 else 0)
            t += ln
        return n
    st = np.array([steps(c) for c in range(147)])
    print("steps per CTA: min", st.min(), "max", st.max(), "; body ns by step count:", {int(s): int(np.median(dur[st == s])) for s in sorted(set(st))})
    order = np.argsort(end)
    print("earliest 8 ends (cta, steps, end ns):", [(int(c), int(st[c]), int(end[c])) for c in order[:8]])
    print("latest 8 ends   (cta, steps, end ns):", [(int(c), int(st[c]), int(end[c])) for c in order[-8:]])
    print("ns per step by CTA index decile:", [int(np.median((dur / st)[i:i + 15])) for i in range(0, 147, 15)])
