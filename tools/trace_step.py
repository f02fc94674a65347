"""Where does the MMA-issuing thread of launch A spend a step?  PNP_DIAG build only:
   PNP_DIAG=1 python -m pnpvcve_b200.build && python tools/trace_step.py [plain|idt|par]
clock64 stamps of CTA 0 (pnp_conv_rows.cu, PNP_TRACING): step start, after each (dx, k) emit, before / after the look-ahead
poll of the next step's barriers, step end (behind the 1x1 MMAs and their commit)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
trace = torch.zeros(4096, dtype=torch.int64, device=dev)
os.environ["PNP_TRACE_PTR"] = str(trace.data_ptr())
from pnpvcve_b200 import ops  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "par"
h, w = 720, 1280
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((1, h, w, 64), generator=g, device=dev).to(torch.bfloat16)
out = ops.new_feature(1, h, w, dev)
war = ops.new_wpack_rowstack(dev, with_par=True)
ops.pack_conv3x3_rowstack(torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.05, war)
for j in range(3):
    ops.pack_rows(torch.randn((64, 64), generator=g, device=dev) * 0.1, war[9 * ops.CHUNK_BYTES:], 64 * j)
par = (torch.rand((1, 3, h, w), generator=g, device=dev) > 0.6).float() / 255.0
bias = torch.randn(64, generator=g, device=dev) * 0.1


def launch():
    if kind == "par":
        ops.conv3x3(x, war, out=out, bias=bias, par=par, act=ops.PNP_ACT_RELU)
    elif kind == "idt":
        ops.conv3x3(x, war, out=out, bias=bias, idt=x)
    else:
        ops.conv3x3(x, war, out=out, bias=bias)


for _ in range(3):
    launch()
torch.cuda.synchronize()
trace.zero_()
launch()
torch.cuda.synchronize()
t = trace.cpu().tolist()
steps = [s for s in range(1, 50) if t[s * 8] and t[(s + 1) * 8]]
print(f"{kind}: step period / issue of the 12 MMAs / [dx=2 poll] / tail behind the last 3x3 MMA (1x1 MMAs, region wait, commit)")
rows = []
for s in steps[4:44]:
    start, nxt = t[s * 8], t[(s + 1) * 8]
    last3x3 = t[1024 + s * 16 + 11]
    poll = t[s * 8 + 2] - t[s * 8 + 7]
    rows.append((nxt - start, last3x3 - start, poll, t[s * 8 + 1] - last3x3))
for r in rows[:16]:
    print("  period %5d   12 MMAs issued after %5d   look-ahead poll %4d   tail %5d" % r)
n = len(rows)
print("  mean period %.0f, 12-MMA issue %.0f, poll %.0f, tail %.0f over %d steps; CTA 0 body %d cycles" % (
    sum(r[0] for r in rows) / n, sum(r[1] for r in rows) / n, sum(r[2] for r in rows) / n, sum(r[3] for r in rows) / n, n,
    t[2048]))

print("per-emit deltas (start -> dx0k0 .. dx2k3 -> step end), a few steps:")
for s_ in steps[10:18]:
    st = [t[s_ * 8]] + [t[1024 + s_ * 16 + i] for i in range(12)] + [t[s_ * 8 + 1]]
    print("  step %2d:" % s_, " ".join("%4d" % (b - a) for a, b in zip(st, st[1:])), "| dx2 poll %4d" % (t[s_ * 8 + 2] - t[s_ * 8 + 7]))

if kind == "par":
    print("per step s (source row s-1 of the first segment), cycles relative to the MMA thread's step start:")
    print("  scout: a_full seen / published | readers (row s-1): par_done seen / region released / parked | main (row s-2): loop top / step_done seen / released / blend available")
    for s_ in steps[10:22]:
        b = t[s_ * 8]
        sc = [t[2560 + (s_ + 1) * 4 + i] - b for i in range(3)]         # the scout works on the NEXT step
        rd = [t[3072 + (s_ - 1) * 4 + i] - b for i in range(4)]
        mn = [t[2816 + (s_ - 2) * 4 + i] - b for i in range(4)]
        print("  step %2d: scout(next) %5d %5d %5d | readers %5d %5d %5d %5d | main %5d %5d %5d %5d | step end %5d" % (
            s_, sc[0], sc[1], sc[2], rd[0], rd[1], rd[2], rd[3], mn[0], mn[1], mn[2], mn[3], t[(s_ + 1) * 8] - b))
