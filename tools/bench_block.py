"""Isolated timing of one BAE residual block at a given shape: fused CTA-pair kernel vs the two-launch path.
   python tools/bench_block.py [H W [N]]      (diagnostic; PNP_TRACE=1 dumps the MMA threads' per-step clocks)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnpvcve_b200 import ops  # noqa: E402

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (720, 1280)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((n, h, w, 64), generator=g, device=dev).to(torch.bfloat16)
par = (torch.rand((n, 3, h, w), generator=g, device=dev) > 0.66).float() / 255.0
wt2 = torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.05
wt1 = torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.05
b = torch.randn(64, generator=g, device=dev) * 0.1
ws1 = ops.new_wpack_rowstack(dev, with_par=True)
ops.pack_conv3x3_rowstack(wt2, ws1)
for j in range(3):
    ops.pack_rows(torch.randn((64, 64), generator=g, device=dev) * 0.1, ws1[9 * ops.CHUNK_BYTES:], 64 * j)
ws2 = ops.new_wpack_rowstack(dev)
ops.pack_conv3x3_rowstack(wt1, ws2)
wa = ops.new_wpack(12, dev)
ops.pack_conv3x3(wt2, wa, center_chunks=4)
t = ops.new_feature(n, h, w, dev)
outs = [ops.new_feature(n, h, w, dev) for _ in range(2)]


def fused(i):
    ops.resblock(x if i % 2 == 0 else outs[0], outs[0] if i % 2 == 0 else outs[1], ws1, ws2, par, bias1=b, bias2=b)


def two_launch(i):
    ops.conv3x3(x, wa, out=t, bias=b, par=par, act=ops.PNP_ACT_RELU)
    ops.conv3x3(t, ws2, out=outs[i % 2], idt=x, bias=b, wlayout=1)


def rows_par(i):
    ops.conv3x3(x, ws1, out=t, bias=b, par=par, act=ops.PNP_ACT_RELU, wlayout=1)


def timeit(fn, iters=40):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


flop = 2.0 * (64 * 64 * 9 * 2 + 3 * 64 * 64) * n * h * w
import subprocess  # noqa: E402
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                       stdout=subprocess.PIPE, text=True)
for name, fn in (("fused pair", lambda i: fused(i)), ("fused pair x400", None), ("two launches", two_launch), ("rows kPar (launch A only)", rows_par)):
    us = timeit(fn) if fn is not None else timeit(fused, iters=400)
    print(f"{name:28s} {us:8.1f} us   {flop / us * 1e-6:7.1f} TFLOP/s (block FLOPs)")

smi.terminate()
clk = [ln.split(",")[0].strip() for ln in smi.communicate()[0].strip().splitlines()]
print("SM clock samples (MHz) during the timing loops:", " ".join(clk[:60]))
if os.environ.get("PNP_TRACE"):
    tr = torch.zeros(6 * 128 * 8 + 8 + 2 * 148, dtype=torch.int64, device=dev)
    os.environ["PNP_TRACE_PTR"] = str(tr.data_ptr())
    for i_ in range(6):
        fused(i_)
    torch.cuda.synchronize()
    del os.environ["PNP_TRACE_PTR"]
    print('CTA 0 body cycles:', int(tr[6 * 128 * 8 + 1] - tr[6 * 128 * 8]))
    gt = tr[6 * 128 * 8 + 8:].view(148, 2).cpu()
    g0 = int(gt[:, 0].min())
    print('per-CTA (start, end) in us since first CTA start:')
    print(' '.join(f'{(int(a) - g0) / 1e3:.0f}-{(int(b) - g0) / 1e3:.0f}' for a, b in gt.tolist()))
    dur = sorted((int(b) - int(a)) / 1e3 for a, b in gt.tolist() if b > a)
    print('CTA durations us: min %.0f median %.0f max %.0f' % (dur[0], dur[len(dur) // 2], dur[-1]))
    v = tr[:6 * 128 * 8].view(6, 128, 8).cpu()
    t0 = int(v[0, 0, 0])
    lo, hi = 20, 34
    for role in range(2):
        print(f"role {role} MMA thread: step  begin  +go-check  +end   (cycles since role-0 step 0; d = delta to previous begin)")
        for s_ in range(lo, hi):
            bgn = int(v[role, s_, 0])
            print(f"  {s_:3d} {bgn - t0:8d} {int(v[role, s_, 1]) - bgn:6d} {int(v[role, s_, 2]) - bgn:6d}   d={bgn - int(v[role, s_ - 1, 0])}"
                  f"   scout(step {s_ + 1}): begin {int(v[role, s_ + 1, 3]) - bgn:6d} input-row {int(v[role, s_ + 1, 4]) - bgn:6d} acc-free {int(v[role, s_ + 1, 5]) - bgn:6d}")
    for wv in (2, 3):
        print(f"role 0 epilogue warp {wv + 2}: row  begin | +acc_ready +par_released +staged")
        for k_ in range(lo, hi):
            r = v[wv, k_]
            print(f"  {k_:3d} {int(r[0]) - t0:8d} | " + " ".join(f"{int(r[i]) - int(r[0]):6d}" for i in (1, 3, 2)))
    print("role 1 epilogue: row  begin | +acc_ready +staged")
    for k_ in range(lo, hi):
        r = v[4, k_]
        print(f"  {k_:3d} {int(r[0]) - t0:8d} | " + " ".join(f"{int(r[i]) - int(r[0]):6d}" for i in range(1, 3)))
