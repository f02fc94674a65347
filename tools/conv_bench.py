"""Hot-loop timing of the conv launches of one BAE block at a BASELINE shape (not a benchmark of the path: bench.py is):
   python tools/conv_bench.py [H W [N]]      # PNP_PAIR=0 selects the single-CTA form of the kernel
Prints us per launch for: plain 64->64, launch B (+identity, bottom-up), launch A (3x3 + three partition 1x1 convs), the
A+B pair back to back, and an output checksum of each (compare across PNP_PAIR settings)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnpvcve_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (720, 1280)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((n, h, w, 64), generator=g, device=dev).to(torch.bfloat16)
t = ops.new_feature(n, h, w, dev)
out = ops.new_feature(n, h, w, dev)
war = ops.new_wpack_rowstack(dev, with_par=True)
ops.pack_conv3x3_rowstack(torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.05, war)
for j in range(3):
    ops.pack_rows(torch.randn((64, 64), generator=g, device=dev) * 0.1, war[9 * ops.CHUNK_BYTES:], 64 * j)
wb = ops.new_wpack_rowstack(dev)
ops.pack_conv3x3_rowstack(torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.05, wb, flip_ky=True)
par = (torch.rand((n, 3, h, w), generator=g, device=dev) > 0.6).float() / 255.0
bias = torch.randn(64, generator=g, device=dev) * 0.1


def plain():
    ops.conv3x3(x, wb, out=out, bias=bias, act=ops.PNP_ACT_LRELU)


def launch_b():
    ops.conv3x3(t, wb, out=out, idt=x, bias=bias, flip_y=True)


def launch_a():
    ops.conv3x3(x, war, out=t, bias=bias, par=par, act=ops.PNP_ACT_RELU)


wl = ops.new_wpack_rowstack(dev, tap_n=16)
ops.pack_conv3x3_rowstack(torch.randn((3, 64, 3, 3), generator=g, device=dev) * 0.05, wl, tap_n=16)
lq = torch.rand((n, 3, h, w), generator=g, device=dev)
outf = torch.empty((n, 3, h, w), device=dev)
bl = torch.randn(3, generator=g, device=dev) * 0.1
lr64 = ops.new_feature(n, h, w, dev, zero=True)


def last():
    ops.conv3x3(x, wl, bias=bl, lq=lq, outf=outf)


def im2col():
    ops.lr_im2col(lq, lr64)          # (pass a (n,h,w,32) tensor for the compact operand)


def pair():
    launch_a()
    launch_b()


def timeit(fn, reps=60):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def sustained(fn, seconds=1.5):
    """us per call, median SM clock (MHz) and max power (W) over a loop long enough for the power cap to settle"""
    import subprocess
    us = timeit(fn, 200)
    reps = int(seconds * 1e6 / us)
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100",
                            "-i", "0"], stdout=subprocess.PIPE, text=True)
    us = timeit(fn, reps)
    smi.terminate()
    rows = [r.split(",") for r in smi.communicate()[0].strip().splitlines() if "," in r]
    clk = sorted(float(r[0]) for r in rows[len(rows) // 3:]) or [0.0]
    pw = [float(r[1]) for r in rows] or [0.0]
    return us, clk[len(clk) // 2], max(pw)


print(f"PNP_PAIR={os.environ.get('PNP_PAIR', '0')} shape {n}x{h}x{w}, resident CTA pairs {_lib.load().pnp_device_pairs()}")
launch_a()
for name, fn, res in (("plain", plain, out), ("launch A", launch_a, t), ("launch B", launch_b, out), ("A+B", pair, out),
                      ("conv_last", last, outf), ("lr_im2col", im2col, lr64)):
    us = timeit(fn)
    line = f"  {name:9s} {us:7.1f} us   checksum {res.float().abs().sum().item():.6e}"
    if os.environ.get("PNP_SUSTAINED", "0") != "0":
        us2, mhz, watts = sustained(fn)
        line += f"   sustained {us2:7.1f} us at {mhz:.0f} MHz, {watts:.0f} W max = {us2 * mhz / 1e3:.1f} kcycles"
    print(line)
