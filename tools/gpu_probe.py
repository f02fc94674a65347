"""Diagnostic probe of the tcgen05 conv kernel on a real B200 (not a test, not a benchmark).

Runs single-tap identity convolutions (output must equal the input shifted by the tap) under both
UMMA base-offset modes, then random convolutions of every epilogue variant against torch fp32
conv2d on the same bf16-rounded operands.  Prints one line per check.
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnpvcve_b200 import _lib, ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def bf(x):
    return x.to(torch.bfloat16).float()


def nhwc(x):  # (N,64,H,W) fp32 -> (N,H,W,64) bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x):  # (N,H,W,64) bf16 -> (N,64,H,W) fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


def report(name, got, ref, tol):
    d = (got - ref).abs()
    bad = int((d > tol).sum())
    print(f"  {name:<46s} max|d|={d.max().item():.3e} mean|d|={d.mean().item():.3e} "
          f"bad={bad}/{d.numel()} {'OK' if bad == 0 else 'FAIL'}", flush=True)
    return bad == 0


def run_conv(x, w, **kw):
    n, c, h, wd = x.shape
    wp = ops.new_wpack(12, dev)
    ops.pack_conv3x3(w.contiguous(), wp)
    out = ops.new_feature(n, h, wd, dev)
    ops.conv3x3(nhwc(x), wp, out=out, **kw)
    torch.cuda.synchronize()
    return nchw(out)


def tap_probe(mode, h, w):
    _lib.check(_lib.load().pnp_set_base_offset_mode(mode))
    g = torch.Generator(device=dev).manual_seed(1)
    x = bf(torch.randn((1, 64, h, w), generator=g, device=dev))
    ok_all = True
    for tap in range(9):
        wt = torch.zeros((64, 64, 3, 3), device=dev)
        wt[:, :, tap // 3, tap % 3] = torch.eye(64, device=dev)
        got = run_conv(x, wt)
        ref = F.conv2d(x, wt, padding=1)
        ok_all &= report(f"mode{mode} {h}x{w} identity tap {tap} (dy={tap // 3 - 1},dx={tap % 3 - 1})",
                         got, ref, 1e-6)
    return ok_all


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    _lib.require_device()
    results = {}
    for mode in (0, 1):
        try:
            results[mode] = tap_probe(mode, 64, 128)
        except Exception as e:  # noqa: BLE001
            print(f"mode {mode} raised: {e}", flush=True)
            results[mode] = False
    print("base-offset mode results:", results, flush=True)
    good = 0 if results.get(0) else (1 if results.get(1) else None)   # 0 = library default
    if good is None:
        print("NO base-offset mode reproduces shifted taps -- stop here", flush=True)
        return 1
    _lib.check(_lib.load().pnp_set_base_offset_mode(good))
    print(f"using base-offset mode {good}", flush=True)

    g = torch.Generator(device=dev).manual_seed(2)
    for (n, h, w) in [(1, 64, 64), (1, 68, 132), (2, 72, 200), (1, 180, 320), (1, 720, 1280)]:
        print(f"shape n={n} {h}x{w}", flush=True)
        x = bf(torch.randn((n, 64, h, w), generator=g, device=dev))
        wt = bf(torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.05)
        bias = torch.randn(64, generator=g, device=dev) * 0.1
        scale = torch.rand(64, generator=g, device=dev) + 0.5
        ref0 = F.conv2d(x, wt, padding=1)
        t0 = time.time()
        report("plain", run_conv(x, wt), bf(ref0), 2e-2)
        print(f"    ({time.time() - t0:.3f}s incl. pack)", flush=True)
        report("bias+lrelu", run_conv(x, wt, bias=bias, act=ops.PNP_ACT_LRELU),
               bf(F.leaky_relu(ref0 + bias.view(1, -1, 1, 1), 0.1)), 2e-2)
        idt = bf(torch.randn((n, 64, h, w), generator=g, device=dev))
        report("scale+bias+id+relu",
               run_conv(x, wt, bias=bias, scale=scale, idt=nhwc(idt), act=ops.PNP_ACT_RELU),
               bf(F.relu(ref0 * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1) + idt)), 3e-2)
        # aux: 3-channel LR im2col
        lr = torch.rand((n, 3, h, w), generator=g, device=dev)
        w_in = bf(torch.randn((64, 131, 3, 3), generator=g, device=dev) * 0.05)
        wp = ops.new_wpack(10, dev)
        ops.pack_conv3x3(w_in, wp, in_begin=3, in_count=64)
        ops.pack_aux(w_in, wp[9 * ops.CHUNK_BYTES:])
        lr64 = ops.new_feature(n, h, w, dev, zero=True)
        ops.lr_im2col(lr, lr64)
        out = ops.new_feature(n, h, w, dev)
        ops.conv3x3(nhwc(x), wp, out=out, aux=lr64, bias=bias, act=ops.PNP_ACT_LRELU)
        ref = F.conv2d(torch.cat([bf(lr), x], 1), w_in[:, :67], bias, padding=1)
        report("aux(lr)+bias+lrelu", nchw(out), bf(F.leaky_relu(ref, 0.1)), 3e-2)
        # par: 3x3 + three partition-modulated 1x1
        par = torch.rand((n, 3, h, w), generator=g, device=dev) * (torch.rand((n, 3, h, w), generator=g,
                                                                             device=dev) > 0.5)
        w1 = [bf(torch.randn((64, 64), generator=g, device=dev) * 0.1) for _ in range(3)]
        wp = ops.new_wpack(12, dev)
        ops.pack_conv3x3(wt, wp, center_chunks=4)
        for j in range(3):
            ops.pack_rows(w1[j], wp, 64 * (j + 1))
        out = ops.new_feature(n, h, w, dev)
        ops.conv3x3(nhwc(x), wp, out=out, scale=scale, bias=bias, par=par, act=ops.PNP_ACT_RELU)
        ref = ref0 * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
        for j in range(3):
            ref = ref + F.conv2d(x, w1[j].view(64, 64, 1, 1)) * par[:, j:j + 1]
        report("par(3x1x1)+scale+bias+relu", nchw(out), bf(F.relu(ref)), 3e-2)
        # last: 64 -> 3, + lq, fp32 NCHW out
        wl = bf(torch.randn((3, 64, 3, 3), generator=g, device=dev) * 0.05)
        bl = torch.randn(3, generator=g, device=dev) * 0.1
        wp = ops.new_wpack(9, dev)
        ops.pack_conv3x3(wl, wp)
        outf = torch.empty((n, 3, h, w), device=dev)
        ops.conv3x3(nhwc(x), wp, bias=bl, lq=lr, outf=outf)
        torch.cuda.synchronize()
        report("last(64->3)+bias+lq", outf, F.conv2d(x, wl, bl, padding=1) + lr, 1e-3)

    # timing of the main kernels at 720p (device time, CUDA events)
    h, w = 720, 1280
    x = nhwc(bf(torch.randn((1, 64, h, w), generator=g, device=dev)))
    out = ops.new_feature(1, h, w, dev)
    wp = ops.new_wpack(12, dev)
    ops.pack_conv3x3(bf(torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.05), wp, center_chunks=4)
    wp9 = ops.new_wpack(9, dev)
    ops.pack_conv3x3(bf(torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.05), wp9)
    par = torch.rand((1, 3, h, w), device=dev)
    idt = nhwc(torch.randn((1, 64, h, w), device=dev))
    flow = (torch.randint(-64, 65, (2, h // 8, w // 8), device=dev).float() / 4).repeat_interleave(
        8, 1).repeat_interleave(8, 2).contiguous()

    def timeit(fn, iters=20):
        for _ in range(3):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters * 1e3

    flops = 2 * 64 * 64 * 9 * h * w
    us = timeit(lambda: ops.conv3x3(x, wp9, out=out))
    print(f"720p conv64x64 plain      : {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
    us = timeit(lambda: ops.conv3x3(x, wp9, out=out, idt=idt))
    print(f"720p conv64x64 +id        : {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
    fl2 = flops + 2 * 3 * 64 * 64 * h * w
    us = timeit(lambda: ops.conv3x3(x, wp, out=out, par=par, act=ops.PNP_ACT_RELU))
    print(f"720p conv64x64 +par(1x1x3): {us:8.1f} us  {fl2 / us / 1e6:7.1f} TFLOP/s", flush=True)
    us = timeit(lambda: ops.mv_warp(x, flow, out))
    byt = h * w * (2 * 64 * 2 + 8)
    print(f"720p mv_warp              : {us:8.1f} us  {byt / us / 1e3:7.1f} GB/s", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
