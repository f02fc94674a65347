"""Which operand of block launch A / B is sensitive to a cold L2?  Rotates the partition map and/or the
source / destination tensors over enough buffers to exceed the 126 MB L2 (diagnostic)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnpvcve_b200 import ops
dev = torch.device("cuda:0"); h, w = 720, 1280
NX, NP = 6, 16
xs = [torch.randn((1, h, w, 64), device=dev).to(torch.bfloat16) for _ in range(NX)]
outs = [ops.new_feature(1, h, w, dev) for _ in range(NX)]
pars = [torch.rand((1, 3, h, w), device=dev) for _ in range(NP)]
bias = torch.randn(64, device=dev)
wp = ops.new_wpack(12, dev); ops.pack_conv3x3(torch.randn((64, 64, 3, 3), device=dev) * 0.05, wp, center_chunks=4)
wr = ops.new_wpack_rowstack(dev); ops.pack_conv3x3_rowstack(torch.randn((64, 64, 3, 3), device=dev) * 0.05, wr)
def timeit(fn, iters=48):
    for i in range(6): fn(i)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for i in range(iters): fn(i)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
for name, fx, fp in (("all hot", lambda i: 0, lambda i: 0), ("par cold", lambda i: 0, lambda i: i % NP),
                     ("x/out cold", lambda i: i % NX, lambda i: 0), ("all cold", lambda i: i % NX, lambda i: i % NP)):
    a = timeit(lambda i: ops.conv3x3(xs[fx(i)], wp, out=outs[fx(i)], par=pars[fp(i)], bias=bias, act=2, wpack_stable=True))
    print(f"launch A (tap-major +par+bias+relu)  {name:12s} {a:6.1f} us")
for name, fx in (("all hot", lambda i: 0), ("x/idt/out cold", lambda i: i % NX)):
    b = timeit(lambda i: ops.conv3x3(xs[fx(i)], wr, out=outs[fx(i)], idt=xs[(fx(i) + 1) % NX], bias=bias, wlayout=1, flip_y=True, wpack_stable=True))
    print(f"launch B (rows +id+bias, flip_y)     {name:14s} {b:6.1f} us")
