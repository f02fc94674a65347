"""Tiny launch sequence for `ncu` captures at the REDS4 shape (720p): per round
   conv3x3_rows_kernel<1,0> (block launch A), conv3x3_rows_kernel<0,0> (block launch B: + identity, bottom-up), mv_warp.
Not a benchmark: numbers printed under a profiler are never bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnpvcve_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (720, 1280)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((1, h, w, 64), generator=g, device=dev).to(torch.bfloat16)
idt = torch.randn((1, h, w, 64), generator=g, device=dev).to(torch.bfloat16)
t = ops.new_feature(1, h, w, dev)
out = ops.new_feature(1, h, w, dev)
war = ops.new_wpack_rowstack(dev, with_par=True)          # launch A, row-stacked layout (the engine's default)
ops.pack_conv3x3_rowstack(torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.05, war)
for j in range(3):
    ops.pack_rows(torch.randn((64, 64), generator=g, device=dev) * 0.1, war[9 * ops.CHUNK_BYTES:], 64 * j)
wb = ops.new_wpack_rowstack(dev)
ops.pack_conv3x3_rowstack(torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.05, wb, flip_ky=True)
par = torch.rand((1, 3, h, w), generator=g, device=dev)
scale = torch.rand(64, generator=g, device=dev) + 0.5
bias = torch.randn(64, generator=g, device=dev) * 0.1
flow = (torch.randint(-64, 65, (2, (h + 7) // 8, (w + 7) // 8), generator=g, device=dev).float() / 4
        ).repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :h, :w].contiguous()
torch.cuda.synchronize()
for _ in range(4):
    ops.conv3x3(x, war, out=t, bias=bias, par=par, act=ops.PNP_ACT_RELU)  # block launch A (row-stacked)
    ops.conv3x3(t, wb, out=out, idt=x, bias=bias, flip_y=True)           # block launch B
    ops.mv_warp(x, flow, out)
torch.cuda.synchronize()
print("done")
