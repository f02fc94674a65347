"""lr_im2col at 720p under a cold L2 (rotating buffers)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnpvcve_b200 import ops
dev = torch.device("cuda:0"); h, w = 720, 1280
lrs = [torch.rand((1, 3, h, w), device=dev) for _ in range(8)]
for ch in (64, 32):        # 128-byte pixels with dead upper halves / compact 64-byte pixels (the engine's operand)
    dsts = [torch.zeros((1, h, w, ch), dtype=torch.bfloat16, device=dev) for _ in range(8)]
    for i in range(8): ops.lr_im2col(lrs[i % 8], dsts[i % 8])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(200): ops.lr_im2col(lrs[i % 8], dsts[i % 8])
    e.record(); torch.cuda.synchronize()
    print(f"lr_im2col 720p, {ch}-channel destination pixels: {s.elapsed_time(e) / 200 * 1e3:.1f} us per launch")
