"""lr_im2col at 720p under a cold L2 (rotating buffers)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnpvcve_b200 import ops
dev = torch.device("cuda:0"); h, w = 720, 1280
lrs = [torch.rand((1, 3, h, w), device=dev) for _ in range(8)]
dsts = [ops.new_feature(1, h, w, dev, zero=True) for _ in range(4)]
for i in range(8): ops.lr_im2col(lrs[i % 8], dsts[i % 4])
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(200): ops.lr_im2col(lrs[i % 8], dsts[i % 4])
e.record(); torch.cuda.synchronize()
print(f"lr_im2col 720p: {s.elapsed_time(e) / 200 * 1e3:.1f} us per launch")
