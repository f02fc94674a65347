"""Replays the engine's block loop (launch A -> t -> launch B + identity, ping-pong buffers, per-block weights)
outside the engine and times it two ways: whole-loop time per block, and CUDA events around every 8th launch
(the bench's method).  Diagnostic: separates what the A->B->A sequence costs from what the engine adds."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pnpvcve_b200 import ops
dev = torch.device("cuda:0"); h, w = 720, 1280
NB = 16
ROWS_PAR = os.environ.get("PNP_ROWS_PAR", "0") != "0"
g = torch.Generator(device=dev).manual_seed(0)
xa = torch.randn((1, h, w, 64), generator=g, device=dev).to(torch.bfloat16)
xb = ops.new_feature(1, h, w, dev); t = ops.new_feature(1, h, w, dev)
pars = [(torch.nn.functional.one_hot(torch.randint(0, 3, (1, h // 8, w // 8), device=dev), 3).permute(0, 3, 1, 2).float()
         .repeat_interleave(8, 2).repeat_interleave(8, 3).contiguous() / 255) for _ in range(4)]
bias = [torch.randn(64, generator=g, device=dev) * 0.01 for _ in range(NB)]
wa, wb = [], []
for k in range(NB):
    w3 = torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.02
    if ROWS_PAR:
        b = ops.new_wpack_rowstack(dev, with_par=True); ops.pack_conv3x3_rowstack(w3, b)
        for j in range(3): ops.pack_rows(torch.randn((64, 64), generator=g, device=dev) * 0.05, b[9 * ops.CHUNK_BYTES:], 64 * j)
    else:
        b = ops.new_wpack(12, dev); ops.pack_conv3x3(w3, b, center_chunks=4)
        for j in range(3): ops.pack_rows(torch.randn((64, 64), generator=g, device=dev) * 0.05, b, 64 * (j + 1))
    wa.append(b)
    b = ops.new_wpack_rowstack(dev); ops.pack_conv3x3_rowstack(torch.randn((64, 64, 3, 3), generator=g, device=dev) * 0.02, b, flip_ky=True)
    wb.append(b)
torch.cuda.synchronize()
MODE = sys.argv[1] if len(sys.argv) > 1 else "ab"

def loop(frames, every):
    evs = {"a": [], "b": []}
    cnt = 0
    x, o = xa, xb
    for f in range(frames):
        par = pars[f % 4]
        for k in range(NB):
            ta = every and cnt % every == 0
            if ta:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
            if MODE != "b_only":
                ops.conv3x3(x, wa[k], out=t, bias=bias[k], par=par, act=ops.PNP_ACT_RELU, wlayout=1 if ROWS_PAR else 0, wpack_stable=True)
            if ta:
                e1.record(); evs["a"].append((e0, e1))
            tb = every and cnt % every == 4
            if tb:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
            if MODE != "a_only":
                ops.conv3x3(t, wb[k], out=o, idt=x, bias=bias[k], wlayout=1, flip_y=True, wpack_stable=True)
            if tb:
                e1.record(); evs["b"].append((e0, e1))
            x, o = o, x
            cnt += 1
    return evs

for every in (0, 8, 1):
    loop(2, every); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); evs = loop(12, every); e.record(); torch.cuda.synchronize()
    per_block = s.elapsed_time(e) / (12 * NB) * 1e3
    ma = sum(a.elapsed_time(b) for a, b in evs["a"]) / max(len(evs["a"]), 1) * 1e3
    mb = sum(a.elapsed_time(b) for a, b in evs["b"]) / max(len(evs["b"]), 1) * 1e3
    print(f"mode {MODE} rows_par {int(ROWS_PAR)} events every {every}: {per_block:6.1f} us per block   bracketed A {ma:6.1f} us  B {mb:6.1f} us")
