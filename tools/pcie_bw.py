import torch, time
dev = torch.device("cuda:0")
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4): fn()
    torch.cuda.synchronize()
    print(f"{name}: {4 * (1 << 30) / (time.perf_counter() - t0) / 1e9:.1f} GB/s (pinned, 1 GiB copies)")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); d2 = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"both directions at once: {4 * (1 << 30) / dt / 1e9:.1f} GB/s each way")
