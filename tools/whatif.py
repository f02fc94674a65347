"""What-if timing of the conv kernel with parts of the pipeline disabled (PNP_DEBUG_SKIP bits)."""
import os, sys, subprocess
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from pnpvcve_b200 import ops
    import functools
    CONV = functools.partial(ops.conv3x3, wpack_stable=bool(os.environ.get("PNP_WSTABLE")))
    dev = torch.device("cuda:0"); h, w = 720, 1280
    x = torch.randn((1, h, w, 64), device=dev).to(torch.bfloat16)
    WS = float(os.environ.get("PNP_WSCALE", "0.05"))
    if os.environ.get("PNP_ZERO"): x.zero_()
    idt = torch.randn((1, h, w, 64), device=dev).to(torch.bfloat16)
    out = ops.new_feature(1, h, w, dev)
    wp9 = ops.new_wpack(9, dev); ops.pack_conv3x3(torch.randn((64, 64, 3, 3), device=dev) * WS, wp9)
    wp = ops.new_wpack(12, dev); ops.pack_conv3x3(torch.randn((64, 64, 3, 3), device=dev) * WS, wp, center_chunks=4)
    par = torch.rand((1, 3, h, w), device=dev)
    def timeit(fn, iters=30):
        for _ in range(5): fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters): fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / iters * 1e3
    a = timeit(lambda: CONV(x, wp9, out=out))
    b = timeit(lambda: CONV(x, wp9, out=out, idt=idt))
    c = timeit(lambda: CONV(x, wp, out=out, par=par, act=2))
    scale = torch.rand(64, device=dev) + 0.5; bias = torch.randn(64, device=dev)
    c2 = timeit(lambda: CONV(x, wp, out=out, par=par, scale=scale, bias=bias, act=2))
    par1 = torch.nn.functional.one_hot(torch.randint(0, 3, (1, h // 8, w // 8), device=dev), 3).permute(0, 3, 1, 2).float().repeat_interleave(8, 2).repeat_interleave(8, 3).contiguous() / 255
    c3 = timeit(lambda: CONV(x, wp, out=out, par=par1, scale=scale, bias=bias, act=2))
    print(f"tap-major +par+scale+bias {c2:6.1f} us ; with one-hot/255 partition map {c3:6.1f} us", flush=True)
    wr = ops.new_wpack_rowstack(dev); ops.pack_conv3x3_rowstack(torch.randn((64, 64, 3, 3), device=dev) * WS, wr)
    ra = timeit(lambda: CONV(x, wr, out=out, wlayout=1))
    rb = timeit(lambda: CONV(x, wr, out=out, idt=idt, wlayout=1))
    wrp = ops.new_wpack_rowstack(dev, with_par=True); ops.pack_conv3x3_rowstack(torch.randn((64, 64, 3, 3), device=dev) * WS, wrp)
    rc = timeit(lambda: CONV(x, wrp, out=out, par=par, act=2, wlayout=1))
    print(f"skip={os.environ.get('PNP_DEBUG_SKIP','0'):>2s}: rowstack plain {ra:6.1f} us   +id {rb:6.1f} us   +par {rc:6.1f} us", flush=True)
    print(f"skip={os.environ.get('PNP_DEBUG_SKIP','0'):>2s}: plain {a:6.1f} us   +id {b:6.1f} us   +par {c:6.1f} us", flush=True)
else:
    for bits in (0, 1, 2, 4, 3, 6, 7):
        env = dict(os.environ, PNP_DEBUG_SKIP=str(bits))
        subprocess.run([sys.executable, __file__, "run"], env=env)
