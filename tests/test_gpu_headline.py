"""Parity at the HEADLINE lengths and shapes (BASELINE.json configs 2-4): the residual stream and the recurrent
features are stored in bf16 through 2 x T dependent frame steps x 16 blocks, so the short-clip goldens do not
prove the 100-frame configs.  The checker is the oracle restatement (oracle/bae_oracle.py, plain PyTorch) run on
the GPU in fp32 with TF32 off; it is first held to the CPU oracle on a small clip.

Per-frame max-abs error and PSNR delta are written to gpurun_out/r02_error_vs_frame.json (copied to profiles/).
north_star tolerance: max-abs 2e-3 on [0,1] frames, PSNR delta <= 0.02 dB -- asserted PER FRAME.
"""
import json
import os

import pytest
import torch

from oracle import bae_oracle as O
from pnpvcve_b200 import synthetic, weights

from test_gpu_parity import TOL, build, run

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def fp32_oracle():
    """The GPU oracle must be real fp32: no TF32 in cuDNN convolutions or matmuls."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def gpu_oracle(sd, clip, dev):
    with torch.no_grad():
        ref = O.generator_forward(sd, *[a.to(dev) for a in synthetic.generator_args(clip)])
    torch.cuda.synchronize()
    return ref


def per_frame_errors(out, ref, seed=0):
    """(n,T) max-abs and PSNR delta (tensor2img uint8 + psnr of the reference's test loop, against a synthetic ground
    truth near the reference output), computed on the device."""
    n, t = out.shape[:2]
    g = torch.Generator(device=out.device).manual_seed(seed)
    err = (out - ref).abs().amax(dim=(2, 3, 4))
    gt = (ref + 0.02 * torch.randn(ref.shape, generator=g, device=ref.device)).clamp(0, 1)
    q = lambda x: (x.clamp(0, 1) * 255.0).round()

    def psnr(a):
        mse = ((q(a) - q(gt)).double() ** 2).mean(dim=(2, 3, 4))
        return 20.0 * torch.log10(255.0 / mse.sqrt())
    return err.cpu(), (psnr(out) - psnr(ref)).abs().cpu()


def record(name, err, dps, extra=None):
    path = os.path.join(ROOT, "gpurun_out", "r02_error_vs_frame.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = {}
    if os.path.isfile(path):
        with open(path) as f:
            data = json.load(f)
    t = err.shape[1]
    pick = sorted({0, t // 4, t // 2, (3 * t) // 4, t - 1})
    data[name] = dict(frames=t, clips=err.shape[0], max_abs_all=float(err.max()), psnr_delta_all=float(dps.max()),
                      max_abs_per_frame=[round(float(v), 7) for v in err.amax(0)],
                      psnr_delta_per_frame=[round(float(v), 6) for v in dps.amax(0)],
                      at_frames={str(i): dict(max_abs=float(err[:, i].max()), psnr_delta=float(dps[:, i].max()))
                                 for i in pick}, **(extra or {}))
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
    print(f"{name}: max-abs {float(err.max()):.2e} psnr-delta {float(dps.max()):.4f} dB; at frames "
          + ", ".join(f"{i}: {float(err[:, i].max()):.2e}" for i in pick))


def test_gpu_oracle_equals_cpu_oracle(dev):
    """The checker used below (oracle on cuda, fp32, TF32 off) against the pinned CPU oracle."""
    sd = weights.random_state_dict(5)
    clip = synthetic.make_clip(64, 96, 5, seed=77, crf=35)
    ref_cpu = O.generator_forward(sd, *synthetic.generator_args(clip))
    ref_gpu = gpu_oracle(sd, clip, dev).cpu()
    assert (ref_cpu - ref_gpu).abs().max().item() <= 2e-5


def test_c2_720p_t13_crf35_random_qps(dev):
    """C2 (non-IPB: per-frame random QPs -> one expert-mixed pack per distinct (CRF, QP)), 1280x720, T=13
    I B B P B B P B B P B B P: five real key frames, CRF 35."""
    sd = weights.random_state_dict(21)
    clip = synthetic.make_config_clip("C2", clip_idx=5, t=13, crf=35)
    assert len({float(q) for q in clip["QPs"].flatten()}) >= 5
    ref = gpu_oracle(sd, clip, dev)
    net = build(sd, dev)
    out = run(net, clip, dev)
    err, dps = per_frame_errors(out, ref)
    record("c2_720p_t13_crf35", err, dps, dict(distinct_qps=len({float(q) for q in clip["QPs"].flatten()})))
    assert err.max().item() <= TOL
    assert dps.max().item() <= 0.02


def test_c3_720p_t12_ipb_crf15(dev):
    """C3 (IPB conditioning), 1280x720, T=12, CRF 15."""
    sd = weights.random_state_dict(22)
    clip = synthetic.make_config_clip("C3", clip_idx=6, t=12, crf=15)
    ref = gpu_oracle(sd, clip, dev)
    out = run(build(sd, dev), clip, dev)
    err, dps = per_frame_errors(out, ref)
    record("c3_720p_t12_ipb_crf15", err, dps)
    assert err.max().item() <= TOL
    assert dps.max().item() <= 0.02


def test_c4_lr_t100_batched_mixed_crf(dev):
    """C4 at its full length: 320x180, T=100, n=4 clips with DIFFERENT CRFs in one call (the many-clip LR workload):
    does the bf16 recurrent stream drift over 100 frames?  Error asserted per frame, table recorded."""
    sd = weights.random_state_dict(23)
    clips = [synthetic.make_config_clip("C4", clip_idx=k, crf=(15, 25, 35, 25)[k]) for k in range(4)]
    clip = synthetic.cat_clips(clips)
    ref = gpu_oracle(sd, clip, dev)
    net = build(sd, dev)
    out = run(net, clip, dev)
    err, dps = per_frame_errors(out, ref)
    record("c4_lr_t100_n4_mixed_crf", err, dps, dict(gpu_launches=net.gpu_launches))
    assert err.max().item() <= TOL
    assert dps.max().item() <= 0.02
    # no drift: the last quarter of the clip is not worse than 2x the first quarter
    q = err.shape[1] // 4
    assert err[:, -q:].max().item() <= 2.0 * err[:, :q].max().item() + 1e-4


def test_c2_t100_strip_no_drift(dev):
    """The headline length of C2/C3 on a 1280-wide strip (128 rows): T=100 frames, non-IPB random QPs, CRF 25 -- the
    same 2 x 100 dependent frame steps as the bench workload at a height the fp32 oracle finishes in seconds."""
    sd = weights.random_state_dict(24)
    clip = synthetic.make_clip(128, 1280, 100, seed=2100, crf=25, mv_qpel=64, ipb=False)
    ref = gpu_oracle(sd, clip, dev)
    out = run(build(sd, dev), clip, dev)
    err, dps = per_frame_errors(out, ref)
    record("c2_strip_1280x128_t100_crf25", err, dps)
    assert err.max().item() <= TOL
    assert dps.max().item() <= 0.02


def test_pytorch_on_the_same_gpu_is_recorded_beside_the_cuda_path(dev):
    """The reference is a PyTorch model: on a GPU box its users run it through ATen / cuDNN.  This records what the
    reference's algorithm (the oracle restatement, op for op) reaches on THIS B200 at 720p -- fp32 as the reference
    runs it, with TF32 allowed, and under bf16 autocast (our arithmetic type) -- next to the CUDA path on the same clip,
    in gpurun_out/r02_error_vs_frame.json.  Informational; the only assertion is that the hand-written path is not
    slower than the library path it replaces."""
    import time
    sd = weights.random_state_dict(25)
    t = 6
    clip = synthetic.make_config_clip("C2", clip_idx=9, t=t, crf=25)
    args = [a.to(dev) for a in synthetic.generator_args(clip)]
    sd_dev = {k: v.to(dev) for k, v in sd.items()}

    def timed(fn, reps=2):
        fn()                                             # warm-up (cuDNN algorithm selection, graph capture)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return t * reps / (time.perf_counter() - t0)

    def oracle_fn():
        with torch.no_grad():
            return O.generator_forward(sd_dev, *args)

    res = {}
    res["torch_fp32_frames_per_s"] = timed(oracle_fn)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    res["torch_tf32_frames_per_s"] = timed(oracle_fn)

    def autocast_fn():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return oracle_fn()
    try:
        res["torch_bf16_autocast_frames_per_s"] = timed(autocast_fn)
    except RuntimeError as e:                            # (an op of the restatement that autocast cannot mix)
        res["torch_bf16_autocast_error"] = str(e).splitlines()[0][:200]
    net = build(sd, dev)

    def ours():
        with torch.no_grad():
            return net(*args)
    res["cuda_path_frames_per_s"] = timed(ours, reps=4)
    res["clip"] = f"C2 1280x720, T={t} (short clip: key frames at both ends), CRF 25, resident inputs"
    path = os.path.join(ROOT, "gpurun_out", "r02_error_vs_frame.json")
    data = {}
    if os.path.isfile(path):
        with open(path) as f:
            data = json.load(f)
    data["pytorch_on_this_gpu"] = res
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
    print("720p frames/s on this GPU: " + ", ".join(f"{k} {v:.1f}" for k, v in res.items() if k.endswith("_per_s")))
    assert res["cuda_path_frames_per_s"] >= max(v for k, v in res.items() if k.startswith("torch_") and k.endswith("_per_s"))


def test_generator_on_non_current_device_stream(dev):
    """The launches follow the INPUT's device and the caller's current stream there: run from a side
    stream and, when the box has a second GPU, on a device that is not current."""
    sd = weights.random_state_dict(3, num_blocks=2)
    clip = synthetic.make_clip(64, 96, 4, seed=31, crf=25)
    net = build(sd, dev, num_blocks=2)
    base = run(net, clip, dev).clone()
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        again = run(net, clip, dev)
    side.synchronize()
    assert torch.equal(base, again)
    if torch.cuda.device_count() > 1:
        dev1 = torch.device("cuda:1")
        net1 = build(sd, dev1, num_blocks=2)
        assert torch.cuda.current_device() == 0
        out1 = run(net1, clip, dev1)
        torch.cuda.synchronize(dev1)
        assert torch.equal(base.cpu(), out1.cpu())


def test_sparse_val_only_in_eval_mode(dev):
    """sr_backbone_utils.py:307 takes the sparse path only when `not self.training`: a module left in train() under
    no_grad computes the DENSE blend."""
    sd = weights.random_state_dict(4, num_blocks=2)
    clip = synthetic.make_clip(64, 64, 3, seed=41, crf=25)
    synthetic.overlap_partitions(clip, 5)
    sparse_net = build(sd, dev, num_blocks=2, sparse_val=True)
    dense_net = build(sd, dev, num_blocks=2, sparse_val=False)
    dense = run(dense_net, clip, dev)
    sparse = run(sparse_net, clip, dev)
    assert not torch.equal(dense, sparse)
    sparse_net.train()
    assert torch.equal(run(sparse_net, clip, dev), dense)


def test_weight_update_through_data_invalidates_packs(dev):
    """Packed weights follow load_state_dict / .to() / p.data re-seating; in-place .data writes need invalidate()."""
    sd = weights.random_state_dict(6, num_blocks=2)
    sd2 = weights.random_state_dict(7, num_blocks=2)
    clip = synthetic.make_clip(64, 64, 2, seed=51, crf=25)
    net = build(sd, dev, num_blocks=2)
    a = run(net, clip, dev).clone()
    net.load_state_dict(sd2)
    b = run(net, clip, dev).clone()
    assert not torch.equal(a, b)
    with torch.no_grad():
        for k, p in net.named_parameters():
            p.data.copy_(sd[k])
    net.invalidate_packed_weights()
    assert torch.equal(run(net, clip, dev), a)
