"""Quality metrics (PSNR / SSIM of the reference's test loop): oracle against the reference's own values on the
CPU, pnp_frame_quality against the oracle and those values on the GPU."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M

from helpers import GOLDEN, metric_cases

GOLD = np.load(os.path.join(GOLDEN, "metrics_cases.npz"))


@pytest.mark.parametrize("name", sorted(metric_cases()))
def test_metrics_oracle_matches_reference_values(name):
    out, gt, crop = metric_cases()[name]
    o8, g8 = M.tensor2img_u8(out.numpy()), M.tensor2img_u8(gt.numpy())
    assert int(o8.astype(np.int64).sum() * 1000003 + g8.astype(np.int64).sum()) == int(GOLD[name + "/u8sum"])
    p, s = M.psnr(o8, g8, crop), M.ssim(o8, g8, crop)
    if math.isinf(float(GOLD[name + "/psnr"])):
        assert math.isinf(p)
    else:
        assert abs(p - float(GOLD[name + "/psnr"])) <= 1e-5
    assert abs(s - float(GOLD[name + "/ssim"])) <= 1e-12
    # psnr from the exact integer error sum (what the GPU accumulates) is the same number
    if not math.isinf(p):
        n = o8[crop:o8.shape[0] - crop, crop:o8.shape[1] - crop].size
        assert abs(20 * math.log10(255 / math.sqrt(M.sse_u8(o8, g8, crop) / n)) - p) <= 1e-5


def test_tensor2img_rounds_half_to_even_and_clamps():
    x = torch.tensor([-0.3, 0.0, 0.5 / 255, 1.5 / 255, 2.5 / 255, 0.999, 1.0, 1.7]).view(1, 1, 8).repeat(3, 1, 1)
    u = M.tensor2img_u8(x.numpy())
    assert u[0, :, 0].tolist() == [0, 0, 0, 2, 2, 255, 255, 255]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(metric_cases()))
def test_frame_quality_kernel_matches_reference_values(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pnpvcve_b200 import metrics
    dev = torch.device("cuda:0")
    out, gt, crop = metric_cases()[name]
    q = metrics.frame_quality(out[None].to(dev), gt[None].to(dev), crop)
    o8, g8 = M.tensor2img_u8(out.numpy()), M.tensor2img_u8(gt.numpy())
    assert int(q["sse"].item()) == M.sse_u8(o8, g8, crop)                  # integer work: bit exact
    gp, gs = float(GOLD[name + "/psnr"]), float(GOLD[name + "/ssim"])
    if math.isinf(gp):
        assert math.isinf(q["psnr"].item())
    else:
        assert abs(q["psnr"].item() - gp) <= 1e-5
    assert abs(q["ssim"].item() - gs) <= 1e-9


@pytest.mark.gpu
def test_frame_quality_sequence_720p_views_and_evaluate():
    """(n,T,3,H,W) sequence at the REDS4 shape through strided views; evaluate() = mean over frames."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pnpvcve_b200 import metrics
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    big = torch.rand((1, 3, 3, 724, 1288), generator=g, device=dev)
    gt = big[..., 2:722, 4:1284]                                           # non-contiguous view, unit x stride
    out = (gt + 0.03 * torch.randn(gt.shape, generator=g, device=dev)).contiguous()
    q = metrics.frame_quality(out, gt, 0)
    assert q["psnr"].shape == (1, 3) and q["ssim"].dtype == torch.float64
    o8 = M.tensor2img_u8(out[0, 1].cpu().numpy())
    g8 = M.tensor2img_u8(gt[0, 1].cpu().numpy())
    assert int(q["sse"][0, 1].item()) == M.sse_u8(o8, g8)
    assert abs(q["psnr"][0, 1].item() - M.psnr(o8, g8)) <= 1e-5
    assert abs(q["ssim"][0, 1].item() - M.ssim(o8, g8)) <= 1e-9
    ev = metrics.evaluate(out, gt)
    assert abs(ev["PSNR"].item() - q["psnr"].mean().item()) < 1e-12
    with pytest.raises(Exception):
        metrics.frame_quality(out[..., :12, :12], gt[..., :12, :12], crop_border=1)   # cropped frame < 11 x 11


@pytest.mark.gpu
def test_driver_gathers_psnr_ssim_per_frame():
    """enhance_clips with ground truth: [max-abs, mse, PSNR, SSIM] per frame through the fixed-shape gather."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pnpvcve_b200 as P
    from pnpvcve_b200 import driver, metrics, synthetic, weights
    from test_gpu_parity import GENERATOR_CFG
    dev = torch.device("cuda:0")
    net = P.build_backbone(dict(GENERATOR_CFG, num_blocks=2))
    net.load_state_dict(weights.random_state_dict(3, num_blocks=2), strict=True)
    net = net.to(dev).eval()
    clips = [{k: v.to(dev) for k, v in synthetic.make_clip(64, 96, 3, seed=50 + i).items()} for i in range(2)]
    gts = [(c["lq"] * 0.9 + 0.05) for c in clips]
    outs, met = driver.enhance_clips(net, clips, gts=gts)
    assert met.shape == (2, 3, 4)
    q = metrics.frame_quality(outs[1], gts[1])
    assert torch.allclose(met[1, :, 2].double(), q["psnr"][0], atol=1e-4)
    assert torch.allclose(met[1, :, 3].double(), q["ssim"][0], atol=1e-6)
