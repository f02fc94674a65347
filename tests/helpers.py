"""Shared helpers for the parity tests (rebuild golden-case inputs without the reference)."""
import json
import os

import numpy as np
import torch

from pnpvcve_b200 import synthetic, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)


def build_case(name):
    """(state_dict, clip dict, golden npz) for a committed golden case."""
    case = golden_cases()[name]
    clips = [synthetic.make_clip(**kw) for kw in case["clips"]]
    clip = synthetic.cat_clips(clips)
    if case.get("par_overlap"):
        synthetic.overlap_partitions(clip, case["par_overlap"])
    if case.get("mirror"):
        t = clip["lq"].shape[1]
        half = clip["lq"][:, : t // 2]
        clip["lq"] = torch.cat([half, half.flip(1)], dim=1).contiguous()
    sd = weights.random_state_dict(case["weight_seed"], num_blocks=case["num_blocks"], vsr=bool(case.get("vsr")))
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    return sd, clip, gold


def residual_base(lq, vsr):
    """The frame the generator adds its conv output to: the reflect-padded LR frame (iconvsr.py:371-394,
    iconvsr_ipb_par.py:146) or, with vsr, its x4 bilinear upsampling (:41,140-141)."""
    n, t, c, h, w = lq.shape
    ph, pw = (4 - h % 4) % 4, (4 - w % 4) % 4
    x = torch.nn.functional.pad(lq.reshape(-1, c, h, w), [0, pw, 0, ph], mode="reflect")
    if vsr:
        x = torch.nn.functional.interpolate(x, scale_factor=4, mode="bilinear", align_corners=False)
    return x.view(n, t, c, *x.shape[-2:])


def check_against_golden(out, gold, tol, lq=None, vsr=False):
    """max-abs error on the stored fp32 lattice, on every other pixel (fp16 residuals: + 4e-5 of quantisation) when
    ``lq`` is given, and relative error of the float64 frame sums."""
    out = out.detach().float().cpu()
    assert tuple(out.shape) == tuple(int(v) for v in gold["shape"])
    lat = torch.from_numpy(gold["lattice"])
    err = (out[..., ::2, ::2] - lat).abs().max().item()
    assert err <= tol, f"lattice max-abs {err:.3e} > {tol:.1e}"
    if lq is not None:
        res = out - residual_base(lq.float().cpu(), vsr)
        off = torch.from_numpy(gold["off_lattice_f16"].astype(np.float32))
        got = torch.stack([res[..., 0::2, 1::2], res[..., 1::2, 0::2], res[..., 1::2, 1::2]])
        err_off = (got - off).abs().max().item()
        assert err_off <= tol + 4e-5, f"off-lattice max-abs {err_off:.3e} > {tol:.1e} + 4e-5"
        err = max(err, err_off - 4e-5)
    fs = out.double().sum(dim=(2, 3, 4)).numpy()
    npix = out.shape[2] * out.shape[3] * out.shape[4]
    sum_err = np.abs(fs - gold["frame_sum"]).max() / npix
    assert sum_err <= tol, f"frame mean error {sum_err:.3e} > {tol:.1e}"
    return err


def warp_case():
    g = torch.Generator().manual_seed(4242)
    x = torch.randn((1, 2, 720, 1280), generator=g)
    flow_b = torch.randint(-64, 65, (1, 2, 90, 160), generator=g).float() / 4.0
    flow = flow_b.repeat_interleave(8, 2).repeat_interleave(8, 3)
    gold = np.load(os.path.join(GOLDEN, "warp_720p.npz"))
    return x, flow, gold


def metric_cases():
    """name -> (out (3,H,W) fp32, gt (3,H,W) fp32, crop_border) for the PSNR / SSIM golden values
    (tests/golden/metrics_cases.npz, written by tests/golden/make_golden_metrics.py from the reference's functions)."""
    cases = {}
    for name, (h, w, seed, noise, crop) in dict(a_64x80=(64, 80, 1, 0.02, 0), b_72x96_crop4=(72, 96, 2, 0.05, 4),
                                                c_40x56_hard=(40, 56, 3, 0.3, 2), d_equal=(32, 48, 4, 0.0, 0)).items():
        g = torch.Generator().manual_seed(seed)
        gt = torch.rand((3, h, w), generator=g) * 1.1 - 0.05
        gt = torch.nn.functional.avg_pool2d(gt[None], 3, 1, 1)[0] * 1.2 - 0.05          # some structure
        out = gt + noise * torch.randn((3, h, w), generator=g)
        cases[name] = (out.contiguous(), gt.contiguous(), crop)
    return cases
