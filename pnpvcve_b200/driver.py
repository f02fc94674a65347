"""Clip-sharded multi-GPU driver: one process per GPU, no data-path collective.

The reference shards its test set by rank-strided sampling (mmedit/datasets/samplers/
distributed_sampler.py:51-72) and gathers pickled results with two all_gathers
(mmedit/apis/test.py:190-234).  Clips are independent, so here clip ``c`` goes to rank
``c mod world`` and the only communication is ONE fixed-shape all_gather of per-frame metrics.
"""
import torch
import torch.distributed as dist

N_METRICS = 2  # per frame: max-abs error vs. a reference frame (or 0), mean squared error


def shard_clips(num_clips, rank, world):
    """Rank-strided clip indices (clip c -> rank c % world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    return list(range(rank, num_clips, world))


def clips_per_rank(num_clips, world):
    return (num_clips + world - 1) // world


def frame_metrics(out, ref=None):
    """out (n,T,3,H,W) -> (n,T,N_METRICS) fp32 on the same device (no host sync)."""
    if ref is None:
        ref = torch.zeros_like(out)
    d = (out - ref).float()
    return torch.stack([d.abs().amax(dim=(2, 3, 4)), (d * d).mean(dim=(2, 3, 4))], dim=-1)


def gather_metrics(local, num_clips, rank, world, group=None):
    """local: (len(shard_clips), T, N_METRICS) -> (num_clips, T, N_METRICS) in clip order on every rank.

    One all_gather of a fixed-shape tensor (ranks with fewer clips pad with NaN rows).
    """
    per = clips_per_rank(num_clips, world)
    t = local.shape[1] if local.dim() == 3 else 0
    nm = local.shape[2] if local.dim() == 3 else N_METRICS      # any per-frame metric vector (e.g. + PSNR, SSIM)
    padded = torch.full((per, t, nm), float("nan"), dtype=torch.float32, device=local.device)
    padded[: local.shape[0]] = local
    if world == 1:
        parts = [padded]
    else:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
    full = torch.empty((num_clips, t, nm), dtype=torch.float32, device=local.device)
    for r in range(world):
        idx = shard_clips(num_clips, r, world)
        if idx:
            full[idx] = parts[r][: len(idx)]
    return full


@torch.no_grad()
def enhance_clips(net, clips, rank=0, world=1, refs=None, gts=None, crop_border=0):
    """Run this rank's share of ``clips`` (list of dicts as produced by pnpvcve_b200.synthetic).

    Returns (outputs for the local clips, gathered metrics for all clips).  With ``gts`` (ground-truth clips,
    (1,T,3,H,W) each) the per-frame metric vector is [max-abs, mse, PSNR, SSIM]: PSNR / SSIM as BasicVSR.evaluate
    computes them (mmedit/models/restorers/basicvsr.py:119-153), but on the device (pnpvcve_b200.metrics) -- the
    reference's test loop copies every frame to the host and pickles per-clip results through two all_gathers
    (mmedit/apis/test.py:190-234); here nothing leaves the GPU before the single fixed-shape gather.
    """
    from .synthetic import generator_args
    from . import metrics as _metrics
    mine = shard_clips(len(clips), rank, world)
    outs, mets = [], []
    for c in mine:
        out = net(*generator_args(clips[c]))
        outs.append(out)
        m = frame_metrics(out, None if refs is None else refs[c])[0]
        if gts is not None:
            q = _metrics.frame_quality(out, gts[c].to(out.device), crop_border)
            m = torch.cat([m, q["psnr"][0].float()[:, None], q["ssim"][0].float()[:, None]], dim=1)
        mets.append(m)
    t = clips[0]["lq"].shape[1]
    dev = clips[0]["lq"].device
    nm = N_METRICS + (2 if gts is not None else 0)
    local = torch.stack(mets, 0) if mets else torch.empty((0, t, nm), device=dev)
    return outs, gather_metrics(local, len(clips), rank, world)
