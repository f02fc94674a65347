"""Seeded synthetic clips with the value ranges the reference's data pipeline produces.

Follows SURVEY.md section 8(d): the reference's loaders (mmedit/datasets/pipelines/
loading_ipb.py:221-397, normalization.py:93-99, formating.py:101-138) hand the generator
    lq        (n,T,3,H,W)   fp32 in [0,1]
    QPs       (n,T,1,1,1)   frame QP / 255, or ord(slice type) / 255 in the IPB configs
    slices    (n,T,1,1,1)   ord('I'|'P'|'B') = 73 | 80 | 66 (not rescaled)
    mvs       (n,T,4,H,W)   pixels; ch0,1 = forward x,y ; ch2,3 = backward x,y; block constant
    base_QPs  (n,T,1,1,1)   CRF / 255
    partitions(n,T,3,H,W)   one-hot block-size mask scaled by 1/255
There is no dataset offline, so clips are synthetic; one seed per clip = 1000*config + clip.
"""
import torch

SLICE_I, SLICE_P, SLICE_B = 73, 80, 66

#: name -> (config number, H, W, default T, MV range in quarter pels, IPB conditioning)
CONFIGS = {
    "C1": dict(idx=1, h=128, w=128, t=7, mv_qpel=32, ipb=False),      # HR_davis_LR_128x128, CPU case
    "C2": dict(idx=2, h=720, w=1280, t=100, mv_qpel=64, ipb=False),   # REDS4 shape, CRF 15/25/35
    "C3": dict(idx=3, h=720, w=1280, t=100, mv_qpel=64, ipb=True),    # _IPB, clip-sharded
    "C4": dict(idx=4, h=180, w=320, t=100, mv_qpel=32, ipb=True),     # _IPB_LR_test, many clips
    "C5": dict(idx=5, h=376, w=1244, t=2, mv_qpel=64, ipb=True),      # KITTI pairs pre-padded to x4
}


def gop_slices(t, pattern="IBBP"):
    """``I (B B P)*`` truncated to t frames (C5 uses ``I P``)."""
    if pattern == "IP":
        seq = [SLICE_I] + [SLICE_P] * (t - 1)
    elif pattern == "allB":
        seq = [SLICE_B] * t
    elif pattern == "allkey":
        seq = [SLICE_I] + [SLICE_P] * (t - 1)
    else:
        seq = [SLICE_I]
        while len(seq) < t:
            seq += [SLICE_B, SLICE_B, SLICE_P]
    return seq[:t]


def make_clip(h, w, t, seed, crf=25, mv_qpel=64, ipb=False, pattern="IBBP", device="cpu",
              block=8):
    """One clip (n=1).  Returns dict(lq, QPs, slices, mvs, base_QPs, partitions)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    lq = torch.rand((1, t, 3, h, w), generator=g, device=dev, dtype=torch.float32)
    sl = gop_slices(t, pattern)
    slices = torch.tensor(sl, dtype=torch.float32, device=dev).view(1, t, 1, 1, 1)
    hb, wb = (h + block - 1) // block, (w + block - 1) // block
    mv_b = torch.randint(-mv_qpel, mv_qpel + 1, (1, t, 4, hb, wb), generator=g, device=dev)
    mv_b = mv_b.to(torch.float32) / 4.0
    mvs = mv_b.repeat_interleave(block, dim=3).repeat_interleave(block, dim=4)[..., :h, :w]
    pidx = torch.randint(0, 3, (1, t, hb, wb), generator=g, device=dev)
    par_b = torch.nn.functional.one_hot(pidx, 3).permute(0, 1, 4, 2, 3).to(torch.float32) / 255.0
    par = par_b.repeat_interleave(block, dim=3).repeat_interleave(block, dim=4)[..., :h, :w]
    is_i = (slices == SLICE_I).view(1, t, 1, 1, 1)
    mvs = torch.where(is_i, torch.zeros_like(mvs), mvs).contiguous()
    par = torch.where(is_i, torch.zeros_like(par), par).contiguous()
    base = torch.full((1, t, 1, 1, 1), crf / 255.0, dtype=torch.float32, device=dev)
    if ipb:
        qps = slices / 255.0
    else:
        q = torch.randint(crf - 5, crf + 10, (1, t, 1, 1, 1), generator=g, device=dev)
        qps = q.to(torch.float32) / 255.0
    return dict(lq=lq, QPs=qps, slices=slices, mvs=mvs, base_QPs=base, partitions=par)


def overlap_partitions(clip, seed, block=8):
    """Make the partition maps overlap and carry other values than 1/255 (in place): the dense path uses the
    VALUES, the reference's sparse_val path only whether they are non-zero and which class comes last."""
    par = clip["partitions"]
    n, t, _, h, w = par.shape
    g = torch.Generator().manual_seed(int(seed))
    hb, wb = (h + block - 1) // block, (w + block - 1) // block
    extra = (torch.rand((n, t, 3, hb, wb), generator=g) > 0.6).float() * \
        torch.randint(1, 4, (n, t, 3, hb, wb), generator=g).float() / 255.0
    extra = extra.repeat_interleave(block, dim=3).repeat_interleave(block, dim=4)[..., :h, :w]
    clip["partitions"] = (par + extra.to(par.device)).contiguous()
    return clip


def make_config_clip(name, clip_idx=0, t=None, crf=25, device="cpu", pattern=None):
    c = CONFIGS[name]
    pat = pattern or ("IP" if name == "C5" else "IBBP")
    return make_clip(c["h"], c["w"], t or c["t"], 1000 * c["idx"] + clip_idx, crf=crf,
                     mv_qpel=c["mv_qpel"], ipb=c["ipb"], pattern=pat, device=device)


def cat_clips(clips):
    """Batch several equally shaped clips along n."""
    return {k: torch.cat([c[k] for c in clips], dim=0) for k in clips[0]}


def generator_args(clip):
    """Positional argument order of the generator call (basicvsr.py:179)."""
    return (clip["lq"], clip["QPs"], clip["slices"], clip["mvs"], clip["base_QPs"],
            clip["partitions"])
