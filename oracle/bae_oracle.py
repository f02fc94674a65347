"""CPU restatement of the reference's BAE+CAA generator forward (the ORACLE).

TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this file; the
product package ``pnpvcve_b200`` never does (it fails loudly without its CUDA library).

Parity status: PINNED against the reference itself.  The reference ships no tests
or golden vectors, so the pin is: ``tests/golden/make_golden.py``
imports the unmodified reference files from ``/root/reference`` (``oracle/refshim.py``),
runs them on seeded synthetic clips and commits the outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks this restatement against those vectors (and, when the
reference tree is present, against the live reference).

Everything is plain PyTorch fp32 and follows the reference's operation ORDER so that the
result is reproduced to ~1 ulp.  Third-party arithmetic the reference relies on lives in
PyTorch (``F.conv2d``, ``F.grid_sample``, ``nn.Linear`` ...; the reference pins torch 1.8 via
its docker tag, README.md:57-61) -- ``warp_bilinear`` below restates ATen's
``grid_sampler_2d`` (bilinear, zeros, align_corners=True) explicitly so that the integer tap
indices are observable.

All ``file:line`` citations are relative to ``/root/reference``.
"""
import torch
import torch.nn.functional as F

NUM_BLOCKS = 8
MID = 64
NUM_EXPERTS = 6


# --------------------------------------------------------------------------------------
# CAA heads
# --------------------------------------------------------------------------------------
def base_predictor(sd, base_qps):
    """Expert weights from CRF.  mmedit/models/backbones/sr_backbones/domain_aware.py:172-183

    ``softmax(Linear(64,6)(relu(Linear(1,64)(crf))))`` applied to ``CRFs.view(-1,1)``.
    base_qps: (n,t,1,1,1) -> (n,t,6)
    """
    n, t = base_qps.shape[:2]
    x = base_qps.reshape(-1, 1)
    h = F.relu(F.linear(x, sd["BasePredictor.BaseNet.0.weight"], sd["BasePredictor.BaseNet.0.bias"]))
    o = F.linear(h, sd["BasePredictor.BaseNet.2.weight"], sd["BasePredictor.BaseNet.2.bias"])
    return torch.softmax(o, dim=1).view(n, t, -1)


def se_module(sd, qps):
    """Per-frame 64-channel SE gain.  domain_aware.py:210-222, Hsigmoid :201-207

    ``relu6(Linear(4,64,nobias)(relu(Linear(1,4,nobias)(qp))) + 3) / 3`` (note: /3, range [0,2]).
    qps: (n,t,1,1,1) -> (n,t,64)
    """
    n, t = qps.shape[:2]
    x = qps.reshape(-1, 1)
    h = F.relu(F.linear(x, sd["BiasePredictor.fc.0.weight"]))
    o = F.linear(h, sd["BiasePredictor.fc.2.weight"])
    return (F.relu6(o + 3.0) / 3.0).view(n, t, -1)


# --------------------------------------------------------------------------------------
# MV-guided alignment
# --------------------------------------------------------------------------------------
def warp_coords(flow, h, w):
    """Un-normalised sampling coordinates exactly as the reference computes them.

    mmedit/models/common/flow_warp.py:33-43 builds ``grid + flow``, normalises with
    ``2.0 * g / max(w-1,1) - 1.0`` (fp32 multiply, TRUE division, subtract), and ATen's
    grid sampler un-normalises with ``((g + 1) / 2) * (size - 1)`` (align_corners=True).
    The round trip is NOT the identity in fp32, so it is replayed op by op.

    flow: (2,h,w) fp32 (x then y).  Returns (ix, iy) fp32 of shape (h,w).
    """
    dev = flow.device
    gy, gx = torch.meshgrid(torch.arange(0, h, device=dev, dtype=torch.float32),
                            torch.arange(0, w, device=dev, dtype=torch.float32), indexing="ij")
    # device tensors as divisors => true division on every backend (a python scalar divisor is
    # turned into a reciprocal multiply by ATen's CUDA kernel; the CPU kernel divides).
    dw = torch.tensor(float(max(w - 1, 1)), device=dev)
    dh = torch.tensor(float(max(h - 1, 1)), device=dev)
    nx = (2.0 * (gx + flow[0])) / dw - 1.0
    ny = (2.0 * (gy + flow[1])) / dh - 1.0
    ix = ((nx + 1.0) / 2.0) * float(w - 1)
    iy = ((ny + 1.0) / 2.0) * float(h - 1)
    return ix, iy


def warp_taps(flow, h, w):
    """Integer north-west tap indices (floor of the coordinates).  int32 (h,w) each."""
    ix, iy = warp_coords(flow, h, w)
    return torch.floor(ix).to(torch.int32), torch.floor(iy).to(torch.int32)


def warp_bilinear(x, flow):
    """``flow_warp(x, flow.permute(0,2,3,1))`` for one sample.

    flow_warp.py:6-50 via VOSAlignment.forward (iconvsr_mv.py:17-18); bilinear, zeros padding,
    align_corners=True.  Tap weights / accumulation order follow ATen grid_sampler_2d:
    nw=(ix_se-ix)(iy_se-iy), ne=(ix-ix_sw)(iy_sw-iy), sw=(ix_ne-ix)(iy-iy_ne), se=(ix-ix_nw)(iy-iy_nw),
    out = nw_val*nw + ne_val*ne + sw_val*sw + se_val*se with out-of-range taps contributing 0.

    x: (c,h,w) fp32, flow: (2,h,w) fp32 -> (c,h,w)
    """
    c, h, w = x.shape
    ix, iy = warp_coords(flow, h, w)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    wnw = (x1 - ix) * (y1 - iy)
    wne = (ix - x0) * (y1 - iy)
    wsw = (x1 - ix) * (iy - y0)
    wse = (ix - x0) * (iy - y0)
    flat = x.reshape(c, h * w)

    def tap(xx, yy, wt):
        ok = (xx >= 0) & (xx <= w - 1) & (yy >= 0) & (yy <= h - 1)
        idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).long().reshape(-1)
        v = flat[:, idx].reshape(c, h, w)
        return v * (wt * ok.to(wt.dtype))

    return tap(x0, y0, wnw) + tap(x1, y0, wne) + tap(x0, y1, wsw) + tap(x1, y1, wse)


# --------------------------------------------------------------------------------------
# orchestration helpers
# --------------------------------------------------------------------------------------
def spatial_padding(lrs):
    """Reflect-pad H,W up to a multiple of 4 (right/bottom).  iconvsr.py:371-394"""
    n, t, c, h, w = lrs.shape
    pad_h = (4 - h % 4) % 4
    pad_w = (4 - w % 4) % 4
    x = F.pad(lrs.reshape(-1, c, h, w), [0, pad_w, 0, pad_h], mode="reflect")
    return x.view(n, t, c, h + pad_h, w + pad_w)


def is_mirror_extended(lrs):
    """iconvsr.py:396-410: even T and first half == flipped second half."""
    if lrs.size(1) % 2 == 0:
        a, b = torch.chunk(lrs, 2, dim=1)
        return bool(torch.norm(a - b.flip(1)) == 0)
    return False


def compute_flow(mvs, mirror):
    """iconvsr_ipb.py:33-46: fwd = mvs[:,1:,:2], bwd = mvs[:,:t-1,2:]; the mirror case
    returns (None, cat(bwd, 0, fwd.flip(1)))."""
    n, t, c, h, w = mvs.shape
    fwd = mvs[:, 1:, :2]
    bwd = mvs[:, :t - 1, 2:]
    if mirror:
        zero = torch.zeros((n, 1, 2, h, w), device=mvs.device, dtype=mvs.dtype)
        return None, torch.cat([bwd, zero, fwd.flip(dims=[1])], dim=1)
    return fwd, bwd


def keyframe_mask(slices):
    """iconvsr_ipb_par.py:60-62: I (73) and P (80) frames, first and last forced.  (n,t) bool"""
    s = slices[:, :, 0, 0, 0]
    key = (s == 73) | (s == 80)
    key = key.clone()
    key[:, -1] = True
    key[:, 0] = True
    return key


def key_schedule(key_row):
    """Integer schedule for one clip.  iconvsr_ipb_par.py:81 and :116.

    Returns (bwd_key, fwd_key): bwd_key[i] = first key index > i (i < t-1),
    fwd_key[i] = last key index < i (i > 0); -1 where undefined.
    """
    t = len(key_row)
    bwd = [-1] * t
    fwd = [-1] * t
    for i in range(t - 1):
        bwd[i] = next(j for j in range(i + 1, t) if key_row[j])
    for i in range(1, t):
        fwd[i] = max(j for j in range(0, i) if key_row[j])
    return bwd, fwd


# --------------------------------------------------------------------------------------
# BAE blocks
# --------------------------------------------------------------------------------------
def dynamic_conv_se(x, weight, bias, w_experts, gamma):
    """Dynamic_conv2d_se.forward, sr_backbone_utils.py:193-209, for ONE sample.

    ``groups=batch`` there is exactly a per-sample conv, so one sample at a time is the same
    arithmetic.  x (1,64,h,w); weight (6,64,64,3,3); bias (6,64); w_experts (6,); gamma (64,)
    """
    k = weight.shape[0]
    agg_w = torch.mm(w_experts.view(1, k), weight.view(k, -1)).view(-1, weight.shape[2], 3, 3)
    agg_b = torch.mm(w_experts.view(1, k), bias).view(-1)
    out = F.conv2d(x, agg_w, agg_b, stride=1, padding=1)
    return out * gamma.view(1, -1, 1, 1)


def sparse_dyres(sd, prefix, x, par):
    """ResidualBlockNoBNDynamic_drt.sparse_conv (sr_backbone_utils.py:294-302) with the index lists of
    generate_indices(..., 1) (basicvsr_net.py:456-466, :511-514): for every partition class the 1x1 conv is
    evaluated only at the pixels whose mask is NON-ZERO (gather -> mm -> scatter), scattered results
    OVERWRITE each other in the order 16x16, 16x8, 8x8, and the sum is divided by 255 -- the mask VALUE
    is never used.  x (1,64,h,w), par (1,3,1,h,w)."""
    dyres = torch.zeros_like(x)
    for k, name in enumerate(("conv16x16", "conv16x8", "conv8x8")):
        idx = torch.nonzero(par[:, k].squeeze())                 # (m, 2) row-major pixel list
        hi, wi = idx[:, 0], idx[:, 1]
        roi = x[0, :, hi, wi]                                    # mask_roi: (64, m)
        dyres[0, :, hi, wi] = torch.mm(sd[prefix + name + ".weight"].view(64, -1), roi)   # mask_roi_back
    return dyres / 255


def bae_block(sd, prefix, x, par, w_experts, gamma, sparse=False):
    """ResidualBlockNoBNDynamic_drt.forward (channel_first, one_layer, with_se; dense par, or the
    eval-mode sparse_val=True path).

    sr_backbone_utils.py:304-333.  x (1,64,h,w), par (1,3,1,h,w).
    """
    identity = x
    if sparse:
        dyres = sparse_dyres(sd, prefix, x, par)
    else:
        dyres = (F.conv2d(x, sd[prefix + "conv16x16.weight"]) * par[:, 0]
                 + F.conv2d(x, sd[prefix + "conv16x8.weight"]) * par[:, 1]
                 + F.conv2d(x, sd[prefix + "conv8x8.weight"]) * par[:, 2])
    t = dynamic_conv_se(x, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], w_experts, gamma)
    out = F.relu(t + dyres)
    out = F.conv2d(out, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], padding=1)
    return identity + out * 1.0


def resblocks(sd, branch, x_in, par, w_experts, gamma, num_blocks=NUM_BLOCKS, sparse=False):
    """ResidualBlocksWithInputConvDynamic_drt.forward, basicvsr_net.py:506-519.

    x_in (1,131|195,h,w); par (1,3,h,w) -> viewed (1,3,1,h,w); returns (1,64,h,w).
    """
    b, c, h, w = par.shape
    par5 = par.view(b, c, 1, h, w)
    p = branch + "_resblocks."
    x = F.leaky_relu(F.conv2d(x_in, sd[p + "input_conv.0.weight"], sd[p + "input_conv.0.bias"],
                              padding=1), negative_slope=0.1)
    for k in range(num_blocks):
        x = bae_block(sd, f"{p}main.{k}.", x, par5, w_experts, gamma, sparse)
    return x


# --------------------------------------------------------------------------------------
# the generator forward
# --------------------------------------------------------------------------------------
@torch.no_grad()
def generator_forward(sd, lrs, QPs, slices, mvs, base_QPs, par_map, num_blocks=NUM_BLOCKS,
                      return_features=False, vsr=False, sparse_val=False):
    """IconVSR_restore_wo_refill_mv_ipb_fast_domain_dynamic_with_par.forward

    iconvsr_ipb_par.py:44-149 with the config kwargs of configs/HR_davis_LR_128x128.py:6-25
    (with_cat, use_base_qp, with_bias+with_se, align_key; vsr selects the x4 tail).  ``sd`` is the reference's
    ``state_dict``.  Returns (n,t,3,Hp,Wp) -- padded, NOT cropped (reference quirk).
    """
    sd = {k: v.to(lrs.device, torch.float32) for k, v in sd.items()}
    experts = base_predictor(sd, base_QPs)            # :45-46
    gammas = se_module(sd, QPs)                       # :48
    n, t, c, h_in, w_in = lrs.shape
    assert h_in >= 64 and w_in >= 64, (
        f"The height and width of inputs should be at least 64, but got {h_in} and {w_in}.")
    mirror = is_mirror_extended(lrs)                  # :53
    lrs = spatial_padding(lrs)                        # :54
    h, w = lrs.shape[3:]
    if mvs.shape[3:] != (h, w) or par_map.shape[3:] != (h, w):
        # the reference fails inside flow_warp.py:27-29 / the par multiply for non-x4 sizes
        raise ValueError(f"The spatial sizes of input ({(h, w)}) and flow/partition "
                         f"({tuple(mvs.shape[3:])}) are not the same.")
    flows_fwd, flows_bwd = compute_flow(mvs, mirror)  # :58
    key = keyframe_mask(slices).cpu().tolist()        # :60-62
    sched = [key_schedule(k) for k in key]

    outputs = [[None] * t for _ in range(n)]
    zeros = lrs.new_zeros(1, MID, h, w)
    for b in range(n):
        bwd_key, fwd_key = sched[b]
        # backward-time propagation, :67-100
        for i in range(t - 1, -1, -1):
            lr = lrs[b:b + 1, i]
            if i < t - 1:
                kidx = bwd_key[i]
                key_warp = warp_bilinear(outputs[b][kidx][0], flows_bwd[b, i]).unsqueeze(0)
                neighbor = key_warp if kidx == i + 1 else outputs[b][i + 1]   # align_key, :85-88
            else:
                key_warp, neighbor = zeros, zeros
            feat = torch.cat([lr, key_warp, neighbor], dim=1)
            outputs[b][i] = resblocks(sd, "backward", feat, par_map[b:b + 1, i], experts[b, i],
                                      gammas[b, i], num_blocks, sparse_val)
    bwd_feats = [[o.clone() for o in row] for row in outputs] if return_features else None
    outs = []
    for b in range(n):
        bwd_key, fwd_key = sched[b]
        frames = []
        # forward-time propagation, :102-147
        for i in range(t):
            lr = lrs[b:b + 1, i]
            if i > 0:
                flow = flows_fwd[b, i - 1] if flows_fwd is not None else flows_bwd[b, -i]
                kidx = fwd_key[i]
                key_warp = warp_bilinear(outputs[b][kidx][0], flow).unsqueeze(0)
                neighbor = key_warp if kidx == i - 1 else outputs[b][i - 1]
            else:
                key_warp, neighbor = zeros, zeros
            feat = torch.cat([lr, key_warp, neighbor, outputs[b][i]], dim=1)
            x = resblocks(sd, "forward", feat, par_map[b:b + 1, i], experts[b, i], gammas[b, i],
                          num_blocks, sparse_val)
            outputs[b][i] = x
            if vsr:                                   # x4 tail, :135-142 (PixelShufflePack: common/upsample.py:46-49)
                o = x
                for name in ("upsample1", "upsample2"):
                    o = F.conv2d(o, sd[name + ".upsample_conv.weight"], sd[name + ".upsample_conv.bias"], padding=1)
                    o = F.leaky_relu(F.pixel_shuffle(o, 2), 0.1)
                o = F.leaky_relu(F.conv2d(o, sd["conv_hr.weight"], sd["conv_hr.bias"], padding=1), 0.1)
                o = F.conv2d(o, sd["conv_last.weight"], sd["conv_last.bias"], padding=1)
                base = F.interpolate(lr, scale_factor=4, mode="bilinear", align_corners=False)   # :41
                frames.append(o + base)
                continue
            o = F.leaky_relu(F.conv2d(x, sd["conv_hr.weight"], sd["conv_hr.bias"], padding=1), 0.1)
            o = F.conv2d(o, sd["conv_last.weight"], sd["conv_last.bias"], padding=1)
            frames.append(o + lr)                     # :144-147
        outs.append(torch.cat(frames, dim=0))
    out = torch.stack(outs, dim=0)
    if return_features:
        return out, bwd_feats, outputs
    return out


# --------------------------------------------------------------------------------------
# metric helpers (parity measurement only)
# --------------------------------------------------------------------------------------
def tensor2img_u8(x):
    """mmedit/core/misc.py:9-74 for a (3,h,w) frame: clamp [0,1] -> x255 -> round -> uint8."""
    return (x.clamp(0, 1) * 255.0).round().to(torch.uint8)


def psnr_u8(a, b):
    """mmedit/core/evaluation/metrics.py:170-215 on uint8 frames (crop_border=0)."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if mse == 0:
        return float("inf")
    import math
    return 20.0 * math.log10(255.0 / math.sqrt(mse))
