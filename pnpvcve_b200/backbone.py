"""Drop-in generator: same registry name, constructor kwargs, forward signature and checkpoint
layout as the reference's BAE+CAA backbone, executed by the sm_100a kernels of libpnpvcve.

Boundary being mirrored (all paths relative to the reference tree):
  * class / registration: mmedit/models/backbones/sr_backbones/iconvsr_ipb_par.py:16-17
    (``@BACKBONES.register_module()``), built through ``build_backbone`` (mmedit/models/builder.py:60)
    from ``configs/HR_davis_LR_128x128*.py`` generator dicts;
  * call: ``generator(lq, QPs, slices, mvs, base_QPs, par_map)`` positionally
    (mmedit/models/restorers/basicvsr.py:179, basic_restorer.py:113,160);
  * return: new fp32 ``(n, T, 3, Hp, Wp)`` tensor, H/W rounded up to x4 and NOT cropped
    (``(n, T, 3, 4Hp, 4Wp)`` with ``vsr=True``: PixelShufflePack x2 + bilinear x4 base, :36-41,135-142);
  * ``init_weights(pretrained, strict)``: iconvsr.py:510-523;
  * ``state_dict`` keys: SURVEY.md section 8(b) (see ``pnpvcve_b200.weights``).

The parameter-holder submodules below exist only to give the parameters the reference's names;
they carry no arithmetic.  Inference only: there are no backward kernels.
"""
import torch
import torch.nn as nn

from .engine import BaeEngine
from .registry import BACKBONES, register_with_mmedit


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - holders are never called
        raise RuntimeError("parameter holder: the computation runs in libpnpvcve (see BaeEngine)")


class _ExpertConv(_Holder):
    """``conv2``: E expert kernels (E,64,64,3,3) + biases (E,64); sr_backbone_utils.py:134-164."""

    def __init__(self, mid, num_experts, init_weight):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(num_experts, mid, mid, 3, 3))
        self.bias = nn.Parameter(torch.zeros(num_experts, mid))
        if init_weight:
            for e in range(num_experts):
                nn.init.kaiming_uniform_(self.weight[e])


class _BaeBlock(_Holder):
    """Parameters of one ResidualBlockNoBNDynamic_drt (sr_backbone_utils.py:278-292)."""

    def __init__(self, mid, num_experts, init_weight):
        super().__init__()
        self.conv1 = nn.Conv2d(mid, mid, 3, 1, 1, bias=True)
        self.conv2 = _ExpertConv(mid, num_experts, init_weight)
        self.conv16x16 = nn.Conv2d(mid, mid, 1, 1, 0, bias=False)
        self.conv16x8 = nn.Conv2d(mid, mid, 1, 1, 0, bias=False)
        self.conv8x8 = nn.Conv2d(mid, mid, 1, 1, 0, bias=False)
        # default_init_weights(m, 0.1): kaiming-normal (fan_in, relu) x 0.1, zero bias -- it only
        # touches nn.Conv2d modules, i.e. not the expert conv (sr_backbone_utils.py:41-57,291-292)
        for conv in (self.conv1, self.conv16x16, self.conv16x8, self.conv8x8):
            nn.init.kaiming_normal_(conv.weight, a=0, mode="fan_in", nonlinearity="relu")
            with torch.no_grad():
                conv.weight.mul_(0.1)
            if conv.bias is not None:
                nn.init.zeros_(conv.bias)


class _PropagationBranch(_Holder):
    """``input_conv`` + ``main`` of ResidualBlocksWithInputConvDynamic_drt (basicvsr_net.py:478-519)."""

    def __init__(self, in_channels, mid, num_blocks, num_experts, init_weight):
        super().__init__()
        self.input_conv = nn.Sequential(nn.Conv2d(in_channels, mid, 3, 1, 1, bias=True),
                                        nn.LeakyReLU(negative_slope=0.1, inplace=True))
        self.main = nn.Sequential(*[_BaeBlock(mid, num_experts, init_weight) for _ in range(num_blocks)])


class _ExpertPredictor(_Holder):
    """Base_Predictor parameters (domain_aware.py:172-179): Linear(1,nf) ReLU Linear(nf,E) Softmax."""

    def __init__(self, nf, num_experts):
        super().__init__()
        self.BaseNet = nn.Sequential(nn.Linear(1, nf), nn.ReLU(True), nn.Linear(nf, num_experts),
                                     nn.Softmax(1))


class _GainPredictor(_Holder):
    """SEModule parameters (domain_aware.py:210-218): bias-free Linear(1,c/16) ReLU Linear(c/16,c)."""

    def __init__(self, channel, reduction=16):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(1, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False))


class _PixelShufflePack(_Holder):
    """PixelShufflePack parameters (common/upsample.py:27-41): conv 64 -> 64*r*r, 3x3; default_init_weights(self, 1)."""

    def __init__(self, cin, cout, scale):
        super().__init__()
        self.upsample_conv = nn.Conv2d(cin, cout * scale * scale, 3, padding=1)
        nn.init.kaiming_normal_(self.upsample_conv.weight, a=0, mode="fan_in", nonlinearity="relu")
        nn.init.zeros_(self.upsample_conv.bias)


_REQUIRED = dict(mid_channels=64, num_group=1, expert_softmax=True, use_base_qp=True, with_bias=True,
                 with_se=True, one_layer=True, blocktype="drt", channel_first=True,
                 align_key=True, with_cat=True, deform="vos", flow_inter="bilinear")


@BACKBONES.register_module()
class IconVSR_restore_wo_refill_mv_ipb_fast_domain_dynamic_with_par(nn.Module):
    """BAE+CAA generator on B200 (tcgen05 conv stack + MV-guided warp), reference-compatible.

    Constructor kwargs are those of the reference class and its parents
    (iconvsr_ipb_par.py:18, iconvsr_ipb.py:16, iconvsr.py:346-351).  The combination used by the
    three shipped configs (configs/HR_davis_LR_128x128.py:6-25) is implemented; any kwarg that
    would select a different network raises ``NotImplementedError`` instead of silently diverging.
    """

    def __init__(self, mid_channels=64, num_blocks=30, num_experts=10, num_group=1,
                 expert_softmax=False, use_base_qp=False, with_bias=False, with_se=False,
                 with_par=False, init_weight=False, one_layer=False, small_sft=False,
                 blocktype="default", channel_first=False, drconv=False, sparse_val=False, vsr=False,
                 align_key=False, with_cat=False, deform="vos", max_residue_magnitude=10,
                 flow_inter="bilinear", keyframe_stride=5, padding=2):
        super().__init__()
        given = dict(mid_channels=mid_channels, num_group=num_group, expert_softmax=expert_softmax,
                     use_base_qp=use_base_qp, with_bias=with_bias, with_se=with_se,
                     one_layer=one_layer, blocktype=blocktype, channel_first=channel_first,
                     align_key=align_key, with_cat=with_cat,
                     deform=deform, flow_inter=flow_inter)
        bad = {k: v for k, v in given.items() if v != _REQUIRED[k]}
        if bad:
            raise NotImplementedError(
                "pnpvcve_b200 implements the generator configuration of configs/HR_davis_LR_128x128*.py; "
                f"unsupported kwargs: {bad} (required: { {k: _REQUIRED[k] for k in bad} })")
        if not (1 <= int(num_experts) <= 16) or int(num_blocks) < 1:
            raise NotImplementedError("num_experts must be in [1,16] and num_blocks >= 1")
        self.mid_channels = mid_channels
        self.num_blocks = int(num_blocks)
        self.num_experts = int(num_experts)
        self.padding = padding
        self.keyframe_stride = keyframe_stride
        self.flow_inter = flow_inter
        self.with_cat, self.use_base_qp, self.with_bias = with_cat, use_base_qp, with_bias
        self.with_par, self.vsr, self.align_key = with_par, vsr, align_key
        #: the reference's eval-mode "sparse conv" (sr_backbone_utils.py:294-302,307-308): per pixel the 1x1 conv of the
        #: last partition class with a non-zero mask, / 255.  Defined per image (the reference's index lists only
        #: work for n == 1: basicvsr_net.py:458 squeezes the batch away).
        self.sparse_val = bool(sparse_val)
        self.is_mirror_extended = False

        self.BiasePredictor = _GainPredictor(mid_channels)
        self.BasePredictor = _ExpertPredictor(mid_channels, self.num_experts)
        self.backward_resblocks = _PropagationBranch(2 * mid_channels + 3, mid_channels, self.num_blocks,
                                                     self.num_experts, init_weight)
        self.forward_resblocks = _PropagationBranch(3 * mid_channels + 3, mid_channels, self.num_blocks,
                                                    self.num_experts, init_weight)
        self.conv_hr = nn.Conv2d(64, 64, 3, 1, 1)
        self.conv_last = nn.Conv2d(64, 3, 3, 1, 1)
        if vsr:   # x4 tail: PixelShufflePack x2, bilinear base (iconvsr_ipb_par.py:36-41); img_upsample has no parameters
            self.upsample1 = _PixelShufflePack(mid_channels, mid_channels, 2)
            self.upsample2 = _PixelShufflePack(mid_channels, 64, 2)
        self._engine = BaeEngine(self)

    # packed weights follow the parameters: every path that rewrites them drops the engine's caches
    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._engine.invalidate()
        return out

    def load_state_dict(self, *args, **kwargs):
        res = super().load_state_dict(*args, **kwargs)
        self._engine.invalidate()
        return res

    def invalidate_packed_weights(self):
        """Call after writing parameters through ``.data`` (EMA swaps, ``p.data.copy_``): such writes do not bump
        the parameter version the packed-weight cache is keyed on."""
        self._engine.invalidate()

    # ------------------------------------------------------------------ reference API
    def init_weights(self, pretrained=None, strict=True):
        """iconvsr.py:510-523: str -> load checkpoint, None -> keep init, anything else -> TypeError."""
        if isinstance(pretrained, str):
            ckpt = torch.load(pretrained, map_location="cpu")
            sd = ckpt.get("state_dict", ckpt)
            if any(k.startswith("generator.") for k in sd):     # BasicVSR checkpoints prefix the generator
                sd = {k[len("generator."):]: v for k, v in sd.items() if k.startswith("generator.")}
            self.load_state_dict(sd, strict=strict)
        elif pretrained is not None:
            raise TypeError(f'"pretrained" must be a str or None. But received {type(pretrained)}.')

    def forward(self, lrs, QPs=None, slices=None, mvs=None, base_QPs=None, par_map=None):
        """lrs (n,T,3,H,W); QPs, slices, base_QPs (n,T,1,1,1); mvs (n,T,4,H,W); par_map (n,T,3,H,W)."""
        tensors = dict(lrs=lrs, QPs=QPs, slices=slices, mvs=mvs, base_QPs=base_QPs, par_map=par_map)
        for name, tns in tensors.items():
            if tns is None:
                raise TypeError(f"{name} is required by the BAE+CAA generator")
        if torch.is_grad_enabled() and self.training:
            raise RuntimeError("pnpvcve_b200 is inference-only (no backward kernels): call under "
                               "torch.no_grad() / model.eval()")
        if not lrs.is_cuda:
            raise RuntimeError("pnpvcve_b200 runs on sm_100 CUDA devices only; there is no CPU fallback")
        return self._engine.forward(lrs, QPs, slices, mvs, base_QPs, par_map)

    def forward_streamed(self, lrs, QPs, slices, mvs, base_QPs, par_map, cond_host=None, frame_ready=None,
                         frame_done=None, out=None):
        """Same computation as ``forward`` for a clip that is still arriving / already leaving: ``frame_ready(i)`` is
        called before frame i's inputs are first read (backward-time order T-1 .. 0), ``frame_done(i, out)`` after
        frame i of the output has been enqueued, ``cond_host`` = host copies of (slices, base_QPs, QPs) saves the one
        device->host copy.  Used by ``pnpvcve_b200.driver.stream_clips`` to overlap H2D / D2H with the kernels."""
        if not lrs.is_cuda:
            raise RuntimeError("pnpvcve_b200 runs on sm_100 CUDA devices only; there is no CPU fallback")
        return self._engine.forward(lrs, QPs, slices, mvs, base_QPs, par_map, cond_host=cond_host,
                                    frame_ready=frame_ready, frame_done=frame_done, out=out)

    def forward_with_features(self, lrs, QPs, slices, mvs, base_QPs, par_map):
        """Test hook: also returns the backward / forward propagation features (bf16 NHWC)."""
        return self._engine.forward(lrs, QPs, slices, mvs, base_QPs, par_map, return_features=True)

    @property
    def gpu_launches(self):
        """Kernels launched by the last forward call."""
        return self._engine.launch_count


BAEGenerator = IconVSR_restore_wo_refill_mv_ipb_fast_domain_dynamic_with_par
register_with_mmedit(BAEGenerator)


def num_parameters(module):
    return sum(p.numel() for p in module.parameters())
