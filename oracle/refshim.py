"""Loader for the UNMODIFIED reference generator (test infrastructure only).

TEST INFRASTRUCTURE -- never imported by the product package ``pnpvcve_b200``.

The reference (``/root/reference``, ZeldaM1/PnP-VCVE) is a fork of MMEditing that
needs ``mmcv-full`` 1.3.13..1.6 (``mmedit/__init__.py:22-32``), which is not
installable here.  The twelve files on the BAE+CAA hot path contain no mmcv
arithmetic, so they run unmodified once a stub ``mmcv`` exposing the handful of
names they import is placed in ``sys.modules`` and the heavy ``mmedit`` package
``__init__`` files are replaced by bare namespace modules.

This file only exists to (a) pin ``oracle/bae_oracle.py`` against the real
reference in the build container and (b) generate the committed golden vectors
(``tests/golden/make_golden.py``).  ``/root/reference`` does not exist on the GPU
box, so nothing that runs there may import this module; ``available()`` tells.
"""
import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("PNP_REFERENCE_ROOT", "/root/reference")

REF_CLASS = "IconVSR_restore_wo_refill_mv_ipb_fast_domain_dynamic_with_par"


def available():
    return os.path.isfile(os.path.join(
        REFERENCE_ROOT, "mmedit/models/backbones/sr_backbones/iconvsr_ipb_par.py"))


class _Registry:
    """Just enough of mmcv.utils.Registry for ``@BACKBONES.register_module()``."""

    def __init__(self, name, parent=None, **kw):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)


def _kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0,
                  distribution="normal"):
    # mmcv 1.x mmcv/cnn/utils/weight_init.py: kaiming_init
    if hasattr(module, "weight") and module.weight is not None:
        if distribution == "uniform":
            nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        else:
            nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _constant_init(module, val, bias=0):
    if hasattr(module, "weight") and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, "bias") and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _install_mmcv_stub():
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "__pnp_stub__", False):
        return

    def mod(name):
        m = types.ModuleType(name)
        m.__pnp_stub__ = True
        sys.modules[name] = m
        return m

    mmcv = mod("mmcv")
    mmcv.__version__ = "1.5.0"
    cnn = mod("mmcv.cnn")
    runner = mod("mmcv.runner")
    ops = mod("mmcv.ops")
    utils = mod("mmcv.utils")
    pw = mod("mmcv.utils.parrots_wrapper")
    mmcv.cnn, mmcv.runner, mmcv.ops, mmcv.utils = cnn, runner, ops, utils
    utils.parrots_wrapper = pw

    class ConvModule(nn.Module):  # imported by name only on the hot path
        def __init__(self, *a, **k):
            raise NotImplementedError("mmcv stub: ConvModule is not on the BAE+CAA path")

    class ModulatedDeformConv2d(nn.Module):  # base class of the unused DCN alignments
        def __init__(self, *a, **k):
            raise NotImplementedError("mmcv stub: DCN is out of scope (deform='vos')")

    def modulated_deform_conv2d(*a, **k):
        raise NotImplementedError("mmcv stub: DCN is out of scope (deform='vos')")

    def load_checkpoint(model, filename, map_location=None, strict=False, logger=None):
        ckpt = torch.load(filename, map_location=map_location or "cpu")
        sd = ckpt.get("state_dict", ckpt)
        model.load_state_dict(sd, strict=strict)
        return ckpt

    def build_from_cfg(cfg, registry, default_args=None):
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        cls = registry.get(args.pop("type"))
        return cls(**args)

    cnn.ConvModule = ConvModule
    cnn.constant_init = _constant_init
    cnn.kaiming_init = _kaiming_init
    cnn.MODELS = _Registry("model")
    runner.load_checkpoint = load_checkpoint
    ops.ModulatedDeformConv2d = ModulatedDeformConv2d
    ops.modulated_deform_conv2d = modulated_deform_conv2d
    utils.Registry = _Registry
    utils.get_logger = lambda *a, **k: None
    pw._BatchNorm = nn.modules.batchnorm._BatchNorm
    mmcv.build_from_cfg = build_from_cfg


def _load_file(modname, relpath):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[modname] = m
    spec.loader.exec_module(m)
    return m


def _namespace(name, relpath):
    m = types.ModuleType(name)
    m.__path__ = [os.path.join(REFERENCE_ROOT, relpath)]
    sys.modules[name] = m
    return m


_loaded = None


def load_reference():
    """Returns the reference generator CLASS, imported from the reference's own files."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    _install_mmcv_stub()
    mmedit = _namespace("mmedit", "mmedit")
    models = _namespace("mmedit.models", "mmedit/models")
    _namespace("mmedit.models.backbones", "mmedit/models/backbones")
    _namespace("mmedit.models.backbones.sr_backbones", "mmedit/models/backbones/sr_backbones")
    common = _namespace("mmedit.models.common", "mmedit/models/common")
    mutils = types.ModuleType("mmedit.utils")
    mutils.get_root_logger = lambda *a, **k: None
    sys.modules["mmedit.utils"] = mutils
    mmedit.models, mmedit.utils = models, mutils

    reg = _load_file("mmedit.models.registry", "mmedit/models/registry.py")
    models.registry = reg
    _load_file("mmedit.models.common.partition_aware", "mmedit/models/common/partition_aware.py")
    sbu = _load_file("mmedit.models.common.sr_backbone_utils", "mmedit/models/common/sr_backbone_utils.py")
    fw = _load_file("mmedit.models.common.flow_warp", "mmedit/models/common/flow_warp.py")
    up = _load_file("mmedit.models.common.upsample", "mmedit/models/common/upsample.py")
    for name in dir(sbu):
        if name.startswith("ResidualBlock") or name in ("make_layer", "default_init_weights"):
            setattr(common, name, getattr(sbu, name))
    common.flow_warp = fw.flow_warp
    common.PixelShufflePack = up.PixelShufflePack

    pkg = "mmedit.models.backbones.sr_backbones"
    m = importlib.import_module(pkg + ".iconvsr_ipb_par")
    _loaded = getattr(m, REF_CLASS)
    return _loaded


#: generator kwargs of configs/HR_davis_LR_128x128.py:6-25 (minus ``type``)
GENERATOR_KWARGS = dict(
    mid_channels=64, num_blocks=8, padding=3, with_cat=True, use_base_qp=True,
    num_experts=6, expert_softmax=True, init_weight=True, with_bias=True, with_se=True,
    with_par=True, one_layer=True, blocktype="drt", channel_first=True, sparse_val=False,
    align_key=True, vsr=False)


def build_reference(seed=0, **overrides):
    """Reference generator with the config's kwargs, random-init under ``seed``."""
    cls = load_reference()
    kw = dict(GENERATOR_KWARGS)
    kw.update(overrides)
    torch.manual_seed(seed)
    net = cls(**kw)
    net.eval()
    return net


_restorer = None


def load_restorer():
    """The reference's own MODEL wrapper ``BasicVSR`` (mmedit/models/restorers/basicvsr.py, basic_restorer.py, base.py,
    builder.py -- loaded unmodified), its ``MODELS`` registry and its ``build_model``.  ``mmedit.core`` is replaced by a
    namespace holding the reference's own ``psnr`` / ``ssim`` / ``tensor2img`` functions (their sources are executed from
    the reference files; the modules themselves import cv2-free parts of mmcv that the stub does not provide).  Used by
    tests/test_restorer_dropin.py to drive the B200 generator through the reference's L2 -> L1 call."""
    global _restorer
    if _restorer is not None:
        return _restorer
    import ast
    import math

    import numpy as np
    load_reference()
    mmcv = sys.modules["mmcv"]
    runner = sys.modules["mmcv.runner"]

    def auto_fp16(apply_to=None, out_fp32=False):     # mmcv.runner.auto_fp16 is the identity unless fp16_enabled
        def deco(fn):
            return fn
        return deco

    runner.auto_fp16 = auto_fp16
    mmcv.imwrite = lambda *a, **k: None

    def load_functions(relpath, names, ns):
        path = os.path.join(REFERENCE_ROOT, relpath)
        tree = ast.parse(open(path).read())
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                exec(compile(ast.Module([node], []), path + ":" + node.name, "exec"), ns)
        return ns

    try:
        import cv2
    except ImportError:  # psnr does not need it
        cv2 = None
    ns = dict(np=np, cv2=cv2, torch=torch, math=math, make_grid=None, mmcv=None)
    load_functions("mmedit/core/evaluation/metrics.py", {"reorder_image", "psnr", "_ssim", "ssim"}, ns)
    load_functions("mmedit/core/misc.py", {"tensor2img"}, ns)
    core = types.ModuleType("mmedit.core")
    core.psnr, core.ssim, core.tensor2img = ns["psnr"], ns["ssim"], ns["tensor2img"]
    sys.modules["mmedit.core"] = core
    sys.modules["mmedit"].core = core

    _load_file("mmedit.models.base", "mmedit/models/base.py")
    builder = _load_file("mmedit.models.builder", "mmedit/models/builder.py")
    _namespace("mmedit.models.losses", "mmedit/models/losses")
    _load_file("mmedit.models.losses.utils", "mmedit/models/losses/utils.py")
    _load_file("mmedit.models.losses.pixelwise_loss", "mmedit/models/losses/pixelwise_loss.py")
    _namespace("mmedit.models.restorers", "mmedit/models/restorers")
    _load_file("mmedit.models.restorers.basic_restorer", "mmedit/models/restorers/basic_restorer.py")
    vsr = _load_file("mmedit.models.restorers.basicvsr", "mmedit/models/restorers/basicvsr.py")
    reg = sys.modules["mmedit.models.registry"]
    _restorer = dict(BasicVSR=vsr.BasicVSR, MODELS=reg.MODELS, BACKBONES=reg.BACKBONES,
                     build_model=builder.build_model, build_backbone=builder.build_backbone)
    return _restorer
