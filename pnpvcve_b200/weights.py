"""Checkpoint layout of the generator and a seeded random-init recipe.

The key names / shapes are the reference's ``state_dict`` (verified in
``tests/golden/make_golden.py`` by a strict ``load_state_dict`` into the reference class):

    {backward,forward}_resblocks.input_conv.0.{weight (64,131|195,3,3), bias}
    {...}_resblocks.main.{k}.conv1.{weight (64,64,3,3), bias}            sr_backbone_utils.py:283
    {...}_resblocks.main.{k}.conv2.{weight (E,64,64,3,3), bias (E,64)}   sr_backbone_utils.py:284
    {...}_resblocks.main.{k}.conv{16x16,16x8,8x8}.weight (64,64,1,1)     sr_backbone_utils.py:285-287
    conv_hr.{weight (64,64,3,3), bias}, conv_last.{weight (3,64,3,3), bias}   iconvsr.py:365-366
    upsample{1,2}.upsample_conv.{weight (256,64,3,3), bias}   (vsr=True only)  iconvsr_ipb_par.py:36-40
    BiasePredictor.fc.0.weight (4,1), BiasePredictor.fc.2.weight (64,4)   domain_aware.py:213-218
    BasePredictor.BaseNet.0.{weight (64,1), bias}, .2.{weight (E,64), bias}   domain_aware.py:175-179

``random_state_dict`` draws every tensor from its own seeded generator with the scale of the
reference's initialisers (kaiming-uniform experts, kaiming-normal x0.1 for conv1 and the 1x1
convs, torch defaults elsewhere).  Unlike the reference's init it gives the biases small
non-zero values so that the bias paths are exercised.  The recipe does not depend on the
reference, so the GPU box can rebuild exactly the weights the golden vectors were made with.
"""
import math
import zlib

import torch

MID = 64


def state_dict_shapes(num_blocks=8, num_experts=6, mid=MID, vsr=False):
    shapes = {}
    for branch, cin in (("backward", 2 * mid + 3), ("forward", 3 * mid + 3)):
        p = f"{branch}_resblocks."
        shapes[p + "input_conv.0.weight"] = (mid, cin, 3, 3)
        shapes[p + "input_conv.0.bias"] = (mid,)
        for k in range(num_blocks):
            q = f"{p}main.{k}."
            shapes[q + "conv1.weight"] = (mid, mid, 3, 3)
            shapes[q + "conv1.bias"] = (mid,)
            shapes[q + "conv2.weight"] = (num_experts, mid, mid, 3, 3)
            shapes[q + "conv2.bias"] = (num_experts, mid)
            shapes[q + "conv16x16.weight"] = (mid, mid, 1, 1)
            shapes[q + "conv16x8.weight"] = (mid, mid, 1, 1)
            shapes[q + "conv8x8.weight"] = (mid, mid, 1, 1)
    shapes["conv_hr.weight"] = (mid, mid, 3, 3)
    shapes["conv_hr.bias"] = (mid,)
    shapes["conv_last.weight"] = (3, mid, 3, 3)
    shapes["conv_last.bias"] = (3,)
    if vsr:   # PixelShufflePack x2 (iconvsr_ipb_par.py:36-40, common/upsample.py:27-31)
        for name in ("upsample1", "upsample2"):
            shapes[name + ".upsample_conv.weight"] = (4 * mid, mid, 3, 3)
            shapes[name + ".upsample_conv.bias"] = (4 * mid,)
    shapes["BiasePredictor.fc.0.weight"] = (mid // 16, 1)
    shapes["BiasePredictor.fc.2.weight"] = (mid, mid // 16)
    shapes["BasePredictor.BaseNet.0.weight"] = (mid, 1)
    shapes["BasePredictor.BaseNet.0.bias"] = (mid,)
    shapes["BasePredictor.BaseNet.2.weight"] = (num_experts, mid)
    shapes["BasePredictor.BaseNet.2.bias"] = (num_experts,)
    return shapes


def _fan_in(shape):
    if len(shape) == 5:      # (E, out, in, kh, kw): per-expert fan-in
        return shape[2] * shape[3] * shape[4]
    if len(shape) == 4:
        return shape[1] * shape[2] * shape[3]
    return shape[1]


def random_state_dict(seed=0, num_blocks=8, num_experts=6, vsr=False):
    sd = {}
    shapes = state_dict_shapes(num_blocks, num_experts, vsr=vsr)
    for key, shape in shapes.items():
        g = torch.Generator()
        g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        leaf = key.rsplit(".", 2)
        if key.endswith("bias"):
            wkey = key[:-4] + "weight"
            bound = 1.0 / math.sqrt(_fan_in(shapes[wkey]))
            if ".conv1." in key or ".conv2." in key:
                bound *= 0.1
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif ".conv2.weight" in key:                       # kaiming_uniform_ per expert
            bound = math.sqrt(6.0 / _fan_in(shape))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif "upsample_conv.weight" in key:                # default_init_weights(self, 1): kaiming-normal
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / _fan_in(shape))
        elif ".conv1.weight" in key or "conv16x" in key or "conv8x8" in key:
            t = torch.randn(shape, generator=g) * (0.1 * math.sqrt(2.0 / _fan_in(shape)))
        else:                                              # torch default: U(+-1/sqrt(fan_in))
            bound = 1.0 / math.sqrt(_fan_in(shape))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        del leaf
        sd[key] = t.to(torch.float32)
    return sd
