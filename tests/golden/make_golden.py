"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case the reference generator class (imported from the reference's own files through
``oracle/refshim.py``) is constructed with the config kwargs, the seeded recipe weights
(``pnpvcve_b200.weights.random_state_dict``) are loaded with ``strict=True`` -- which also pins
the checkpoint key layout -- and the forward is run on a seeded synthetic clip
(``pnpvcve_b200.synthetic``).  Stored per case: the output on a stride-2 pixel lattice (fp32),
float64 per-frame sums of the full output, and the case description needed to rebuild the inputs.
The warp case stores the reference ``flow_warp`` on a 720x1280 plane (stride-8 lattice).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from oracle import refshim  # noqa: E402
from pnpvcve_b200 import synthetic, weights  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

#: name -> dict(clips=[make_clip kwargs...], weight_seed, num_blocks, mirror)
CASES = {
    "c1_128x128_t7": dict(weight_seed=0, num_blocks=8, clips=[
        dict(h=128, w=128, t=7, seed=1000, crf=25, mv_qpel=32, ipb=False, pattern="IBBP")]),
    "n2_64x96_t6_ipb": dict(weight_seed=1, num_blocks=8, clips=[
        dict(h=64, w=96, t=6, seed=3000, crf=15, mv_qpel=32, ipb=True, pattern="IBBP"),
        dict(h=64, w=96, t=6, seed=3001, crf=35, mv_qpel=32, ipb=True, pattern="IP")]),
    "allB_64x64_t5": dict(weight_seed=2, num_blocks=8, clips=[
        dict(h=64, w=64, t=5, seed=3100, crf=25, mv_qpel=64, ipb=True, pattern="allB")]),
    "mirror_64x64_t6": dict(weight_seed=3, num_blocks=8, mirror=True, clips=[
        dict(h=64, w=64, t=6, seed=3200, crf=25, mv_qpel=32, ipb=False, pattern="IBBP")]),
    "t1_64x64": dict(weight_seed=4, num_blocks=8, clips=[
        dict(h=64, w=64, t=1, seed=3300, crf=35, mv_qpel=32, ipb=False, pattern="IBBP")]),
    "edge_68x132_t3": dict(weight_seed=5, num_blocks=8, clips=[
        dict(h=68, w=132, t=3, seed=3400, crf=15, mv_qpel=64, ipb=True, pattern="IBBP")]),
    # vsr=True: PixelShufflePack x2 + bilinear x4 base tail (iconvsr_ipb_par.py:36-41,135-142); output is 4H x 4W
    # sparse_val=True eval path (sr_backbone_utils.py:294-302) on overlapping, non-1/255 partition maps; and the
    # dense path on the same maps (the values matter there)
    "sparse_64x64_t3": dict(weight_seed=7, num_blocks=8, sparse_val=True, par_overlap=77, clips=[
        dict(h=64, w=64, t=3, seed=3600, crf=35, mv_qpel=32, ipb=True, pattern="IBBP")]),
    "densepar_64x64_t3": dict(weight_seed=7, num_blocks=8, par_overlap=77, clips=[
        dict(h=64, w=64, t=3, seed=3600, crf=35, mv_qpel=32, ipb=True, pattern="IBBP")]),
    "vsr_64x96_t3": dict(weight_seed=6, num_blocks=8, vsr=True, clips=[
        dict(h=64, w=96, t=3, seed=3500, crf=25, mv_qpel=32, ipb=True, pattern="IBBP")]),
}


def build_inputs(case):
    clips = [synthetic.make_clip(**kw) for kw in case["clips"]]
    clip = synthetic.cat_clips(clips)
    if case.get("par_overlap"):
        synthetic.overlap_partitions(clip, case["par_overlap"])
    if case.get("mirror"):
        # even-T exact mirror: frame i == frame t-1-i (iconvsr.py:396-410)
        t = clip["lq"].shape[1]
        half = clip["lq"][:, : t // 2]
        clip["lq"] = torch.cat([half, half.flip(1)], dim=1).contiguous()
    return clip


def main():
    torch.set_num_threads(os.cpu_count())
    only = set(sys.argv[1:])           # optional: regenerate just these cases (cases.json is merged)
    meta = {}
    cases_path = os.path.join(HERE, "cases.json")
    if only and os.path.isfile(cases_path):
        with open(cases_path) as f:
            meta = json.load(f)
    for name, case in CASES.items():
        if only and name not in only:
            continue
        vsr = bool(case.get("vsr", False))
        net = refshim.build_reference(seed=0, num_blocks=case["num_blocks"], vsr=vsr,
                                      sparse_val=bool(case.get("sparse_val", False)))
        sd = weights.random_state_dict(case["weight_seed"], num_blocks=case["num_blocks"], vsr=vsr)
        net.load_state_dict(sd, strict=True)
        clip = build_inputs(case)
        with torch.no_grad():
            out = net(*synthetic.generator_args(clip))
        out = out.contiguous()
        # every pixel OFF the stride-2 lattice as an fp16 residual against the frame the network adds its output to
        # (the padded LR frame, or its x4 bilinear upsampling with vsr): |residual| ~ 0.07, i.e. ~3e-5 absolute
        res = (out - helpers.residual_base(clip["lq"], vsr)).numpy()
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            off_lattice_f16=np.stack([res[..., 0::2, 1::2], res[..., 1::2, 0::2], res[..., 1::2, 1::2]]).astype(np.float16),
            lattice=out[..., ::2, ::2].numpy().astype(np.float32),
            frame_sum=out.double().sum(dim=(2, 3, 4)).numpy(),
            frame_abs_sum=out.double().abs().sum(dim=(2, 3, 4)).numpy(),
            shape=np.array(out.shape))
        meta[name] = case
        print(name, tuple(out.shape), float(out.min()), float(out.max()))

    if only and "warp_720p" not in only:
        with open(cases_path, "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        return
    # reference flow_warp on a 720p plane
    fw = sys.modules["mmedit.models.common.flow_warp"].flow_warp
    g = torch.Generator().manual_seed(4242)
    x = torch.randn((1, 2, 720, 1280), generator=g)
    flow_b = torch.randint(-64, 65, (1, 2, 90, 160), generator=g).float() / 4.0
    flow = flow_b.repeat_interleave(8, 2).repeat_interleave(8, 3)
    y = fw(x, flow.permute(0, 2, 3, 1))
    np.savez_compressed(os.path.join(HERE, "warp_720p.npz"),
                        lattice=y[..., ::8, ::8].numpy(), out_sum=y.double().sum().item(),
                        out_abs_sum=y.double().abs().sum().item())
    meta["warp_720p"] = dict(seed=4242, c=2, h=720, w=1280, block=8, mv_qpel=64)
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
