"""MV / partition rasteriser: oracle vs. the reference's golden vectors (CPU), CUDA kernel vs. both (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import mv_raster_oracle as R
from pnpvcve_b200 import sideinfo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["raster_64x96_ibbp", "raster_72x80_messy", "raster_128x128_ip"]


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    offs = g["offsets"]
    recs = [g["records"][offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    h, w = (int(v) for v in g["shape"])
    return g, recs, [str(c) for c in g["pattern"]], h, w


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_loop_golden(name):
    g, recs, pattern, h, w = load_case(name)
    mv, par = R.rasterize_clip(recs, pattern, h, w)
    assert np.array_equal(mv, g["mvs"]) and np.array_equal(par, g["partitions"])
    assert set(np.unique(par)) <= {np.float32(0), np.float32(1) / np.float32(255)}


def test_p_targets_follow_p_offset_rule():
    assert sideinfo.p_targets("IBBPBBP") == [-1, -1, -1, 0, -1, -1, 3]
    assert sideinfo.p_targets("IPPP") == [-1, 0, 1, 2]
    assert sideinfo.p_targets("PBP") == [-1, -1, 0]


def test_oracle_errors_like_reference():
    bad_area = [np.zeros((0, 10), np.float32), np.array([[-1, 4, 4, 8, 8, 8, 8, 1, 1, 4]], np.float32)]
    with pytest.raises(KeyError):
        R.rasterize_clip(bad_area, ["I", "P"], 32, 32)
    first_p = [np.array([[1, 8, 8, 8, 8, 8, 8, 1, 1, 4]], np.float32)]
    with pytest.raises(UnboundLocalError):
        R.rasterize_clip(first_p, ["P"], 32, 32)


def test_synthetic_records_tile_the_frame():
    recs = sideinfo.synthetic_records(64, 96, "IBP", seed=3)
    assert recs[0].shape == (0, 10)
    mv, par = R.rasterize_clip(recs, list("IBP"), 64, 96)
    assert np.allclose(par[1].sum(0), 1.0 / 255.0)          # one-hot everywhere on coded frames
    assert par[0].sum() == 0


def test_pack_side_layout():
    """Host form of the compact feed: records (R,10) fp32 + ONE int32 vector [offsets (T+1) | is_b (T) | p_target (T)]."""
    recs = sideinfo.synthetic_records(32, 32, "IBPB", seed=1)
    flat = np.concatenate(recs, 0)
    offs = np.cumsum([0] + [len(r) for r in recs])
    side = sideinfo.pack_side(flat, offs, list("IBPB"))
    assert side["t"] == 4 and tuple(side["records"].shape) == (len(flat), 10) and side["records"].dtype == torch.float32
    meta = side["meta"].tolist()
    assert meta[:5] == offs.tolist() and meta[5:9] == [0, 1, 0, 1] and meta[9:] == sideinfo.p_targets("IBPB")
    with pytest.raises(ValueError):
        sideinfo.pack_side(flat, offs[:-1], list("IBPB"))


def test_side_from_files_follows_the_loader_paths(tmp_path):
    """Record tables are found where LoadImageFromFileList_ipb looks for them (loading_ipb.py:316-325) and packed in
    frame order; the oracle of the loader loop on the same tables gives the dense planes the GPU path must produce."""
    pattern = "IBP"
    recs = sideinfo.synthetic_records(32, 48, pattern, seed=4)
    root = tmp_path / "crf25"
    (root / "png" / "000").mkdir(parents=True)
    (root / "mv" / "000").mkdir(parents=True)
    paths = []
    for f, r in enumerate(recs):
        paths.append(str(root / "png" / "000" / f"{f:08d}.png"))
        np.save(str(root / "mv" / "000" / f"{f:08d}.npy"), r.astype(np.float64))       # the codec tool writes float64
    assert sideinfo.mv_record_path(paths[1]) == str(root / "mv" / "000" / "00000001.npy")
    assert sideinfo.mv_record_path("/d/png/00001/0001/im3.png", "vimeo") == "/d/mv/00001/0001/00000002.npy"
    assert sideinfo.mv_record_path("/d/x_crf25/png/000012_11.png", "kitti") == "/d/x_crf25/mv/000012/00000001.npy"
    side = sideinfo.side_from_files(paths, pattern)
    assert side["t"] == 3 and side["records"].shape[0] == sum(len(r) for r in recs)
    assert torch.equal(side["records"], torch.from_numpy(np.concatenate(recs, 0)))
    assert side["meta"][:4].tolist() == np.cumsum([0] + [len(r) for r in recs]).tolist()
    with pytest.raises(ValueError):
        sideinfo.side_from_files(paths, "IB")


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_reference_golden_bit_exact(dev, name):
    g, recs, pattern, h, w = load_case(name)
    mv, par = sideinfo.rasterize_clip(g["records"], g["offsets"], pattern, h, w, device=dev)
    assert torch.equal(mv.cpu(), torch.from_numpy(g["mvs"]))
    assert torch.equal(par.cpu(), torch.from_numpy(g["partitions"]))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,pattern,messy", [(720, 1280, "IBBP", True), (376, 1244, "IP", True),
                                                (180, 320, "IBBPBBPB", False)])
def test_kernel_matches_oracle_at_full_sizes(dev, h, w, pattern, messy):
    recs = sideinfo.synthetic_records(h, w, pattern, seed=h + w, messy=messy)
    flat = np.concatenate(recs, 0)
    offs = np.cumsum([0] + [len(r) for r in recs])
    mv, par = sideinfo.rasterize_clip(flat, offs, list(pattern), h, w, device=dev)
    emv, epar = R.rasterize_clip(recs, list(pattern), h, w)
    assert torch.equal(mv.cpu(), torch.from_numpy(emv))
    assert torch.equal(par.cpu(), torch.from_numpy(epar))


@pytest.mark.gpu
def test_kernel_reports_reference_errors(dev):
    bad = np.array([[-1, 4, 4, 8, 8, 8, 8, 1, 1, 4]], np.float32)
    with pytest.raises(KeyError):
        sideinfo.rasterize_clip(bad, [0, 0, 1], ["I", "P"], 32, 32, device=dev)
    first_p = np.array([[1, 8, 8, 8, 8, 8, 8, 1, 1, 4]], np.float32)
    with pytest.raises(ValueError):
        sideinfo.rasterize_clip(first_p, [0, 1], ["P"], 32, 32, device=dev)
    mv, par = sideinfo.rasterize_clip(np.zeros((0, 10), np.float32), [0, 0], ["I"], 64, 64, device=dev)
    assert mv.abs().max().item() == 0 and par.abs().max().item() == 0


@pytest.mark.gpu
def test_rasterised_side_info_drives_the_generator(dev):
    """records -> pnp_mv_rasterize -> generator == dense reference-style inputs -> generator."""
    import pnpvcve_b200 as P
    from pnpvcve_b200 import synthetic, weights
    from test_gpu_parity import GENERATOR_CFG
    h, w, pattern = 64, 96, "IBBP"
    recs = sideinfo.synthetic_records(h, w, pattern, seed=5)
    emv, epar = R.rasterize_clip(recs, list(pattern), h, w)
    flat = np.concatenate(recs, 0)
    offs = np.cumsum([0] + [len(r) for r in recs])
    mv, par = sideinfo.rasterize_clip(flat, offs, list(pattern), h, w, device=dev)
    clip = synthetic.make_clip(h, w, len(pattern), seed=9, pattern="IBBP")
    net = P.build_backbone(dict(GENERATOR_CFG))
    net.load_state_dict(weights.random_state_dict(2), strict=True)
    net = net.to(dev).eval()
    args = [a.to(dev) for a in synthetic.generator_args(clip)]
    with torch.no_grad():
        a = net(args[0], args[1], args[2], mv[None], args[4], par[None])
        b = net(args[0], args[1], args[2], torch.from_numpy(emv)[None].to(dev), args[4],
                torch.from_numpy(epar)[None].to(dev))
    assert torch.equal(a, b)


def _side_entries(h, w, t, n, seed0, oracle_dense):
    """(dense host entries, compact host entries) of the same clips: the dense planes come from the numpy oracle
    of the reference's loader loop (or are left to the caller), the compact form carries the records themselves."""
    from pnpvcve_b200 import synthetic
    dense, compact = [], []
    for e in range(2):
        clips, sides = [], []
        for c in range(n):
            clip = synthetic.make_clip(h, w, t, seed=seed0 + 10 * e + c, crf=(15, 35)[c % 2], ipb=True)
            types = [chr(int(v)) for v in clip["slices"].flatten()]
            recs = sideinfo.synthetic_records(h, w, types, seed=seed0 + 100 * e + c, messy=(c == 0))
            flat = np.concatenate(recs, 0)
            offs = np.cumsum([0] + [len(r) for r in recs])
            emv, epar = oracle_dense(recs, types, h, w)
            clip["mvs"], clip["partitions"] = torch.from_numpy(emv)[None], torch.from_numpy(epar)[None]
            clips.append(clip)
            sides.append(sideinfo.pack_side(flat, offs, types))
        entry = synthetic.cat_clips(clips)
        dense.append(entry)
        compact.append(dict({k: v for k, v in entry.items() if k not in ("mvs", "partitions")}, side=sides))
    return dense, compact


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,t,n", [(64, 96, 7, 2), (720, 1280, 4, 1)])
def test_enhance_clips_takes_compact_side_information(dev, h, w, t, n):
    """driver.enhance_clips on host entries that carry the codec's per-block records (`side`) instead of the dense
    mvs / partitions planes: records are uploaded and rasterised on the device inside the streamer -- the frames equal
    those of the dense feed built by the ORACLE of the reference's loader loop, bit for bit, for n = 2 clips per entry,
    with 1/50th of the side-information bytes over the bus."""
    import pnpvcve_b200 as P
    from pnpvcve_b200 import driver, weights
    from test_gpu_parity import GENERATOR_CFG
    net = P.build_backbone(dict(GENERATOR_CFG, num_blocks=2))
    net.load_state_dict(weights.random_state_dict(2, num_blocks=2), strict=True)
    net = net.to(dev).eval()
    dense, compact = _side_entries(h, w, t, n, 40, R.rasterize_clip)
    outs_d, met_d = driver.enhance_clips(net, dense, device=dev, chunk=3)
    outs_c, met_c = driver.enhance_clips(net, compact, device=dev, chunk=3)
    torch.cuda.synchronize()
    for a, b in zip(outs_d, outs_c):
        assert b.is_pinned() and torch.equal(a, b)
    assert torch.equal(met_d, met_c)
    st = driver.ClipStreamer(net, dev, chunk=3)
    st.upload(dense[0])
    dense_bytes = st.h2d_bytes
    st.upload(compact[0])
    assert st.h2d_bytes < dense_bytes / 2
    st.finish(check=True)
    # malformed records surface as the reference loader's errors once the streamer is finished
    bad = dict(compact[0])
    rec = bad["side"][0]["records"].clone()
    rec[0, 1:3] = 4.0                                     # 4x4 block: area 16 is not in the partition table
    bad["side"] = [dict(bad["side"][0], records=rec.pin_memory())] + list(bad["side"][1:])
    with pytest.raises(KeyError):
        driver.enhance_clips(net, [bad], device=dev)
