"""Golden vectors for the MV / partition rasteriser, produced by the reference's OWN loop.

Run in the build container only:   python tests/golden/make_golden_raster.py

The per-record loop of LoadImageFromFileList_ipb.__call__ (mmedit/datasets/pipelines/loading_ipb.py, the
`for idx in range(mv_npy.shape[0])` block and the p_offset update) is embedded in file / JSON handling that
needs mmcv and a dataset on disk, so its source lines are READ from /root/reference at run time (never
copied into this repository), de-indented and executed on seeded synthetic records.
"""
import os
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.environ.get("PNP_REFERENCE_ROOT", "/root/reference"),
                   "mmedit/datasets/pipelines/loading_ipb.py")

from pnpvcve_b200 import sideinfo  # noqa: E402


def reference_loop_source():
    lines = open(REF).read().splitlines()
    start = next(i for i, l in enumerate(lines) if "for idx in range(mv_npy.shape[0]):" in l and i > 300)
    end = next(i for i in range(start, len(lines)) if "p_offset = p_offset + 1 if is_B_frame else 1" in lines[i])
    body = lines[start:end + 1]        # loop, `partitions.append`, `mvs.append`, p_offset update
    src = textwrap.dedent("\n".join(body))
    return src


class _Self:
    load_partition = True
    drconv = True


def run_reference(records, slice_types, h_img, w_img):
    src = reference_loop_source()
    code = compile(src, REF + ":mv-loop", "exec")
    mvs, partitions = [], []
    ns = dict(np=np, self=_Self(), mvs=mvs, partitions=partitions)
    for rec, st in zip(records, slice_types):
        ns["mv_npy"] = np.asarray(rec, dtype=np.float32).reshape(-1, 10).astype(np.float32)
        ns["is_B_frame"] = (st == "B")
        ns["mv"] = np.zeros((h_img, w_img, 4)).astype(np.float32)
        ns["partition"] = np.zeros((h_img, w_img, 3)).astype(np.float32)
        ns["partition_ch"] = {'256': 0, '128': 1, '64': 2}
        ns["h"], ns["w"] = h_img, w_img
        exec(code, ns)            # runs the loop, `partitions.append`, `mvs.append`, p_offset update
    mv_out = np.stack(mvs, 0).transpose(0, 3, 1, 2).copy().astype(np.float32)
    par_out = (np.stack(partitions, 0).astype(np.float32) / 255).transpose(0, 3, 1, 2).copy()
    return mv_out, par_out


CASES = {
    "raster_64x96_ibbp": dict(h=64, w=96, pattern="IBBPBBP", seed=11, messy=False),
    "raster_72x80_messy": dict(h=72, w=80, pattern="IPBBPBP", seed=12, messy=True),
    "raster_128x128_ip": dict(h=128, w=128, pattern="IPPP", seed=13, messy=True),
}


def main():
    print(reference_loop_source())
    for name, c in CASES.items():
        recs = sideinfo.synthetic_records(c["h"], c["w"], c["pattern"], seed=c["seed"], messy=c["messy"])
        mv, par = run_reference(recs, list(c["pattern"]), c["h"], c["w"])
        flat = np.concatenate([r.reshape(-1, 10) for r in recs], 0).astype(np.float32)
        offs = np.cumsum([0] + [len(r) for r in recs]).astype(np.int32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), records=flat, offsets=offs,
                            pattern=np.array(list(c["pattern"])), mvs=mv, partitions=par,
                            shape=np.array([c["h"], c["w"]]))
        print(name, flat.shape, float(np.abs(mv).max()), float(par.sum()))


if __name__ == "__main__":
    main()
