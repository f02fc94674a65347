"""Golden values of the reference's OWN psnr / ssim / tensor2img (build container only; needs /root/reference, cv2).

The function sources are READ from the reference at run time (mmedit/core/evaluation/metrics.py, mmedit/core/misc.py)
and executed in a namespace holding numpy / cv2 / torch -- the modules themselves import mmcv and the whole mmedit
package and cannot be imported here.  Nothing is copied into this repository.
    python tests/golden/make_golden_metrics.py
"""
import ast
import math
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PNP_REFERENCE_ROOT", "/root/reference")


def load_functions(path, names, ns):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path + ":" + node.name, "exec"), ns)
    return ns


sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import metric_cases  # noqa: E402


def main():
    ns = dict(np=np, cv2=cv2, torch=torch, math=math, make_grid=None, mmcv=None)
    load_functions(os.path.join(REF, "mmedit/core/evaluation/metrics.py"), {"reorder_image", "psnr", "_ssim", "ssim"}, ns)
    load_functions(os.path.join(REF, "mmedit/core/misc.py"), {"tensor2img"}, ns)
    res = {}
    for name, (out, gt, crop) in metric_cases().items():
        o8, g8 = ns["tensor2img"](out[None]), ns["tensor2img"](gt[None])          # (1,3,H,W) -> HWC BGR uint8
        p = ns["psnr"](o8, g8, crop)
        s = ns["ssim"](o8, g8, crop)
        res[name + "/psnr"] = np.float64(p)
        res[name + "/ssim"] = np.float64(s)
        res[name + "/u8sum"] = np.int64(o8.astype(np.int64).sum() * 1000003 + g8.astype(np.int64).sum())
        print(name, o8.shape, p, s)
    np.savez(os.path.join(HERE, "metrics_cases.npz"), **res)


if __name__ == "__main__":
    main()
