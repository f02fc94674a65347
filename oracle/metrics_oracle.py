"""CPU restatement of the reference's per-frame quality metrics -- TEST INFRASTRUCTURE ONLY (only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import it; the product path is pnp_frame_quality on the GPU).

Follows, for one frame pair (BasicVSR.evaluate, mmedit/models/restorers/basicvsr.py:119-153):
  * tensor2img   mmedit/core/misc.py:9-74      clamp to [0,1], x255, round half to even, uint8, CHW RGB -> HWC BGR
  * psnr         mmedit/core/evaluation/metrics.py:170-215   float32 images, mean squared error, 20 log10(255/sqrt(mse))
  * ssim/_ssim   mmedit/core/evaluation/metrics.py:262-355   float64, 11x11 Gaussian window (sigma 1.5) as the outer
                 product of cv2.getGaussianKernel(11, 1.5), correlation, valid part [5:-5, 5:-5], mean over the map,
                 mean over the 3 channels
convert_to=None only (the shipped configs do not convert to Y).  Parity PINNED: tests/golden/metrics_cases.npz holds
the values of the reference's own functions (executed from /root/reference by tests/golden/make_golden_metrics.py).
"""
import numpy as np
from scipy.ndimage import correlate


def tensor2img_u8(x):
    """x: (3,H,W) float array/tensor in RGB -> (H,W,3) uint8 BGR.  misc.py:52-71"""
    a = np.asarray(x, dtype=np.float32)
    a = np.clip(a, 0.0, 1.0)
    a = (a - 0.0) / (1.0 - 0.0)
    a = np.transpose(a[[2, 1, 0], :, :], (1, 2, 0))
    return (a * 255.0).round().astype(np.uint8)


def gaussian_kernel_11():
    """cv2.getGaussianKernel(11, 1.5): exp(-(i-5)^2 / (2 sigma^2)) normalised to sum 1, float64."""
    i = np.arange(11, dtype=np.float64) - 5.0
    g = np.exp(-(i * i) / (2.0 * 1.5 * 1.5))
    return g / g.sum()


def _crop(img, crop_border):
    if crop_border != 0:
        # metrics.py:208-210 indexes with a trailing None: (H-2c, W-2c, 3, 1); the extra axis changes no value
        img = img[crop_border:-crop_border, crop_border:-crop_border]
    return img


def psnr(img1, img2, crop_border=0):
    """metrics.py:170-215 (input_order HWC, convert_to None)."""
    a = _crop(img1.astype(np.float32), crop_border)
    b = _crop(img2.astype(np.float32), crop_border)
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return float("inf")
    return 20.0 * np.log10(255.0 / np.sqrt(mse))


def sse_u8(img1, img2, crop_border=0):
    """Exact integer sum of squared differences of the uint8 images (what the GPU kernel accumulates)."""
    a = _crop(img1.astype(np.int64), crop_border)
    b = _crop(img2.astype(np.int64), crop_border)
    return int(((a - b) ** 2).sum())


def _ssim(c1, c2):
    """metrics.py:262-291 for one channel (float64)."""
    C1 = (0.01 * 255) ** 2
    C2 = (0.03 * 255) ** 2
    c1 = c1.astype(np.float64)
    c2 = c2.astype(np.float64)
    k = gaussian_kernel_11()
    window = np.outer(k, k)

    def filt(img):     # cv2.filter2D(img, -1, window): correlation, BORDER_REFLECT_101; then the valid part
        return correlate(img, window, mode="mirror")[5:-5, 5:-5]

    mu1, mu2 = filt(c1), filt(c2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    s1 = filt(c1 ** 2) - mu1_sq
    s2 = filt(c2 ** 2) - mu2_sq
    s12 = filt(c1 * c2) - mu1_mu2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean()


def ssim(img1, img2, crop_border=0):
    """metrics.py:294-355 (input_order HWC, convert_to None): mean of the per-channel SSIM.

    Reference quirk, kept: with crop_border != 0 the crop is written ``img[c:-c, c:-c, None]`` (:347-349), which
    inserts an axis -> (H-2c, W-2c, 1, 3); the channel loop then runs over ``shape[2] == 1`` and ``img[..., 0]``
    picks channel 0 of the LAST axis, i.e. only the first channel of the BGR image (blue) is evaluated.
    """
    a, b = _crop(img1, crop_border), _crop(img2, crop_border)
    channels = range(a.shape[2]) if crop_border == 0 else [0]
    return float(np.mean([_ssim(a[..., i], b[..., i]) for i in channels]))
