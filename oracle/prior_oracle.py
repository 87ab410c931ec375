"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's prior-detection heat maps
(SURVEY.md §8 row f2): utils.gaussian_2d / draw_gaussian_2d / make_prior_kp_input, reference
lib/utils/utils.py:356-411.  Only tests/ may import this; the product path (suo_slam_b200/) never does.

The one third-party call of that code, ``cv2.GaussianBlur`` (OpenCV, present in this image as it is in the
reference's requirements.txt), is called here exactly as the reference calls it.

PARITY PIN: tests/test_oracle_net.py checks this file against tests/golden/prior.npz, produced by the UNMODIFIED
reference function (oracle/gen_golden_prior.py), and live against the reference when /root/reference is mounted.
"""
from __future__ import annotations

import math

import cv2
import numpy as np


def gaussian_2d(size: int) -> np.ndarray:
    """utils.py:356-361: blur a centred delta with a size x size kernel (sigma derived by OpenCV), max-normalised."""
    assert size % 2 == 1
    g = np.zeros((size, size), dtype=np.float32)
    g[size // 2, size // 2] = 1
    g = cv2.GaussianBlur(g, (size, size), 0)
    return g / np.max(g)


def draw_gaussian_2d(img: np.ndarray, pt, sigma: int = 15) -> np.ndarray:
    """utils.py:364-385: paste (assign, not add) the stamp around pt = (x, y); the slice ends are exclusive, so
    the window is 2*tmpSize wide although the stamp is 2*tmpSize + 1."""
    h, w = img.shape
    t = int(math.ceil(3 * sigma))
    ul = (int(math.floor(pt[0] - t)), int(math.floor(pt[1] - t)))
    br = (int(math.floor(pt[0] + t)), int(math.floor(pt[1] + t)))
    if ul[0] > w or ul[1] > h or br[0] < 1 or br[1] < 1:
        return img
    g = gaussian_2d(2 * t + 1)
    gx0, gy0 = max(0, -ul[0]), max(0, -ul[1])
    x0, x1 = max(0, ul[0]), min(br[0], w)
    y0, y1 = max(0, ul[1]), min(br[1], h)
    img[y0:y1, x0:x1] = g[gy0:gy0 + (y1 - y0), gx0:gx0 + (x1 - x0)]
    return img


def make_prior_kp_input(kp_uv: np.ndarray, kp_uv_mask: np.ndarray, img_shape, ndc: bool = True) -> np.ndarray:
    """utils.py:398-411.  The pixel arithmetic runs in the dtype of kp_uv, as NumPy does for the reference
    (float32 for ObjectSLAM's prior_uv_full, lib/object_slam.py:510); round() is half-to-even."""
    n = kp_uv.shape[0]
    vh, vw = int(img_shape[0]), int(img_shape[1])
    x = np.zeros((n, vh, vw), dtype=np.float32)
    ft = kp_uv.dtype.type if kp_uv.dtype.kind == "f" else np.float64
    for i in range(n):
        if not kp_uv_mask[i] or not np.all(np.isfinite(kp_uv[i, :2])):
            continue
        u, v = ft(kp_uv[i, 0]), ft(kp_uv[i, 1])
        if ndc:
            cu, cv = min(max(u, ft(-1)), ft(1)), min(max(v, ft(-1)), ft(1))
            u = ft(ft(ft(cu * ft(vw)) / ft(2)) + ft(vw / 2)) - ft(0.5)
            v = ft(vh - 0.5) - ft(ft(ft(cv * ft(vh)) / ft(2)) + ft(vh / 2))
        pt = (int(round(float(u))), int(round(float(v))))
        draw_gaussian_2d(x[i], pt)
    return x
