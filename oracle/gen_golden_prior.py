"""TEST INFRASTRUCTURE ONLY — generates tests/golden/prior.npz by running the UNMODIFIED reference
``utils.make_prior_kp_input`` (lib/utils/utils.py:398-411, imported read-only from /root/reference through
oracle/ref_shims.py) on seeded keypoints, including the awkward ones: exact .5 pixel ties (Python's round() is
half-to-even), NDC values on and beyond +-1 (clamped), NaN / inf (skipped), masked-out keypoints, windows clipped
by every image border, pixel-unit input (ndc=False).  Run in the build container:

    python -m oracle.gen_golden_prior
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "prior.npz")


def cases():
    rng = np.random.default_rng(5)
    out = {}
    # 256x256 NDC (what ObjectSLAM uses, lib/object_slam.py:513-514), float32 like prior_uv_full (:510)
    uv = rng.uniform(-1.05, 1.05, size=(2, 12, 2)).astype(np.float32)
    uv[0, 0] = [(100.5 + 0.5 - 128) / 128, (128 - 0.5 - 60.5) / 128 - 0.0]     # pixel (100.5, 60.5): ties -> (100, 60)
    uv[0, 1] = [(101.5 + 0.5 - 128) / 128, (255.5 - 61.5 - 128) / 128]         # ties -> (102, 62)
    uv[0, 2] = [1.0, -1.0]
    uv[0, 3] = [-1.0, 1.0]
    uv[0, 4] = [3.0, -7.0]                                                     # clamped
    uv[0, 5] = [np.nan, 0.1]
    uv[0, 6] = [0.2, np.inf]
    uv[1, 0] = [0.0, 0.0]
    mask = rng.random((2, 12)) < 0.8
    mask[0, :7] = True
    mask[0, 7] = False
    out["a"] = (uv, mask, (256, 256), True)
    # 64x64 NDC (the small network of the other goldens): the 90x90 window always overhangs
    uv = rng.uniform(-1.0, 1.0, size=(1, 8, 2)).astype(np.float32)
    out["b"] = (uv, np.ones((1, 8), bool), (64, 64), True)
    # pixel coordinates, non-square image, windows partly / entirely outside
    uv = np.array([[[10.2, 20.7], [-44.0, 5.0], [-46.0, 5.0], [139.5, 99.5], [186.0, 50.0], [60.0, -45.4], [60.0, 146.0]]], np.float32)
    out["c"] = (uv, np.ones((1, 7), bool), (100, 140), False)
    return out


def main():
    ref_utils = ref_shims.import_reference_utils()
    blob = {}
    for name, (uv, mask, shape, ndc) in cases().items():
        planes = np.stack([ref_utils.make_prior_kp_input(uv[i], mask[i], shape, ndc=ndc) for i in range(len(uv))])
        blob[name + "_uv"], blob[name + "_mask"], blob[name + "_shape"], blob[name + "_ndc"], blob[name + "_planes"] = uv, mask, np.array(shape), ndc, planes
        print(name, planes.shape, int((planes > 0).sum()))
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
