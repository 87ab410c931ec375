"""TEST INFRASTRUCTURE ONLY — generates tests/golden/net_*.npz by running the
UNMODIFIED reference Python (imported read-only from /root/reference with the
shims in oracle/ref_shims.py).  Run in the build container:

    python -m oracle.gen_golden_net

The fixtures pin oracle/net_oracle.py (and through it the CUDA path):
  net_small.npz    reference PkpNet(input_res=(64,64)) on 3 crops of a 120x160
                   frame, seeded synthetic weights (suo_slam_b200.synth), no prior
  net_prior.npz    same with non-zero prior planes from the reference's
                   utils.make_prior_kp_input (lib/utils/utils.py:398-411)
  reduce.npz       reference spatial_softmax + post_process_kp on seeded logits
  kbbox.npz        reference utils.fix_K_for_bbox_ndc
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from suo_slam_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def small_case(seed, with_prior, ref_utils):
    rng = np.random.default_rng(seed)
    H, W = 120, 160
    img = rng.random((1, 3, H, W), dtype=np.float32)
    boxes = np.array([[10.0, 8.0, 90.0, 100.0], [40.5, 20.25, 150.0, 110.0], [100.0, 30.0, 112.0, 45.0]],
                     dtype=np.float32)
    prior = None
    if with_prior:
        prior = np.zeros((3, 41, 64, 64), dtype=np.float32)
        for k in (0, 2):
            uv = rng.uniform(-1.1, 1.1, size=(41, 2))
            m = rng.random(41) < 0.4
            prior[k] = ref_utils.make_prior_kp_input(uv, m, (64, 64))
    return img, boxes, prior


def kbbox_f32():
    """kbbox_f32.npz: the reference's utils.fix_K_for_bbox_ndc on FLOAT32 bboxes (what lib/datasets/bop.py:540-552 produces and
    lib/object_slam.py:1086 passes) with fractional coordinates, as the reference stores it (float32, :1082) — the float32 scalar
    arithmetic of ``x2 - x1`` and ``2.0 / w`` is part of the result.  `python -m oracle.gen_golden_net --kbbox-f32` writes only this file."""
    ref_utils = ref_shims.import_reference_utils()
    rng = np.random.default_rng(12)
    bbs = np.stack([np.array([x, y, x + w, y + h], np.float32) for x, y, w, h in
                    zip(rng.uniform(0, 400, 64), rng.uniform(0, 300, 64), rng.uniform(20, 240, 64), rng.uniform(20, 180, 64))])
    raw = np.stack([ref_utils.fix_K_for_bbox_ndc(synth.K_YCBV, bb) for bb in bbs])
    f32 = np.zeros((len(bbs), 3, 3), np.float32)
    for k, bb in enumerate(bbs):
        f32[k] = ref_utils.fix_K_for_bbox_ndc(synth.K_YCBV, bb)
    np.savez(os.path.join(OUT, "kbbox_f32.npz"), K=synth.K_YCBV, bbox=bbs, K_bbox=raw, K_bbox_f32=f32, numpy=np.__version__)
    print("wrote kbbox_f32.npz (numpy", np.__version__ + ")")


def main():
    if "--kbbox-f32" in sys.argv:
        return kbbox_f32()
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shims.import_reference_pkpnet()
    ref_utils = ref_shims.import_reference_utils()
    torch.manual_seed(0)
    torch.set_num_threads(8)

    sd = synth.make_synthetic_state_dict(seed=0, peaky=4.0)
    net = ref.PkpNet(input_res=(64, 64), calc_cov=True)
    net.load_state_dict(sd, strict=True)  # proves key/shape compatibility with the reference
    net.eval()
    for name, with_prior in (("net_small", False), ("net_prior", True)):
        img, boxes, prior = small_case(7, with_prior, ref_utils)
        with torch.no_grad():
            out = net(torch.from_numpy(img), [torch.from_numpy(boxes)],
                      None if prior is None else [torch.from_numpy(prior)])
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"), img=img, boxes=boxes,
            prior=(np.zeros(0, np.float32) if prior is None else prior),
            uv=out["uv"].numpy(), cov=out["cov"].numpy(), logits=out["prob_logits"].numpy(),
            kp_mask=out["kp_mask"].numpy(), kp_mask_logits=out["kp_mask_logits"].numpy(),
            weights_seed=0, peaky=4.0)
        print(name, out["uv"].shape, float(out["prob_logits"].abs().max()))

    # heat-map reduction alone (a3): peaked bumps + noise, and a flat map
    rng = np.random.default_rng(11)
    logits = rng.normal(scale=0.5, size=(2, 41, 32, 32)).astype(np.float32)
    for b in range(2):
        for k in range(41):
            cy, cx = rng.integers(2, 30, size=2)
            yy, xx = np.mgrid[0:32, 0:32]
            logits[b, k] += (8.0 * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 8.0)).astype(np.float32)
    logits[1, 5] = 0.0  # flat map: uv = 0, cov = var U(-1,1)
    t = torch.from_numpy(logits)
    prob = ref.spatial_softmax(t)
    pp = ref.post_process_kp(prob, z=None, calc_sigma=True)
    np.savez_compressed(os.path.join(OUT, "reduce.npz"), logits=logits, prob=prob.numpy(),
                        uv=pp["uv"].numpy(), cov=pp["cov"].numpy())

    # fix_K_for_bbox_ndc
    Ks, bbs, outs = [], [], []
    for i in range(6):
        bb = np.array([10.0 + 7 * i, 20.0 + 3 * i, 200.0 + 11 * i, 230.0 + 5 * i])
        Ks.append(synth.K_YCBV)
        bbs.append(bb)
        outs.append(ref_utils.fix_K_for_bbox_ndc(synth.K_YCBV, bb))
    np.savez(os.path.join(OUT, "kbbox.npz"), K=np.array(Ks), bbox=np.array(bbs), K_bbox=np.array(outs))
    kbbox_f32()
    print("done")


if __name__ == "__main__":
    main()
