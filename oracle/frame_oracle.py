"""TEST INFRASTRUCTURE ONLY — the reference's single-view frame path restated on the CPU:
PkpNet forward (oracle/net_oracle.py) -> keypoint gating (lib/object_slam.py:1100-1115) ->
per-object pnp() (:1123-1165, oracle/geom.py) -> optimize() in single-view mode
(:703-930: camera fixed, one vertex per object, its=[10]*4).  Used by tests/ and by
bench.py's cpu_baseline / --impl reference legs only.

Pinned: tests/golden/slam_seq.npz ("sv_*") holds what the UNMODIFIED reference ObjectSLAM(single_view_mode=True).process_view leaves in
its state for two marker frames with real outliers (oracle/gen_golden_slam.py: lib/object_slam.py imported from /root/reference, its two
native extension modules replaced by oracle/geom.py's PnP / LM); tests/test_marker_cpu.py replays them here: gating, PnP acceptance and
BA inlier sets identical, poses 1e-6 of the scene scale."""
from __future__ import annotations

import numpy as np
import torch

from . import geom, net_oracle


def solve_from_keypoints(uv, cov, kp_mask, model_kps, model_mask, K_bbox, diameter, box_img, kp_var_thresh=0.2,
                         bbox_thresh=0.9, seed=0, run_ba=True):
    """Everything after the network, on given uv/cov/kp_mask (numpy).  Mirrors suo_frames' outputs."""
    L, K = model_mask.shape
    used = net_oracle.gate_keypoints(uv, cov, kp_mask, model_mask.astype(bool), bbox_thresh, kp_var_thresh)
    T_pnp = np.tile(np.eye(4), (L, 1, 1))
    accepted = np.zeros(L, bool)
    for c in range(L):
        m = used[c]
        res = geom.pnp(model_kps[c][m].astype(np.float64), uv[c][m].astype(np.float64), K_bbox[c], seed=seed, obj_key=c)
        if res is not None:
            T_pnp[c, :3] = res[0]
            accepted[c] = res[0][2, 3] > 0.5 * diameter[c] and m.sum() >= 4       # object_slam.py:1147-1148
    T_ba = np.tile(np.eye(4)[:3], (L, 1, 1))
    ba_inl = np.zeros((L, K), bool)
    if run_ba:
        for f in np.unique(box_img):
            crops = [c for c in np.nonzero(box_img == f)[0] if accepted[c]]
            if not crops:
                continue
            n = len(crops)
            poses = np.zeros((n + 1, 3, 4))
            poses[n, :, :3] = np.eye(3)
            fixed = np.zeros(n + 1, np.uint8)
            fixed[n] = 1
            e_obj, cam_k, p, uvs, info, src = [], [], [], [], [], []
            for i, c in enumerate(crops):
                poses[i] = T_pnp[c, :3]
                Kb = K_bbox[c]
                for k in np.nonzero(used[c])[0]:
                    e_obj.append(i)
                    cam_k.append([Kb[0, 0], Kb[1, 1], Kb[0, 2], Kb[1, 2]])
                    p.append(model_kps[c, k])
                    uvs.append(uv[c, k].astype(np.float64))
                    S = cov[c, k].astype(np.float64)
                    det = S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]
                    info.append([S[1, 1] / det, -S[0, 1] / det, -S[1, 0] / det, S[0, 0] / det])
                    src.append((c, k))
            P, inl, _ = geom.ba_optimize(poses, fixed, np.asarray(e_obj, np.int32), np.full(len(e_obj), n, np.int32),
                                         np.asarray(cam_k), np.asarray(p), np.asarray(uvs), np.asarray(info),
                                         np.ones(len(e_obj)), [10, 10, 10, 10])
            for i, c in enumerate(crops):
                T_ba[c] = P[i]
            for (c, k), v in zip(src, inl):
                ba_inl[c, k] = v
    return dict(T_pnp=T_pnp, T_ba=T_ba, kp_used=used, ba_inliers=ba_inl, accepted=accepted)


@torch.no_grad()
def run_frames(sd, images, boxes, box_img, model_kps, model_mask, K_bbox, diameter, input_res=(256, 256), **kw):
    """Full CPU frame path (network included)."""
    boxes_l = [torch.as_tensor(boxes[box_img == i], dtype=torch.float32) for i in range(images.shape[0])]
    out = net_oracle.pkpnet_forward(sd, torch.as_tensor(images), boxes_l, None, input_res)
    uv, cov, km = out["uv"].numpy(), out["cov"].numpy(), out["kp_mask"].numpy()
    res = solve_from_keypoints(uv, cov, km, model_kps, model_mask, K_bbox, diameter, box_img, **kw)
    res.update(uv=uv, cov=cov, kp_mask=km, logits=out["prob_logits"].numpy())
    return res
