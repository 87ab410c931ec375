"""TEST INFRASTRUCTURE ONLY — numpy restatement of the hypothesis scoring between the hot-path calls of a SLAM-mode
frame: ObjectSLAM.__estimate_camera_pose (reference lib/object_slam.py:975-1072) and
__maybe_reinit_objects (:595-697).  Only tests/ may import this; the product path never does.

Pinned through oracle/slam_frame_oracle.py: tests/golden/slam_seq.npz holds what the UNMODIFIED reference class reaches on marker sequences
(oracle/gen_golden_slam.py imports lib/object_slam.py with its two native extension modules replaced by the oracle's PnP / LM); one of the
sequences corrupts a map pose so that the vote has to reject an object and __maybe_reinit_objects has to replace it — the restatement takes
the same decisions and reaches the same poses (tests/test_marker_cpu.py).  It follows the cited lines, including the float32 staging of
poses and np.linalg.inv on the float32 covariances."""
from __future__ import annotations

import numpy as np


def _inv_se3(T):   # utils.invert_SE3, lib/utils/utils.py:431-435
    out = np.eye(4, dtype=T.dtype)
    out[:3, :3] = T[:3, :3].T
    out[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return out


def _as44(T, dtype=np.float64):
    out = np.zeros((4, 4), dtype)
    out[:3] = np.asarray(T)[:3]
    out[3, 3] = 1
    return out


def chi2_of(T_OtoC, det, use_inliers, manual_kp_std):
    """:1040-1064 / :655-679 for one (pose, detection): chi2 of the keypoints in front of the camera."""
    sel = np.asarray(det["inliers"], bool) if use_inliers else np.ones(len(det["uv_pred"]), bool)
    if sel.sum() == 0:
        return np.zeros(0)
    p = np.asarray(det["model_kp"])[sel] @ T_OtoC[:3, :3].T + T_OtoC[:3, 3]          # utils.transform_pts
    h = p @ np.asarray(det["K"]).T
    front = h[:, 2] > 0
    proj = (h[:, :2] / h[:, 2:3])[front]
    if len(proj) == 0:
        return np.zeros(0)
    res = np.asarray(det["uv_pred"])[sel][front] - proj
    if det.get("cov_pred") is not None:
        cov = np.array(np.asarray(det["cov_pred"])[sel][front])
        cov[:, [0, 1], [0, 1]] = np.maximum(cov[:, [0, 1], [0, 1]], 1e-4)
        inf = np.linalg.inv(cov)
    else:
        inf = np.zeros((len(res), 2, 2), np.float32)
        inf[:, [0, 1], [0, 1]] = 1 / manual_kp_std ** 2
    return (res[:, None, :] @ inf @ res[:, :, None]).reshape(-1)


def estimate_camera_pose(obj_poses, curr_det, min_num_inliers=4, manual_kp_std=0.05):
    """-> (T_GtoC or None, per-hypothesis inlier counts, all chi2 values seen)."""
    ids = [o for o in curr_det if curr_det[o].get("pose") is not None and o in obj_poses]
    if not ids:
        return None, None, np.zeros(0)
    hyp = np.stack([np.asarray(curr_det[o]["pose"]) @ _inv_se3(_as44(obj_poses[o])) for o in ids])
    OtoG = np.stack([_as44(obj_poses[o], np.float32) for o in ids])
    counts, seen = np.zeros(len(ids), int), []
    for i in range(len(ids)):
        for j, o in enumerate(ids):
            c = chi2_of(hyp[i] @ OtoG[j], curr_det[o], True, manual_kp_std)
            counts[i] += int((c <= 5.991).sum())
            seen.append(c)
    best, best_n = None, -1
    for i in range(len(ids)):
        if counts[i] >= min_num_inliers and counts[i] > best_n:
            best, best_n = hyp[i], counts[i]
    return best, counts, np.concatenate(seen) if seen else np.zeros(0)


def maybe_reinit_objects(obj_poses, cam_poses, detections, view_ids, view_id, check_n_views=15, manual_kp_std=0.05):
    """-> ({obj_id: new T_OtoG}, {obj_id: {"pnp": n, "estim": n}}, all chi2 values seen)."""
    if len(view_ids) < 2 or view_id not in cam_poses:
        return {}, {}, np.zeros(0)
    n_views = min(len(view_ids), check_n_views)
    cur = detections[view_id]
    ids = [o for o in obj_poses if cur.get(o, {}).get("pose") is not None]
    views = [view_ids[-(i + 1)] for i in range(n_views)]
    cam_inv = _inv_se3(_as44(cam_poses[view_id]))
    new, num, seen = {}, {}, []
    for o in ids:
        T_pnp_G = cam_inv @ np.asarray(cur[o]["pose"])
        T_est = _as44(obj_poses[o], np.float32)
        n = {"pnp": 0, "estim": 0}
        for v in views:
            if o not in detections[v]:
                continue
            GtoC = _as44(cam_poses[v], np.float32)
            for key, T in (("pnp", T_pnp_G), ("estim", T_est)):
                c = chi2_of(GtoC @ T, detections[v][o], False, manual_kp_std)
                n[key] += int((c <= 5.991).sum())
                seen.append(c)
        num[o] = n
        if n["pnp"] >= 3 and n["pnp"] > 3 * n["estim"]:
            new[o] = T_pnp_G
    return new, num, np.concatenate(seen) if seen else np.zeros(0)
