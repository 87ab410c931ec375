"""Bundle-adjustment entry points: packed-array form of ObjectSLAM.optimize()'s g2o solve
(lib/object_slam.py:703-930).  ``ba_batch`` is the thin wrapper over suo_ba_batch;
``optimize_single_view`` builds the graph the reference builds for one frame in
single-view mode (camera fixed at the first view = identity world, one vertex per object,
one EdgeSE3ProjectFromObject per gated keypoint, information = inv(cov), :790-837)."""
from __future__ import annotations

import numpy as np

from . import _lib, runtime

HUBER_DELTA = float(np.sqrt(5.991))   # object_slam.py:831
CHI2_GATE = 5.991                     # object_slam.py:860,886


def ba_batch(prob_vert, prob_edge, poses, fixed, e_obj, e_cam, cam_k, p, uv, info, inliers, its,
             huber_delta=HUBER_DELTA, chi2_gate=CHI2_GATE, init_with_outliers=False, ctx=None, return_errors=False):
    """return_errors: also return err [n_edges,2], each edge's error as g2o would hold it after optimize() (the error of
    the last evaluated LM trial, rejected or not — what e.chi2() reads at lib/object_slam.py:881-883)."""
    ctx = ctx or runtime.get_context()
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    prob_vert, prob_edge = c(prob_vert, np.int32), c(prob_edge, np.int32)
    poses = c(poses, np.float64).reshape(-1, 12).copy()
    fixed = c(fixed, np.uint8)
    e_obj, e_cam = c(e_obj, np.int32), c(e_cam, np.int32)
    cam_k, p = c(cam_k, np.float64).reshape(-1, 4), c(p, np.float64).reshape(-1, 3)
    uv, info = c(uv, np.float64).reshape(-1, 2), c(info, np.float64).reshape(-1, 4)
    inl = c(inliers, np.uint8).copy()
    its = c(its, np.int32)
    n_prob = len(prob_vert) - 1
    stats = np.zeros((n_prob, 3), np.int32)
    ctx.check(_lib.lib().suo_ba_batch(
        ctx.handle, n_prob, _lib.ptr(prob_vert), _lib.ptr(prob_edge), _lib.ptr(poses), _lib.ptr(fixed), len(poses),
        _lib.ptr(e_obj), _lib.ptr(e_cam), _lib.ptr(cam_k), _lib.ptr(p), _lib.ptr(uv), _lib.ptr(info), _lib.ptr(inl),
        len(e_cam), _lib.ptr(its), len(its), float(huber_delta), float(chi2_gate), int(init_with_outliers),
        _lib.ptr(stats), 0, None))
    if return_errors:
        err = np.zeros((len(e_cam), 2))
        if len(e_cam):
            ctx.check(_lib.lib().suo_ba_last_errors(ctx.handle, _lib.ptr(err), len(e_cam), 0, None))
        return poses.reshape(-1, 3, 4), inl.astype(bool), stats, err
    return poses.reshape(-1, 3, 4), inl.astype(bool), stats


def single_view_graph(obj_poses, model_kps, uvs, covs, K_bboxes):
    """Pack one frame: objects 0..n-1 then the fixed identity camera (vertex n)."""
    n = len(obj_poses)
    poses = np.zeros((n + 1, 3, 4))
    for i, T in enumerate(obj_poses):
        poses[i] = np.asarray(T)[:3, :4]
    poses[n, :, :3] = np.eye(3)
    fixed = np.zeros(n + 1, np.uint8)
    fixed[n] = 1
    e_obj, cam_k, p, uv, info = [], [], [], [], []
    for i in range(n):
        Kb = K_bboxes[i]
        for k in range(len(uvs[i])):
            e_obj.append(i)
            cam_k.append([Kb[0, 0], Kb[1, 1], Kb[0, 2], Kb[1, 2]])        # object_slam.py:799
            p.append(model_kps[i][k])
            uv.append(uvs[i][k])
            info.append(np.eye(2) if covs is None else np.linalg.inv(covs[i][k]))   # :825-828 (no clamping)
    e_obj = np.asarray(e_obj, np.int32)
    return dict(poses=poses, fixed=fixed, e_obj=e_obj, e_cam=np.full(len(e_obj), n, np.int32),
                cam_k=np.asarray(cam_k), p=np.asarray(p), uv=np.asarray(uv), info=np.asarray(info).reshape(-1, 4))
