"""Hypothesis scoring between the three hot-path calls of a SLAM-mode frame: the camera-pose
vote (ObjectSLAM.__estimate_camera_pose, reference lib/object_slam.py:975-1072) and the re-initialisation test
(__maybe_reinit_objects, :595-697).  Both are nested Python loops around one primitive — count the keypoints a pose
hypothesis explains (chi2 <= 5.991) — which runs here as ONE kernel launch over all (hypothesis, detection) pairs
(``suo_chi2_inlier_counts``, csrc/frames.cu).  The few 4x4 products that compose the hypotheses stay in numpy, written
exactly as the reference writes them (including its float32 staging of the object / camera poses).

A *detection* is the reference's ``detections[view][obj]`` dict: keys ``pose`` (4x4 T_OtoC from PnP or None),
``model_kp`` [n,3], ``K`` [3,3], ``uv_pred`` [n,2], ``cov_pred`` [n,2,2] or None, ``inliers`` [n] bool.
"""
from __future__ import annotations

import numpy as np

from . import _lib, runtime

CHI2_GATE = 5.991
BACKUP_KEY = 10 ** 6        # RANSAC stream of the bbox-centroid PnP of __backup_estimate_camera_pose (the crops use 0..L-1)


def invert_SE3(T):
    """utils.invert_SE3 (lib/utils/utils.py:431-435)."""
    T = np.asarray(T)
    out = np.eye(4, dtype=T.dtype)
    out[:3, :3] = T[:3, :3].T
    out[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return out


def _as44(T, dtype=np.float64):
    out = np.zeros((4, 4), dtype)
    out[:3, :] = np.asarray(T)[:3, :]
    out[3, 3] = 1
    return out


def chi2_inlier_counts(T_pairs, pair_det, dets, use_inlier_mask, manual_kp_std=0.05, gate=CHI2_GATE, ctx=None):
    """counts[p] = #keypoints of detection dets[pair_det[p]] explained by T_pairs[p] (a T_OtoC)."""
    ctx = ctx or runtime.get_context()
    T = np.ascontiguousarray(np.asarray(T_pairs, np.float64)[:, :3, :4]).reshape(-1, 12)
    pd = np.ascontiguousarray(pair_det, np.int32)
    off = np.zeros(len(dets) + 1, np.int32)
    off[1:] = np.cumsum([len(d["uv_pred"]) for d in dets])
    mk = np.ascontiguousarray(np.concatenate([np.asarray(d["model_kp"], np.float64).reshape(-1, 3) for d in dets]))
    K = np.ascontiguousarray(np.stack([np.asarray(d["K"], np.float64) for d in dets])).reshape(-1, 9)
    uv = np.ascontiguousarray(np.concatenate([np.asarray(d["uv_pred"], np.float32).reshape(-1, 2) for d in dets]))
    has_cov = [d.get("cov_pred") is not None for d in dets]
    if any(has_cov) and not all(has_cov):
        raise ValueError("either every detection carries cov_pred or none does (ObjectSLAM.no_network_cov is global)")
    cov = np.ascontiguousarray(np.concatenate([np.asarray(d["cov_pred"], np.float32).reshape(-1, 4) for d in dets])) if all(has_cov) else None
    use = np.ascontiguousarray(np.concatenate([np.asarray(d["inliers"]).astype(np.uint8) for d in dets])) if use_inlier_mask else None
    counts = np.zeros(len(pd), np.int32)
    ctx.check(_lib.lib().suo_chi2_inlier_counts(ctx.handle, len(pd), _lib.ptr(T), _lib.ptr(pd), len(dets), _lib.ptr(off), _lib.ptr(mk),
                                                _lib.ptr(K), _lib.ptr(uv), _lib.ptr(cov), _lib.ptr(use), float(manual_kp_std), float(gate),
                                                _lib.ptr(counts), 0, None))
    return counts


def estimate_camera_pose(obj_poses, curr_det, min_num_inliers=4, manual_kp_std=0.05, ctx=None, return_counts=False):
    """ObjectSLAM.__estimate_camera_pose (:975-1072).  obj_poses: {obj_id: T_OtoG [>=3,4]}, curr_det: {obj_id: detection}.
    Every object with a PnP pose votes T_GtoC = T_OtoC_pnp @ inv(T_OtoG); the hypothesis explaining the most keypoints
    over all voting objects wins (first one on ties, and only with >= min_num_inliers).  Returns T_GtoC [4,4] or None."""
    obj_ids = [o for o in curr_det if curr_det[o].get("pose") is not None and o in obj_poses]
    if not obj_ids:
        return (None, None) if return_counts else None
    n = len(obj_ids)
    Ts_GtoO = np.stack([invert_SE3(_as44(obj_poses[o])) for o in obj_ids])
    Ts_OtoG = np.stack([_as44(obj_poses[o], np.float32) for o in obj_ids])          # float32 staging, :1005-1008
    Ts_OtoC_pnp = np.stack([np.asarray(curr_det[o]["pose"]) for o in obj_ids])
    Ts_hyp = Ts_OtoC_pnp @ Ts_GtoO                                                  # :1011
    Ts_OtoC_hyp = Ts_hyp[:, None] @ Ts_OtoG[None]                                   # [hypothesis, object], :1030
    dets = [curr_det[o] for o in obj_ids]
    counts = chi2_inlier_counts(Ts_OtoC_hyp.reshape(n * n, 4, 4), np.tile(np.arange(n, dtype=np.int32), n), dets, True,
                                manual_kp_std, ctx=ctx).reshape(n, n).sum(1)
    best, best_n = None, -1
    for i in range(n):                                                              # :1068-1070
        if counts[i] >= min_num_inliers and counts[i] > best_n:
            best, best_n = Ts_hyp[i], int(counts[i])
    return (best, counts) if return_counts else best


def maybe_reinit_objects(obj_poses, cam_poses, detections, view_ids, view_id, check_n_views=15, manual_kp_std=0.05, ctx=None,
                         return_counts=False):
    """ObjectSLAM.__maybe_reinit_objects (:595-697).  For every mapped object with a PnP pose in the current view, count
    the keypoints explained over the last ``check_n_views`` views by (a) the PnP pose moved to the world frame and
    (b) the current estimate; re-initialise when pnp >= 3 and pnp > 3 * estim.  Returns {obj_id: new T_OtoG [4,4]}
    (the caller assigns them to obj_poses, :687)."""
    if len(view_ids) < 2 or view_id not in cam_poses:
        return ({}, {}) if return_counts else {}
    check_n_views = min(len(view_ids), check_n_views)
    curr_det = detections[view_id]
    obj_ids = [o for o in obj_poses if curr_det.get(o, {}).get("pose") is not None]
    if not obj_ids:
        return ({}, {}) if return_counts else {}
    Ts_estim = np.stack([_as44(obj_poses[o], np.float32) for o in obj_ids])         # :619-622
    Ts_pnp_G = invert_SE3(_as44(cam_poses[view_id]))[None] @ np.stack([np.asarray(curr_det[o]["pose"]) for o in obj_ids])   # :627
    views = [view_ids[-(i + 1)] for i in range(check_n_views)]
    Ts_GtoCi = np.stack([_as44(cam_poses[v], np.float32) for v in views])           # :630-633
    T_key = {"pnp": Ts_GtoCi[:, None] @ Ts_pnp_G[None], "estim": Ts_GtoCi[:, None] @ Ts_estim[None]}
    dets, T_pairs, owner = [], [], []
    for j, o in enumerate(obj_ids):
        for i, v in enumerate(views):
            if o in detections[v]:
                dets.append(detections[v][o])
                for key in ("pnp", "estim"):
                    T_pairs.append(T_key[key][i, j]); owner.append((j, key, len(dets) - 1))
    num = {o: {"pnp": 0, "estim": 0} for o in obj_ids}
    if T_pairs:
        c = chi2_inlier_counts(np.stack(T_pairs), np.array([d for _, _, d in owner], np.int32), dets, False, manual_kp_std, ctx=ctx)
        for (j, key, _), v in zip(owner, c):
            num[obj_ids[j]][key] += int(v)
    new = {o: Ts_pnp_G[j] for j, o in enumerate(obj_ids) if num[o]["pnp"] >= 3 and num[o]["pnp"] > 3 * num[o]["estim"]}   # :683-687
    return (new, num) if return_counts else new


# ---- a whole SLAM-mode view on the device ---------------------------------------------------------------------------
class SlamTracker:
    """Host mirror of ObjectSLAM's per-view tracking state around ONE ``suo_slam_frame`` call per view (lib/object_slam.py:327-421):
    the map ``obj_poses`` {obj: T_OtoG [3,4]}, ``cam_poses`` {view: T_GtoC [3,4]}, ``detections`` {view: {obj: det}} and
    ``view_ids`` — the same containers, with the same detection keys (pose, inliers, kp_mask, model_kp, uv_pred, cov_pred, K, bbox,
    prior_uv), so the reference's own bookkeeping (collect_results, ...) can run on top of it.  The two pieces of optimize() that follow
    the per-view solve are here too: the inlier-count check that ends every optimize() (:913-930) and, every ``global_opt_every`` views,
    the full graph (:443-451, :736-778: cameras and objects free) as ONE coupled ``suo_ba_batch`` problem (csrc/ba_global.cu) — and so is
    __backup_estimate_camera_pose (:933-973) for views without a usable vote (``_backup_camera_pose`` + suo_slam_frame's cam_init_mode).
    tests/test_gpu_slam.py replays seven sequences whose states the UNMODIFIED reference class produced (tests/golden/slam_seq.npz)."""

    def __init__(self, model, kp_var_thresh=0.2, bbox_thresh=0.9, manual_kp_std=0.005, init_with_outliers=False, seed=0, check_n_views=15,
                 global_opt_every=10, sfm_mode=False):
        """sfm_mode (ObjectSLAM(sfm_mode=True)): the re-initialisation test looks at ALL earlier views (:417), the per-view solve runs
        its = [10, 10, 40, 40] (:843-846, SUO_OPT_SLAM_SFM) and the full graph is optimised after EVERY view, the first included (:443)."""
        self.model = model
        self.sfm_mode = bool(sfm_mode)
        self.kp_var_thresh, self.bbox_thresh, self.manual_kp_std = kp_var_thresh, bbox_thresh, manual_kp_std
        self.init_with_outliers, self.seed, self.check_n_views = init_with_outliers, seed, check_n_views
        self.global_opt_every = global_opt_every
        self.obj_poses, self.cam_poses, self.detections, self.view_ids = {}, {}, {}, []
        self.obj_num_dets, self.diameters = {}, {}          # ObjectSLAM.obj_num_dets (:149,1153); mesh_db[obj]["diameter"]
        self.record = None        # set to a list to keep the packed input arrays of every suo_slam_frame call (bench.py replays them from HBM)

    def process_view(self, view_id, img, K, obj_ids, bboxes, model_kps, model_kps_masks, is_sym, diameters, cam_pose=None):
        """img [H,W,3] u8; K [3,3]; per detected object: id, bbox xyxy, model keypoints [41,3], their mask [41], symmetric flag
        (mesh_db[obj]["is_symmetric"]), diameter; cam_pose: optional external T_GtoC [>=3,4] (process_view's cam_pose, :349-353: no vote,
        every object is treated as symmetric).  Returns the call's raw outputs (per crop, in the ORIGINAL object order)."""
        assert view_id not in self.cam_poses, f"Repeat view_id {view_id}"                      # :329-330
        ctx = self.model.context()
        obj_ids = list(obj_ids)
        is_sym = np.ones(len(obj_ids), bool) if cam_pose is not None else np.asarray(is_sym, bool)
        order = np.concatenate([np.nonzero(~is_sym)[0], np.nonzero(is_sym)[0]]).astype(int)    # non-symmetric crops first (:394-418)
        L, Kk = len(order), self.model.num_kp
        c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
        img = c(img, np.uint8)
        H, W = img.shape[:2]
        boxes = c(np.asarray(bboxes)[order], np.float32)
        mk, mm = c(np.asarray(model_kps)[order], np.float64), c(np.asarray(model_kps_masks)[order], np.uint8)
        diam = c(np.asarray(diameters)[order], np.float64)
        ids = [obj_ids[i] for i in order]
        map_valid = c([o in self.obj_poses for o in ids], np.uint8)
        T_map = np.tile(np.eye(4)[:3], (L, 1, 1))
        for q, o in enumerate(ids):
            if o in self.obj_poses:
                T_map[q] = np.asarray(self.obj_poses[o])[:3]
        # the objects' detections in the last check_n_views - 1 earlier views (__maybe_reinit_objects, :627-650)
        hc, hT, hK, hoff, hmk, huv, hcov = [], [], [], [0], [], [], []
        n_check = len(self.view_ids) if self.sfm_mode else min(len(self.view_ids), self.check_n_views - 1)
        for v in [self.view_ids[-(i + 1)] for i in range(n_check)]:
            for q, o in enumerate(ids):
                d = self.detections[v].get(o)
                if d is None:
                    continue
                hc.append(q); hT.append(np.asarray(self.cam_poses[v])[:3]); hK.append(d["K"])
                hmk.append(d["model_kp"]); huv.append(d["uv_pred"]); hcov.append(d["cov_pred"])
                hoff.append(hoff[-1] + len(d["uv_pred"]))
        nh = len(hc)
        cat = lambda xs, dt, shape: c(np.concatenate(xs), dt) if xs and sum(len(x) for x in xs) else np.zeros(shape, dt)
        h = dict(crop=c(hc, np.int32), T=c(hT, np.float64).reshape(-1, 12), K=c(hK, np.float64).reshape(-1, 9), off=c(hoff, np.int32),
                 mk=cat(hmk, np.float64, (1, 3)), uv=cat(huv, np.float32, (1, 2)), cov=cat([x.reshape(-1, 4) for x in hcov], np.float32, (1, 4))) if nh else None
        out = dict(T_GtoC=np.zeros((3, 4)), status=np.zeros(8, np.int32), T_pnp=np.zeros((L, 4, 4)), kp_used=np.zeros((L, Kk), np.uint8),
                   ba_inliers=np.zeros((L, Kk), np.uint8), uv=np.zeros((L, Kk, 2), np.float32), cov=np.zeros((L, Kk, 2, 2), np.float32),
                   prior_uv=np.zeros((L, Kk, 2), np.float32), prior_mask=np.zeros((L, Kk), np.uint8), K_bbox=np.zeros((L, 3, 3)),
                   T_OtoG=np.zeros((L, 3, 4)), map_valid=np.zeros(L, np.uint8), reinit=np.zeros(L, np.uint8), reinit_counts=np.zeros((L, 2), np.int32))
        p = _lib.ptr
        n_views = len(self.view_ids) + 1
        n1 = int((~is_sym).sum())
        if self.record is not None:
            self.record.append(dict(img=img, K=c(K, np.float64), boxes=boxes, L=L, n1=n1, mk=mk, mm=mm, diam=diam, map_valid=map_valid,
                                    T_map=c(T_map, np.float64), n_views=n_views, hist=h))

        ctx.set_option(_lib.SUO_OPT_SLAM_SFM, int(self.sfm_mode))

        def call(T_init, mode):
            T_init = None if T_init is None else c(np.asarray(T_init)[:3], np.float64)
            ctx.check(_lib.lib().suo_slam_frame(
                ctx.handle, p(img), H, W, p(c(K, np.float64)), p(boxes), L, n1, p(mk), p(mm), p(diam), p(map_valid), p(c(T_map, np.float64)),
                n_views, nh, *((p(h["crop"]), p(h["T"]), p(h["K"]), p(h["off"]), p(h["mk"]), p(h["uv"]), p(h["cov"])) if nh else (None,) * 7),
                float(self.kp_var_thresh), float(self.bbox_thresh), float(self.manual_kp_std), int(self.init_with_outliers), int(self.seed),
                p(out["T_GtoC"]), p(out["status"]), p(out["T_pnp"]), p(out["kp_used"]), p(out["ba_inliers"]), p(out["uv"]), p(out["cov"]), p(out["prior_uv"]),
                p(out["prior_mask"]), p(out["K_bbox"]), p(out["T_OtoG"]), p(out["map_valid"]), p(out["reinit"]), p(out["reinit_counts"]),
                p(T_init), int(mode), 0, None))

        backup = None
        if cam_pose is not None:
            call(cam_pose, 1)
        elif self.view_ids and n1 == 0:                    # no non-symmetric object to vote with (:372-391): the backup pose BEFORE the passes
            T_b, backup = self._backup_camera_pose(obj_ids, bboxes, K)
            call(T_b, 1)
        else:
            call(None, 0)
            if not out["status"][0] and self.view_ids:     # the vote failed (:404-411): backup pose, then the symmetric pass from it
                T_b, backup = self._backup_camera_pose(obj_ids, bboxes, K)
                call(T_b, 2)
        cam_ok = bool(out["status"][0])
        det = {}
        for q, o in enumerate(ids):
            if q >= n1 and not cam_ok:                    # symmetric objects leave no detection without a camera pose (:413-418)
                continue
            m = out["kp_used"][q].astype(bool)
            T = out["T_pnp"][q]
            ok = (not np.allclose(T, np.eye(4))) and m.sum() >= 4 and T[2, 3] > 0.5 * diam[q]
            det[o] = dict(pose=T.copy() if ok else None, inliers=out["ba_inliers"][q][m].astype(bool) if out["status"][3] > 0 and out["map_valid"][q] else np.ones(int(m.sum()), bool),
                          kp_mask=m, model_kp=mk[q][m], uv_pred=out["uv"][q][m].astype(np.float64), cov_pred=out["cov"][q][m], K=out["K_bbox"][q].copy(),
                          bbox=boxes[q], prior_uv=out["prior_uv"][q] if out["prior_mask"][q].any() else None, crop=int(order[q]))
        self.detections[view_id] = det
        for q, o in enumerate(ids):
            self.diameters[o] = float(diam[q])
            if q < n1 or cam_ok:
                self.obj_num_dets[o] = self.obj_num_dets.get(o, 0) + 1           # one per crop that went through the network (:1153)
        if cam_ok:
            self.cam_poses[view_id] = out["T_GtoC"].copy()
            self.view_ids.append(view_id)
            for q, o in enumerate(ids):
                if out["map_valid"][q]:
                    self.obj_poses[o] = np.vstack([out["T_OtoG"][q], [0, 0, 0, 1.0]])
        inv = np.argsort(order)
        res = {k: (v[inv] if isinstance(v, np.ndarray) and v.shape[:1] == (L,) and k not in ("status",) else v) for k, v in out.items()}
        res["cam_ok"] = cam_ok
        res["reinit_ids"] = sorted(ids[q] for q in range(L) if out["reinit"][q])
        res["backup"] = backup
        res["culled"] = self._cull_objects() if cam_ok and out["status"][3] >= 3 else []       # optimize(curr_only=True) ran to its end
        res["global_stats"] = None
        if cam_ok and (self.sfm_mode or (self.global_opt_every and len(self.view_ids) > 1 and len(self.view_ids) % self.global_opt_every == 0)):      # :443-451
            res["global_stats"] = self.optimize_global()
        return res

    def _backup_camera_pose(self, obj_ids, bboxes, K):
        """ObjectSLAM.__backup_estimate_camera_pose (:933-973): PnP of the bbox centres against the map positions of the objects in view (one
        ``suo_pnp_batch`` object; the objects in the ORDER process_view received them); if that fails — fewer than four mapped objects, or
        PnP reports identity — the constant-velocity guess from the last two camera poses, or the last pose.  -> (T_GtoC [4,4], how)."""
        from . import geometry
        bboxes = np.asarray(bboxes, np.float32)
        cen = [0.5 * (bboxes[i, :2] + bboxes[i, 2:]) for i, o in enumerate(obj_ids) if o in self.obj_poses]
        ctr = [np.asarray(self.obj_poses[o], np.float64)[:3, 3] for o in obj_ids if o in self.obj_poses]
        if len(cen) >= 4:                                   # pnp() (:25-41): normalise with K, lambdatwist.pnp, identity = failure
            KinvT = np.linalg.inv(np.asarray(K, np.float64)).T
            p2n = np.stack(cen) @ KinvT[:2, :2] + KinvT[2:3, :2]
            T = geometry.pnp_batch([np.stack(ctr)], [p2n], seed=self.seed, obj_keys=[BACKUP_KEY], ctx=self.model.context())[0]
            if not np.allclose(T, np.eye(4)):
                return T, "pnp"
        if len(self.view_ids) > 1:
            T1, T2 = _as44(self.cam_poses[self.view_ids[-2]]), _as44(self.cam_poses[self.view_ids[-1]])
            return (T2 @ invert_SE3(T1)) @ T2, "const_vel"
        return _as44(self.cam_poses[self.view_ids[-1]]), "last"

    def _cull_objects(self):
        """The end of ObjectSLAM.optimize() (:913-930): objects whose detections hold too few inliers over all views leave the map."""
        removed = []
        for o in list(self.obj_poses):
            need = 3 if self.obj_num_dets.get(o, 0) < 3 else 6
            n = sum(int(np.count_nonzero(det[o]["inliers"])) for det in self.detections.values() if o in det)
            if n < need:
                self.obj_poses.pop(o)
                removed.append(o)
        return removed

    def optimize_global(self, its=(10, 10, 40, 40)):
        """ObjectSLAM.optimize(curr_only=False) in SLAM mode (:703-930): one vertex per mapped object (ordered as obj_poses) and per view with a
        camera pose (ordered as cam_poses, the first one fixed), one binary edge per gated keypoint of every detection of a mapped object;
        chi2 classification, four rounds with the Huber kernel stripped after the third, all inside one ``suo_ba_batch`` call whose graph
        couples cameras and objects (Schur-complement LM, csrc/ba_global.cu).  Then the objects behind the current camera (:899-911) and
        those with too few inliers (:913-930) are removed."""
        from . import ba
        if not self.view_ids:
            return None
        n_cam_e, n_obj_e = {}, {}
        for v, det in self.detections.items():
            if v in self.cam_poses:
                for o, d in det.items():
                    if o in self.obj_poses:
                        n = int(np.count_nonzero(d["inliers"]))
                        n_cam_e[v] = n_cam_e.get(v, 0) + n
                        n_obj_e[o] = n_obj_e.get(o, 0) + n
        overts = [o for o in self.obj_poses if n_obj_e.get(o, 0) > 0]
        cverts = [(i, v) for i, v in enumerate(self.cam_poses) if n_cam_e.get(v, 0) > 0]
        if not cverts or not overts:
            return None
        oi = {o: j for j, o in enumerate(overts)}
        ci = {v: len(overts) + j for j, (_, v) in enumerate(cverts)}
        poses = np.stack([np.asarray(self.obj_poses[o], np.float64)[:3] for o in overts] + [np.asarray(self.cam_poses[v], np.float64)[:3] for _, v in cverts])
        fixed = np.array([0] * len(overts) + [1 if i == 0 else 0 for i, _ in cverts], np.uint8)
        e_obj, e_cam, P_, K_, U_, I_, owner = [], [], [], [], [], [], []
        for v, det in self.detections.items():
            for o, d in det.items():
                if v in ci and o in oi:
                    n = len(d["uv_pred"])
                    S = np.asarray(d["cov_pred"], np.float64).reshape(n, 2, 2)
                    dt = S[:, 0, 0] * S[:, 1, 1] - S[:, 0, 1] * S[:, 1, 0]
                    I_.append(np.stack([S[:, 1, 1] / dt, -S[:, 0, 1] / dt, -S[:, 1, 0] / dt, S[:, 0, 0] / dt], 1))
                    K_.append(np.tile([d["K"][0, 0], d["K"][1, 1], d["K"][0, 2], d["K"][1, 2]], (n, 1)))
                    P_.append(np.asarray(d["model_kp"], np.float64)); U_.append(np.asarray(d["uv_pred"], np.float64))
                    e_obj += [oi[o]] * n; e_cam += [ci[v]] * n
                    owner += [(v, o, k) for k in range(n)]
        if not owner:
            return None
        P, inl, stats = ba.ba_batch([0, len(poses)], [0, len(owner)], poses, fixed, e_obj, e_cam, np.concatenate(K_), np.concatenate(P_), np.concatenate(U_),
                                    np.concatenate(I_), np.ones(len(owner), np.uint8), list(its), ctx=self.model.context())
        for (v, o, k), f in zip(owner, inl):
            self.detections[v][o]["inliers"][k] = f
        for _, v in cverts:
            self.cam_poses[v] = P[ci[v]].copy()
        cur, behind = self.view_ids[-1], []
        for o in overts:
            self.obj_poses[o] = np.vstack([P[oi[o]], [0, 0, 0, 1.0]])
            if cur in self.cam_poses:
                Tc = self.cam_poses[cur]
                if (Tc[:3, :3] @ self.obj_poses[o][:3, 3] + Tc[:3, 3])[2] < 0.5 * self.diameters[o]:
                    self.obj_poses.pop(o)
                    behind.append(o)
        return dict(rounds=int(stats[0, 0]), outer=int(stats[0, 1]), trials=int(stats[0, 2]), behind=behind, culled=self._cull_objects())
