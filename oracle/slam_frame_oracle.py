"""TEST INFRASTRUCTURE ONLY — the reference's SLAM-mode frame (ObjectSLAM.process_view with single_view_mode = False, reference
lib/object_slam.py:327-451) restated on the CPU over the other oracles: network forward (net_oracle), prior planes
(prior_oracle), PnP and the curr_only LM (geom), camera-pose vote and re-initialisation test (slam_oracle).  Only tests/, smoke()
and bench.py's CPU legs may import this; the product path never does.

Pinned: tests/golden/slam_seq.npz holds the state the UNMODIFIED reference class reaches (lib/object_slam.py imported from /root/reference by
oracle/gen_golden_slam.py: reference control flow, reference network, reference utils; its two native extension modules replaced by the oracle's
PnP / LM) on three marker sequences — 4 views clean, 3 views with a corrupted map pose (vote rejection + re-initialisation), 2 views at 512x512
with the T-LESS thresholds; tests/test_marker_cpu.py replays them here: gating, chi2 classifications, re-init decisions identical, keypoints 1e-5,
poses 1e-6 of the scene scale.  A fourth sequence runs with
global_opt_every = 2, i.e. the periodic full optimize() (:443-451, :736-778: cameras AND objects free, its = [10, 10, 40, 40]) after views 2 and 4.
Three more exercise __backup_estimate_camera_pose (:933-973): every object symmetric (bbox-centroid PnP before the passes), the non-symmetric
objects new to the map (the vote has no hypothesis: centroid PnP after the first pass), and three objects only (centroid PnP impossible:
last pose, then the constant-velocity guess)."""
from __future__ import annotations

import numpy as np
import torch

from . import geom, net_oracle, prior_oracle, slam_oracle


def fix_K_for_bbox_ndc(K, bbox, f32=True):
    """utils.fix_K_for_bbox_ndc (lib/utils/utils.py:416-429).  f32: rounded through float32 as __run_kp_model stores it
    (K_bbox_np float32, :1082,1086) and widened again (:1140); the prior projection (:500) uses the FP64 product as it is."""
    b = np.asarray(bbox)
    if b.dtype == np.float32:        # the reference's float32 scalar arithmetic on a float32 bbox (NumPy >= 2 semantics, see synth.fix_K_for_bbox_ndc)
        sx, sy = float(np.float32(2.0) / np.float32(b[2] - b[0])), float(np.float32(-2.0) / np.float32(b[3] - b[1]))
    else:
        sx, sy = 2.0 / (float(b[2]) - float(b[0])), -2.0 / (float(b[3]) - float(b[1]))
    x1, y1 = float(b[0]), float(b[1])
    T = np.eye(3); T[0, 2], T[1, 2] = -x1, -y1
    S = np.eye(3); S[0, :] *= sx; S[1, :] *= sy; S[0, 2] -= 1.0; S[1, 2] += 1.0
    Kb = S @ T @ np.asarray(K, np.float64)
    return Kb.astype(np.float32).astype(np.float64) if f32 else Kb


class State:
    """ObjectSLAM's map (lib/object_slam.py:100-123): obj_poses {obj: T_OtoG [4,4]}, cam_poses {view: T_GtoC [3,4]},
    detections {view: {obj: det}}, view_ids."""

    def __init__(self):
        self.obj_poses, self.cam_poses, self.detections, self.view_ids = {}, {}, {}, []
        self.num_dets, self.diam = {}, {}          # obj_num_dets (:149,1153) and mesh_db[obj]["diameter"]


def _to44(T):
    T = np.asarray(T, np.float64)
    return T if T.shape[0] == 4 else np.vstack([T, [0, 0, 0, 1.0]])


@torch.no_grad()
def _run_kp_model(sd, img_u8, K, obj_keys, bboxes, model_kps, model_masks, diameters, priors, res, kvt, bt, seed):
    """__run_kp_model (:1077-1167): forward on this group's crops, gating, per-object pnp()."""
    img = torch.from_numpy(img_u8.transpose(2, 0, 1).astype(np.float32) / 255)[None]
    out = net_oracle.pkpnet_forward(sd, img, [torch.as_tensor(np.asarray(bboxes, np.float32))],
                                    None if priors is None else [torch.from_numpy(priors)], (res, res))
    uv, cov, km = out["uv"].numpy(), out["cov"].numpy(), out["kp_mask"].numpy()
    masks = net_oracle.gate_keypoints(uv, cov, km, np.asarray(model_masks, bool), bt, kvt)
    dets = []
    for k in range(len(bboxes)):
        m = masks[k]
        Kb = fix_K_for_bbox_ndc(K, bboxes[k])
        kp_model, uv_pred, cov_pred = model_kps[k][m].astype(np.float64), uv[k][m].astype(np.float64), cov[k][m]
        pose = None
        r = geom.pnp(kp_model, uv_pred, Kb, seed=seed, obj_key=int(obj_keys[k]))
        if r is not None and r[0][2, 3] > 0.5 * diameters[k] and m.sum() >= 4:
            pose = _to44(r[0])
        dets.append(dict(pose=pose, inliers=np.ones(int(m.sum()), bool), kp_mask=m, model_kp=kp_model, uv_pred=uv_pred, cov_pred=cov_pred,
                         K=Kb, uv_full=uv[k], cov_full=cov[k], bbox=np.asarray(bboxes[k])))
    return dets


def _process_objects(st, sd, is_sym, view_id, img, K, idx, keys, obj_ids, bboxes, model_kps, model_masks, diameters, res, kvt, bt, seed, first):
    """__process_objects (:470-593) for the crops `idx` of the frame; keys[c] = RANSAC object key of crop c (its position in the
    non-symmetric-first order, the order the crops are processed in)."""
    if len(idx) == 0:
        return {}
    priors, prior_uv = None, {}
    if is_sym and view_id in st.cam_poses:                      # :486-520
        priors = np.zeros((len(idx), model_masks.shape[1], res, res), np.float32)
        T_GtoC = _to44(st.cam_poses[view_id])
        for q, c in enumerate(idx):
            o = obj_ids[c]
            if o not in st.obj_poses:
                continue
            m = model_masks[c].astype(bool)
            T_OtoC = T_GtoC @ _to44(st.obj_poses[o])
            pc = model_kps[c][m] @ T_OtoC[:3, :3].T + T_OtoC[:3, 3]
            uvd = pc @ fix_K_for_bbox_ndc(K, bboxes[c], f32=False).T
            if np.all(uvd[:, 2] > 0):
                full = np.zeros((len(m), 2), np.float32)
                full[m] = uvd[:, :2] / uvd[:, 2:3]
                prior_uv[o] = full
                priors[q] = prior_oracle.make_prior_kp_input(full, m, (res, res), ndc=True)
    dets = _run_kp_model(sd, img, K, keys[idx], bboxes[idx], model_kps[idx], model_masks[idx], diameters[idx], priors, res, kvt, bt, seed)
    for c in idx:
        st.num_dets[obj_ids[c]] = st.num_dets.get(obj_ids[c], 0) + 1          # :1153
        st.diam[obj_ids[c]] = float(diameters[c])
    detection = {}
    for q, c in enumerate(idx):
        o = obj_ids[c]
        detection[o] = dict(dets[q], prior_uv=prior_uv.get(o), crop=int(c))
        if first and dets[q]["pose"] is not None:               # :541-556: the first view defines the world frame
            st.obj_poses[o] = dets[q]["pose"] if view_id not in st.cam_poses else slam_oracle._inv_se3(_to44(st.cam_poses[view_id])) @ dets[q]["pose"]
    st.detections.setdefault(view_id, {}).update(detection)
    if view_id not in st.cam_poses:                              # :566-575
        if first:
            st.cam_poses[view_id] = np.eye(4)[:3]
        else:
            cam, counts, _ = slam_oracle.estimate_camera_pose(st.obj_poses, st.detections[view_id])
            if cam is None:
                return detection
            st.cam_poses[view_id] = cam[:3]
        st.view_ids.append(view_id)
    for c in idx:                                                # :577-592: objects that could not be initialised before
        o = obj_ids[c]
        if o not in st.obj_poses and detection[o]["pose"] is not None:
            st.obj_poses[o] = slam_oracle._inv_se3(_to44(st.cam_poses[view_id])) @ detection[o]["pose"]
    return detection


def _optimize_curr_only(st, view_id, init_with_outliers, its=(10, 10, 10, 10)):
    """optimize(curr_only=True) (:703-930): one free camera vertex, one EdgeSE3ProjectFromFixedObject per gated keypoint of every
    mapped object of the current view, its = [10] * 4."""
    if view_id not in st.cam_poses:
        return None
    dets = [(o, d) for o, d in st.detections[view_id].items() if o in st.obj_poses]
    if sum(int(np.count_nonzero(d["inliers"])) for _, d in dets) < 3:      # :726-728
        return None
    p, cam_k, uv, info, owner = [], [], [], [], []
    for o, d in dets:
        T = _to44(st.obj_poses[o])
        for k in range(len(d["uv_pred"])):
            p.append(T[:3, :3] @ d["model_kp"][k] + T[:3, 3])
            cam_k.append([d["K"][0, 0], d["K"][1, 1], d["K"][0, 2], d["K"][1, 2]])
            uv.append(d["uv_pred"][k])
            S = d["cov_pred"][k].astype(np.float64)
            det = S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]
            info.append([S[1, 1] / det, -S[0, 1] / det, -S[1, 0] / det, S[0, 0] / det])
            owner.append((o, k))
    n = len(p)
    if n == 0:
        return None
    P, inl, stats = geom.ba_optimize(np.asarray(st.cam_poses[view_id])[None, :3], np.zeros(1, np.uint8), np.full(n, -1, np.int32), np.zeros(n, np.int32),
                                     np.asarray(cam_k), np.asarray(p), np.asarray(uv), np.asarray(info), np.ones(n), list(its),
                                     init_with_outliers=init_with_outliers)
    st.cam_poses[view_id] = P[0]
    for (o, k), v in zip(owner, inl):
        st.detections[view_id][o]["inliers"][k] = v
    return dict(stats, culled=_cull_objects(st))          # (optimize() ends with the inlier-count check in either mode, :913-930)


BACKUP_KEY = 10 ** 6        # RANSAC stream of the bbox-centroid PnP (the per-crop streams use the crop's position 0..L-1)


def _backup_estimate_camera_pose(st, view_id, obj_ids, bboxes, K, seed):
    """__backup_estimate_camera_pose (:933-973): PnP of the bbox centres against the map positions of the objects in view; if that fails
    (fewer than 4 mapped objects, or PnP returns identity) a constant-velocity guess from the last two camera poses, or the last pose.
    Always produces a pose and registers the view."""
    assert st.view_ids and view_id not in st.view_ids and view_id not in st.cam_poses
    cen = [0.5 * (bboxes[i, :2] + bboxes[i, 2:]) for i, o in enumerate(obj_ids) if o in st.obj_poses]
    ctr = [_to44(st.obj_poses[o])[:3, 3] for o in obj_ids if o in st.obj_poses]
    r = geom.pnp(np.stack(ctr), np.stack(cen), np.asarray(K, np.float64), seed=seed, obj_key=BACKUP_KEY) if cen else None
    if r is not None:
        st.cam_poses[view_id], how = r[0], "pnp"
    elif len(st.view_ids) > 1:
        T1, T2 = _to44(st.cam_poses[st.view_ids[-2]]), _to44(st.cam_poses[st.view_ids[-1]])
        st.cam_poses[view_id], how = (T2 @ slam_oracle._inv_se3(T1)) @ T2, "const_vel"
    else:
        st.cam_poses[view_id], how = st.cam_poses[st.view_ids[-1]], "last"
    st.view_ids.append(view_id)
    return how


def _edge_arrays(T_obj, d):
    """cam_k / uv / information of the edges of one detection (:805-832); p in the frame of T_obj (None: the object's own frame)."""
    p, cam_k, uv, info = [], [], [], []
    for k in range(len(d["uv_pred"])):
        p.append(d["model_kp"][k] if T_obj is None else T_obj[:3, :3] @ d["model_kp"][k] + T_obj[:3, 3])
        cam_k.append([d["K"][0, 0], d["K"][1, 1], d["K"][0, 2], d["K"][1, 2]])
        uv.append(d["uv_pred"][k])
        S = d["cov_pred"][k].astype(np.float64)
        det = S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]
        info.append([S[1, 1] / det, -S[0, 1] / det, -S[1, 0] / det, S[0, 0] / det])
    return p, cam_k, uv, info


def _cull_objects(st):
    """The end of optimize() (:913-930): objects whose detections hold too few inliers over all views leave the map."""
    removed = []
    for o in list(st.obj_poses):
        need = 3 if st.num_dets.get(o, 0) < 3 else 6
        n = sum(int(np.count_nonzero(det[o]["inliers"])) for det in st.detections.values() if o in det)
        if n < need:
            st.obj_poses.pop(o)
            removed.append(o)
    return removed


def _optimize_global(st, its=(10, 10, 40, 40)):
    """optimize(curr_only=False) in SLAM mode (:703-930): one vertex per mapped object (id = its position in obj_poses) and per view with a
    camera pose (id = position in cam_poses + len(obj_poses), the first one fixed), one EdgeSE3ProjectFromObject per gated keypoint of every
    detection of a mapped object, chi2 classification of all edges, its = [10, 10, 40, 40] with the Huber kernel stripped after round 2,
    then the objects that ended up behind the current camera (:899-911) or with too few inliers (:913-930) are removed."""
    if not st.view_ids:
        return None
    obj_ids = list(st.obj_poses)
    n_cam_e, n_obj_e = {}, {}
    for v, det in st.detections.items():
        if v in st.cam_poses:
            for o, d in det.items():
                if o in st.obj_poses:
                    n = int(np.count_nonzero(d["inliers"]))
                    n_cam_e[v] = n_cam_e.get(v, 0) + n
                    n_obj_e[o] = n_obj_e.get(o, 0) + n
    overts = [o for o in obj_ids if n_obj_e.get(o, 0) > 0]
    cverts = [(i, v) for i, v in enumerate(st.cam_poses) if n_cam_e.get(v, 0) > 0]
    if not cverts or not overts:
        return None
    oi = {o: j for j, o in enumerate(overts)}
    ci = {v: len(overts) + j for j, (_, v) in enumerate(cverts)}
    poses = np.stack([_to44(st.obj_poses[o])[:3] for o in overts] + [np.asarray(st.cam_poses[v], np.float64)[:3] for _, v in cverts])
    fixed = np.array([0] * len(overts) + [1 if i == 0 else 0 for i, _ in cverts], np.uint8)
    e_obj, e_cam, P_, K_, U_, I_, owner = [], [], [], [], [], [], []
    for v, det in st.detections.items():
        for o, d in det.items():
            if v in ci and o in oi:
                p, cam_k, uv, info = _edge_arrays(None, d)
                e_obj += [oi[o]] * len(p); e_cam += [ci[v]] * len(p)
                P_ += p; K_ += cam_k; U_ += uv; I_ += info
                owner += [(v, o, k) for k in range(len(p))]
    n = len(P_)
    if n == 0:
        return None
    P, inl, stats = geom.ba_optimize(poses, fixed, np.asarray(e_obj, np.int32), np.asarray(e_cam, np.int32), np.asarray(K_), np.asarray(P_), np.asarray(U_),
                                     np.asarray(I_), np.ones(n), list(its))
    for (v, o, k), f in zip(owner, inl):
        st.detections[v][o]["inliers"][k] = f
    for _, v in cverts:
        st.cam_poses[v] = P[ci[v]]
    cur = st.view_ids[-1]
    behind = []
    for o in overts:
        st.obj_poses[o] = P[oi[o]]
        if cur in st.cam_poses:
            Tc = np.asarray(st.cam_poses[cur])
            if (Tc[:3, :3] @ st.obj_poses[o][:3, 3] + Tc[:3, 3])[2] < 0.5 * st.diam[o]:
                st.obj_poses.pop(o)
                behind.append(o)
    return dict(stats, behind=behind, culled=_cull_objects(st))


def process_view(st: State, sd, view_id, img_u8, K, obj_ids, bboxes, model_kps, model_masks, is_sym, diameters, res=256,
                 kp_var_thresh=0.2, bbox_thresh=0.9, manual_kp_std=0.005, init_with_outliers=False, seed=0, global_opt_every=None, cam_pose=None,
                 sfm_mode=False):
    """process_view (:327-451), SLAM mode, no external camera pose, bbox_inflate = 0.  Symmetric crops get the prior heat maps.
    global_opt_every: the periodic full optimize() of :443-451 (None: never, as on sequences shorter than ObjectSLAM's default of 10)."""
    obj_ids, bboxes = list(obj_ids), np.asarray(bboxes, np.float32)
    model_kps, model_masks, is_sym, diameters = np.asarray(model_kps), np.asarray(model_masks).astype(bool), np.asarray(is_sym, bool), np.asarray(diameters, float)
    first = len(st.view_ids) == 0
    if cam_pose is not None:                                     # :349-353: external camera pose, every object gets the prior treatment
        st.cam_poses[view_id] = np.asarray(cam_pose, np.float64)
        st.view_ids.append(view_id)
        is_sym = np.ones(len(obj_ids), bool)
    non, sym = np.nonzero(~is_sym)[0], np.nonzero(is_sym)[0]
    keys = np.zeros(len(obj_ids), int)
    keys[np.concatenate([non, sym]).astype(int)] = np.arange(len(obj_ids))
    args = (keys, obj_ids, bboxes, model_kps, model_masks, diameters, res, kp_var_thresh, bbox_thresh, seed, first)
    backup = None
    if cam_pose is None and not first and len(non) == 0:         # :372-391: no non-symmetric object to vote with
        backup = _backup_estimate_camera_pose(st, view_id, obj_ids, bboxes, K, seed)
    _process_objects(st, sd, False, view_id, img_u8, K, non, *args)
    if view_id not in st.cam_poses:                              # :404-411
        if first:
            st.view_ids.append(view_id)
            st.cam_poses[view_id] = np.eye(4)[:3]
        else:                                                    # the vote failed (no mapped non-symmetric object with a PnP pose / < 4 inliers)
            backup = _backup_estimate_camera_pose(st, view_id, obj_ids, bboxes, K, seed)
    if len(sym):
        _process_objects(st, sd, True, view_id, img_u8, K, sym, *args)
    reinit, counts, _ = slam_oracle.maybe_reinit_objects(st.obj_poses, st.cam_poses, st.detections, st.view_ids, view_id,
                                                          len(st.view_ids) if sfm_mode else 15, manual_kp_std)          # :417
    for o, T in reinit.items():                                  # :683-690
        st.obj_poses[o] = T
    stats = _optimize_curr_only(st, view_id, init_with_outliers, (10, 10, 40, 40) if sfm_mode else (10, 10, 10, 10))       # :843-846
    glob = None
    if sfm_mode or (global_opt_every and len(st.view_ids) > 1 and len(st.view_ids) % global_opt_every == 0):      # :443-451
        glob = _optimize_global(st)
    return dict(cam_ok=True, reinit=sorted(reinit), reinit_counts=counts, ba_stats=stats, global_stats=glob, backup=backup)
