"""TEST INFRASTRUCTURE ONLY — generates tests/golden/slam_seq.npz by running the UNMODIFIED reference class
``lib.object_slam.ObjectSLAM`` (process_view in SLAM mode: the two __process_objects passes, __estimate_camera_pose,
__maybe_reinit_objects, optimize(curr_only=True) with its rounds / chi2 gate / Huber strip — lib/object_slam.py:327-930,975-1072)
on synthetic marker sequences — and, in single_view_mode, on single marker frames (process_view + the full optimize() with the camera
fixed, the path BASELINE configs[1] times) — in THIS container (CPU; /root/reference is not on the GPU box, hence the committed fixture).

What is the reference's and what is substituted:
  * reference, unmodified: lib/object_slam.py (all control flow, gating, map bookkeeping, the optimize() loop), lib/models/pkpnet.py (the
    network, torch CPU), lib/utils/utils.py (fix_K_for_bbox_ndc, make_prior_kp_input, invert_SE3, ...), lib/labeling/kp_config.py;
  * the two native extension modules the reference imports and that cannot be built here (no Eigen / Ceres / SuiteSparse):
      - ``lambdatwist``  -> oracle/geom.py lambdatwist_pnp (RANSAC + refine restatement; its P3P/P4P core is checked against the reference's
                            own p4p.cpp compiled in place, tests/test_oracle_geom.py).  The RANSAC stream is keyed by the crop's position in
                            the non-symmetric-first processing order (what the device kernels and oracle/slam_frame_oracle.py use); the shim
                            recovers that position from the model keypoints it is handed.
      - ``g2o``          -> suo_slam_b200/g2o.py container classes (the product's drop-in: vertices, edges, SparseOptimizer) with the LM solve
                            routed to oracle/geom.py ba_optimize_err instead of the GPU — so the fixture also proves that the LITERAL reference
                            optimize() runs on the drop-in's class surface;
  * stubbed: thirdparty.bop_toolkit...renderer_py (needs glumpy; visualisation only), matplotlib (oracle/ref_shims.py).

The fixture pins oracle/slam_frame_oracle.py + oracle/slam_oracle.py (tests/test_marker_cpu.py) and, through them and directly, the GPU path
(tests/test_gpu_slam.py).

    python -m oracle.gen_golden_slam        # ~1-2 min on 8 cores
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import geom, ref_shims  # noqa: E402
from suo_slam_b200 import synth  # noqa: E402

SEED = 0                     # RANSAC seed (slam_frame_oracle.process_view's default)
BACKUP_KEY = 10 ** 6         # = oracle/slam_frame_oracle.py BACKUP_KEY


class _CpuBa:
    """Stands in for suo_slam_b200.ba inside the g2o drop-in: the same call, solved by the CPU oracle."""

    @staticmethod
    def ba_batch(prob_vert, prob_edge, poses, fixed, e_obj, e_cam, cam_k, p, uv, info, inliers, its, huber_delta, chi2_gate,
                 init_with_outliers, return_errors=True):
        assert list(prob_vert) == [0, len(poses)] and return_errors
        P, inl, st, err = geom.ba_optimize_err(poses, fixed, e_obj, e_cam, cam_k, p, uv, info, inliers, its, huber_delta=huber_delta,
                                               chi2_gate=chi2_gate, init_with_outliers=init_with_outliers)
        return P, inl, np.array([[st["rounds"], st["outer"], st["trials"]]], np.int32), err


class _PnpShim:
    """lambdatwist.pnp(xs, ys, threshold) -> 4x4 (identity on failure).  `crops` = [(model_kps [K,3], key)] of the view being processed."""

    def __init__(self):
        self.crops = []
        self.calls = []

    def pnp(self, xs_in, ys_in, threshold=0.001):
        xs_in = np.asarray(xs_in, np.float64)
        key = [k for mk, k in self.crops if (mk == xs_in[0]).all(-1).any()]
        assert len(key) <= 1, "cannot tell which crop this pnp() call belongs to"
        if not key:                                # not a crop's model keypoints: the bbox-centroid PnP of __backup_estimate_camera_pose (:953)
            key = [BACKUP_KEY]
        self.calls.append(key[0])
        T, _ = geom.lambdatwist_pnp(xs_in, np.asarray(ys_in, np.float64), threshold, seed=SEED, obj_key=key[0])
        return np.eye(4) if T is None else T


def install():
    ref_shims.install()
    for name in ("thirdparty", "thirdparty.bop_toolkit", "thirdparty.bop_toolkit.bop_toolkit_lib", "thirdparty.bop_toolkit.bop_toolkit_lib.renderer_py"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules["thirdparty.bop_toolkit.bop_toolkit_lib.renderer_py"].RendererPython = object
    shim = _PnpShim()
    lt = types.ModuleType("lambdatwist")
    lt.pnp = shim.pnp
    sys.modules["lambdatwist"] = lt
    import suo_slam_b200.g2o as g2o_mod
    g2o_mod._ba = _CpuBa
    sys.modules["g2o"] = g2o_mod
    torch.serialization.add_safe_globals([argparse.Namespace])
    import importlib
    return importlib.import_module("lib.object_slam"), shim


def run_sequence(osl_mod, shim, ckpt, seq, n_views, corrupt_after_first=None, view_objs=None, ext_cam=None, **kw):
    objs = seq["objs"]
    mesh_db = {o["obj_id"]: dict(is_symmetric=bool(o["is_symmetric"]), diameter=float(o["diameter"])) for o in objs}
    with contextlib.redirect_stdout(io.StringIO()):
        slam = osl_mod.ObjectSLAM(ckpt, mesh_db, **kw)
    # ObjectSLAM builds PkpNet() with its default input_res whatever pred_res says (lib/object_slam.py:95 vs :1080): give the network the
    # resolution the priors are made for — an attribute of the reference object, not a code change
    slam.model.input_res = tuple(kw.get("pred_res", (256, 256)))
    out = {}
    for i, v in enumerate(seq["views"][:n_views]):
        present = list(range(len(objs))) if view_objs is None else list(view_objs(i))          # indices of the objects detected in this view
        is_sym = np.array([objs[c]["is_symmetric"] or ext_cam is not None for c in present])     # (an external pose makes every object "symmetric", :353)
        order = [present[j] for j in np.concatenate([np.nonzero(~is_sym)[0], np.nonzero(is_sym)[0]])]
        shim.crops = [(objs[c]["model_kps"], pos) for pos, c in enumerate(order)]
        shim.calls = []
        obj_ids = np.array([v["dets"][c]["obj_id"] for c in present])
        bboxes = np.stack([v["dets"][c]["bbox"] for c in present]).astype(np.float32)
        mk, mm = np.stack([objs[c]["model_kps"] for c in present]), np.stack([objs[c]["model_kps_mask"] for c in present])
        with contextlib.redirect_stdout(io.StringIO()):
            slam.process_view(v["view_id"], v["img"], seq["K"], obj_ids, bboxes.copy(), mk, mm, mm.copy(),
                              cam_pose=None if ext_cam is None else ext_cam(i, v))
        vid = v["view_id"]
        assert vid in slam.cam_poses, "the reference lost the camera"
        out[f"v{i}_cam"] = np.asarray(slam.cam_poses[vid], np.float64)[:3]
        out[f"v{i}_pnp_keys"] = np.asarray(shim.calls, np.int32)
        out[f"v{i}_obj_ids"] = np.array(sorted(slam.obj_poses), np.int32)
        out[f"v{i}_obj_poses"] = np.stack([np.asarray(slam.obj_poses[o], np.float64)[:3] for o in sorted(slam.obj_poses)])
        for o, d in slam.detections[vid].items():
            out[f"v{i}_det{o}_kp_mask"] = d["kp_mask"].astype(np.uint8)
            out[f"v{i}_det{o}_inliers"] = np.asarray(d["inliers"]).astype(np.uint8)
            out[f"v{i}_det{o}_uv"] = d["uv_pred"].astype(np.float64)
            out[f"v{i}_det{o}_cov"] = d["cov_pred"].astype(np.float32)
            out[f"v{i}_det{o}_K"] = np.asarray(d["K"], np.float64)                 # utils.fix_K_for_bbox_ndc on the float32 bbox, through float32 (:1082-1086,1140)
            out[f"v{i}_det{o}_pose"] = np.zeros((0, 4)) if d["pose"] is None else np.asarray(d["pose"], np.float64)[:3]
            out[f"v{i}_det{o}_prior_uv"] = np.zeros((0, 2), np.float32) if d["prior_uv"] is None else d["prior_uv"]
        if corrupt_after_first is not None and i == 0:
            T = np.array(slam.obj_poses[corrupt_after_first], np.float64)
            T[:3, 3] += [70.0, -50.0, 40.0]
            slam.obj_poses[corrupt_after_first] = T
    return out


def run_single_view_frames(osl_mod, shim, ckpt, seed0, n_frames, n_obj=8):
    """single_view_mode (what evaluate.py runs for the single-view tables and what BASELINE configs[1] times): every frame on its own —
    process_view = one __process_objects pass over all crops + the full optimize() with the camera vertex fixed and its = [10] * 4
    (lib/object_slam.py:362-364,443-451,842-846).  RANSAC keys = the crop's index in the batch of frames, as suo_frames / frame_oracle use."""
    out = {}
    for f in range(n_frames):
        fr = synth.make_marker_frame(seed0 + f, n_obj=n_obj)
        objs = fr["objs"]
        ids = np.arange(n_obj) + 100 * (f + 1)
        mesh_db = {int(i): dict(is_symmetric=False, diameter=float(o["diameter"])) for i, o in zip(ids, objs)}
        with contextlib.redirect_stdout(io.StringIO()):
            slam = osl_mod.ObjectSLAM(ckpt, mesh_db, single_view_mode=True)
        shim.crops = [(o["model_kps"], n_obj * f + k) for k, o in enumerate(objs)]
        shim.calls = []
        mk, mm = np.stack([o["model_kps"] for o in objs]), np.stack([o["model_kps_mask"] for o in objs])
        bboxes = np.stack([o["bbox"] for o in objs]).astype(np.float32)
        with contextlib.redirect_stdout(io.StringIO()):
            slam.process_view(f, fr["img"], fr["K"], ids, bboxes.copy(), mk, mm, mm.copy())
        out[f"f{f}_pnp_keys"] = np.asarray(shim.calls, np.int32)
        out[f"f{f}_kept"] = np.array([int(i) in slam.obj_poses for i in ids], np.uint8)           # (objects optimize() did not remove, :899-930)
        out[f"f{f}_T_ba"] = np.stack([np.asarray(slam.obj_poses[int(i)], np.float64)[:3] if int(i) in slam.obj_poses else np.zeros((3, 4)) for i in ids])
        det = slam.detections[f]
        out[f"f{f}_kp_used"] = np.stack([det[int(i)]["kp_mask"] for i in ids]).astype(np.uint8)
        out[f"f{f}_accepted"] = np.array([det[int(i)]["pose"] is not None for i in ids], np.uint8)
        out[f"f{f}_T_pnp"] = np.stack([np.asarray(det[int(i)]["pose"], np.float64)[:3] if det[int(i)]["pose"] is not None else np.zeros((3, 4)) for i in ids])
        K = mm.shape[1]
        inl = np.zeros((n_obj, K), np.uint8)
        for k, i in enumerate(ids):
            inl[k, det[int(i)]["kp_mask"]] = np.asarray(det[int(i)]["inliers"]).astype(np.uint8)
        out[f"f{f}_ba_inliers"] = inl
    return out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "slam_seq.npz"))
    ap.add_argument("--scenarios", default="clean,corrupt,glob,allsym,newnon,cv,sfm,extcam,c5,sv")
    a = ap.parse_args(argv)
    want = a.scenarios.split(",")
    torch.manual_seed(0)
    osl_mod, shim = install()
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    ckpt = os.path.join(ROOT, "build", f"marker_ckpt_{os.getpid()}.pth.tar")
    torch.save({"model": synth.make_marker_state_dict(0), "epoch": 0, "args": argparse.Namespace(synthetic=True)}, ckpt)
    fix = {}
    try:
        seq = synth.make_slam_sequence(3, n_views=4, n_obj=6)
        if "clean" in want:
            # (a) 4 views, 6 objects (3 symmetric), YCBV thresholds (evaluate.py:58-66): the scenario of tests/test_gpu_slam.py
            for k, v in run_sequence(osl_mod, shim, ckpt, seq, 4).items():
                fix["clean_" + k] = v
        if "corrupt" in want:
            # (b) the same sequence with object 13's map pose pushed away after the first view: the vote must reject it, the
            #     re-initialisation test must replace it from the PnP result (lib/object_slam.py:595-697)
            for k, v in run_sequence(osl_mod, shim, ckpt, seq, 3, corrupt_after_first=13).items():
                fix["corrupt_" + k] = v
        if "glob" in want:
            # (b2) the clean sequence with the periodic GLOBAL optimize() (cameras and objects free, LinearSolverCholmod, its = [10, 10, 40, 40],
            #      lib/object_slam.py:443-451,736-778) after views 2 and 4
            for k, v in run_sequence(osl_mod, shim, ckpt, seq, 4, global_opt_every=2).items():
                fix["glob_" + k] = v
        # (b3-b5) __backup_estimate_camera_pose (:933-973)
        if "allsym" in want:       # every object symmetric: bbox-centroid PnP BEFORE the passes (:372-391), every crop gets a prior
            sq = synth.make_slam_sequence(5, n_views=3, n_obj=6, n_sym=6)
            for k, v in run_sequence(osl_mod, shim, ckpt, sq, 3).items():
                fix["allsym_" + k] = v
        if "newnon" in want:       # view 0 sees only the 4 symmetric objects; the 4 non-symmetric ones appear in view 1 and are not in the map:
            sq = synth.make_slam_sequence(6, n_views=3, n_obj=8, n_sym=4)      # the vote has no hypothesis -> centroid PnP AFTER the first pass (:404-411)
            for k, v in run_sequence(osl_mod, shim, ckpt, sq, 3, view_objs=lambda i: range(4) if i == 0 else range(8)).items():
                fix["newnon_" + k] = v
        if "cv" in want:           # three objects: the centroid PnP has < 4 points -> last pose (view 1), constant-velocity guess (view 2)
            sq = synth.make_slam_sequence(7, n_views=3, n_obj=3, n_sym=3)
            for k, v in run_sequence(osl_mod, shim, ckpt, sq, 3).items():
                fix["cv_" + k] = v
        if "sfm" in want:          # sfm_mode: re-initialisation test over all views, its = [10, 10, 40, 40] for the per-view solve too, global optimize() after EVERY view
            for k, v in run_sequence(osl_mod, shim, ckpt, seq, 3, sfm_mode=True).items():
                fix["sfm_" + k] = v
        if "extcam" in want:       # external camera poses (process_view's cam_pose, :349-353): the ground truth perturbed by a few mm / mrad
            def noisy(i, v):
                rng = np.random.default_rng(900 + i)
                w = rng.normal(scale=2e-3, size=3)
                Wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
                T = np.array(v["T_GtoC"][:3], np.float64)
                T[:, :3] = (np.eye(3) + Wx + 0.5 * Wx @ Wx) @ T[:, :3]
                T[:, 3] += rng.normal(scale=3.0, size=3)
                return T
            for k, v in run_sequence(osl_mod, shim, ckpt, seq, 3, ext_cam=noisy).items():
                fix["extcam_" + k] = v
        if "c5" in want:
            # (c) configs[4] shape: 512x512 crops, T-LESS thresholds (evaluate.py:68-76), 4 objects of which 2 symmetric, 2 views
            seq5 = synth.make_slam_sequence(11, n_views=2, n_obj=4, res=512, n_sym=2, radius=2 * synth.MARKER_RADIUS)
            tl = dict(pred_res=(512, 512), kp_var_thresh=0.5, bbox_thresh=1.0, manual_kp_std=0.1, opt_init_with_outliers=True)
            for k, v in run_sequence(osl_mod, shim, ckpt, seq5, 2, **tl).items():
                fix["c5_" + k] = v
        if "sv" in want:
            # (d) single-view mode: the two frames of tests/test_gpu_round2.py::test_pixels_to_poses (8 crops each)
            for k, v in run_single_view_frames(osl_mod, shim, ckpt, 2000, 2).items():
                fix["sv_" + k] = v
    finally:
        os.remove(ckpt)
    np.savez_compressed(a.out, **fix)
    print("wrote", a.out, os.path.getsize(a.out), "bytes,", len(fix), "arrays")
    if "sv" in want:
        for f in range(2):
            print("single view frame", f, "accepted", fix[f"sv_f{f}_accepted"].tolist(), "kept", fix[f"sv_f{f}_kept"].tolist(), "gated", fix[f"sv_f{f}_kp_used"].sum(1).tolist(),
                  "BA inliers", fix[f"sv_f{f}_ba_inliers"].sum(1).tolist())
    for s in [w for w in want if w != "sv"]:
        n = len([k for k in fix if k.startswith(s + "_v") and k.endswith("_cam")])
        print(s, "views", n, "objects in the map at the end", fix[f"{s}_v{n - 1}_obj_ids"].tolist(), "pnp keys of the last view", fix[f"{s}_v{n - 1}_pnp_keys"].tolist())


if __name__ == "__main__":
    main()
