"""GPU: a SLAM-mode sequence through suo_slam_frame (one device-resident call per view: two dependent forwards, PnP, camera-pose vote,
device-rendered priors for the symmetric objects, object initialisation, re-initialisation test, curr_only LM) against the CPU
restatement of ObjectSLAM.process_view (oracle/slam_frame_oracle.py) on the same marker frames."""
import os

import numpy as np
import pytest

from oracle import slam_frame_oracle as sfo
from suo_slam_b200 import _lib, slam, synth
from suo_slam_b200.pkpnet import PkpNet

pytestmark = pytest.mark.gpu


def _check_vs_reference_fixture(G, name, i, trk, vid):
    """The tracker's state after view i against what the UNMODIFIED reference ObjectSLAM.process_view reached on the same sequence
    (tests/golden/slam_seq.npz, oracle/gen_golden_slam.py: reference control flow + reference network on the CPU, leaf solvers = the oracle)."""
    if f"{name}_v{i}_cam" not in G:
        return
    assert _rel(trk.cam_poses[vid], G[f"{name}_v{i}_cam"]) < 1e-3
    ids = G[f"{name}_v{i}_obj_ids"].tolist()
    assert sorted(trk.obj_poses) == ids
    for j, o in enumerate(ids):
        assert _rel(trk.obj_poses[o], G[f"{name}_v{i}_obj_poses"][j]) < 1e-3, o
    flips = 0
    for o, g in trk.detections[vid].items():
        # K_bbox from the device (slam_kbbox_kernel): the reference's float32 scalar arithmetic on the float32 bbox, bit for bit
        assert np.array_equal(g["K"], G[f"{name}_v{i}_det{o}_K"]), o
        gm = G[f"{name}_v{i}_det{o}_kp_mask"].astype(bool)
        if np.array_equal(g["kp_mask"], gm):                  # (a gate within the conv tolerance of its threshold is checked against the oracle)
            np.testing.assert_allclose(g["uv_pred"], G[f"{name}_v{i}_det{o}_uv"], atol=2e-4)
            flips += int((np.asarray(g["inliers"]).astype(bool) != G[f"{name}_v{i}_det{o}_inliers"].astype(bool)).sum())
            assert (g["pose"] is None) == (G[f"{name}_v{i}_det{o}_pose"].shape[0] == 0)
    print(f"[slam vs reference fixture {name} view {i}] cam rel diff {_rel(trk.cam_poses[vid], G[f'{name}_v{i}_cam']):.2e}, chi2 classifications that differ: {flips}")
    assert flips <= 1          # the keypoints differ by the conv tolerance (2e-4 NDC): an edge whose chi2 sits on the 5.991 gate may fall either way


def _view_args(seq, v):
    objs = seq["objs"]
    return (v["view_id"], v["img"], seq["K"], [d["obj_id"] for d in v["dets"]], np.stack([d["bbox"] for d in v["dets"]]),
            np.stack([o["model_kps"] for o in objs]), np.stack([o["model_kps_mask"] for o in objs]),
            np.array([o["is_symmetric"] for o in objs]), np.array([o["diameter"] for o in objs]))


@pytest.fixture(scope="module")
def marker_model():
    m = PkpNet(input_res=(256, 256), max_crops=16)
    m.load_state_dict(synth.make_marker_state_dict(0))
    m.cuda().eval()
    return m


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a)[:3] - np.asarray(b)[:3]) / np.linalg.norm(np.asarray(b)[:3]))


@pytest.mark.parametrize("corrupt", [False, True])
def test_slam_sequence_vs_the_cpu_oracle_and_the_reference_fixture(marker_model, golden_dir, corrupt):
    """4 views, 6 objects (3 symmetric).  corrupt: after the first view one NON-symmetric object's map pose is pushed away on both sides —
    its camera-pose vote must lose, and the re-initialisation test must replace the pose from the PnP result."""
    sd = synth.make_marker_state_dict(0)
    seq = synth.make_slam_sequence(3, n_views=4, n_obj=6)
    trk = slam.SlamTracker(marker_model)
    st = sfo.State()
    G = np.load(os.path.join(golden_dir, "slam_seq.npz"))
    bad = 13
    for i, v in enumerate(seq["views"]):
        a = _view_args(seq, v)
        out = trk.process_view(*a)
        ref = sfo.process_view(st, sd, *a)
        vid = v["view_id"]
        assert out["cam_ok"] == ref["cam_ok"] and out["cam_ok"]
        # network outputs of both forwards (the symmetric crops saw device-rendered priors) and the gating
        for o, d in st.detections[vid].items():
            g = trk.detections[vid][o]
            near = False
            if not np.array_equal(g["kp_mask"], d["kp_mask"]):       # a gate within the conv tolerance of its threshold
                std = np.sqrt(np.abs(d["cov_full"][:, [0, 1], [0, 1]]))
                near = bool(((np.abs(np.abs(d["uv_full"]).max(-1) - 0.9) < 1e-3) | (np.abs(std - 0.4).min(-1) < 1e-3))[g["kp_mask"] != d["kp_mask"]].all())
                assert near, o
            if not near:
                np.testing.assert_allclose(g["uv_pred"], d["uv_pred"], atol=2e-4)
                assert (g["pose"] is None) == (d["pose"] is None)
                if d["prior_uv"] is not None:
                    np.testing.assert_allclose(g["prior_uv"], d["prior_uv"], atol=1e-4)
                else:
                    assert g["prior_uv"] is None
        # camera pose after the vote + curr_only LM, the map, the re-initialisation decisions
        print(f"[slam view {vid} corrupt={corrupt}] cam rel diff {_rel(trk.cam_poses[vid], st.cam_poses[vid]):.2e}, vs ground truth "
              f"{np.linalg.norm(trk.cam_poses[vid][:, 3] - v['T_GtoC'][:3, 3]):.2f} mm, status {out['status'][:6].tolist()}, reinit {out['reinit_ids']}")
        assert _rel(trk.cam_poses[vid], st.cam_poses[vid]) < 1e-3
        assert np.linalg.norm(trk.cam_poses[vid][:, 3] - v["T_GtoC"][:3, 3]) < 12.0
        assert out["reinit_ids"] == ref["reinit"]
        if i >= 1:
            for o, n in ref["reinit_counts"].items():
                q = [d["obj_id"] for d in v["dets"]].index(o)
                got = out["reinit_counts"][q]
                assert abs(int(got[0]) - n["pnp"]) <= 1 and abs(int(got[1]) - n["estim"]) <= 1, (o, got, n)
        assert set(trk.obj_poses) == set(st.obj_poses)
        for o in st.obj_poses:
            assert _rel(trk.obj_poses[o], st.obj_poses[o]) < 1e-3, o
        _check_vs_reference_fixture(G, "corrupt" if corrupt else "clean", i, trk, vid)
        if corrupt and i == 0:
            for m in (trk.obj_poses, st.obj_poses):
                m[bad] = m[bad].copy()
                m[bad][:3, 3] += [70.0, -50.0, 40.0]
        if corrupt and i == 1:
            assert bad in out["reinit_ids"]


def test_slam_sequence_with_the_periodic_global_optimisation(marker_model, golden_dir):
    """global_opt_every = 2: after views 2 and 4 the tracker runs the full graph (cameras and objects free, its = [10, 10, 40, 40]) as one
    coupled suo_ba_batch problem (csrc/ba_global.cu) — against the oracle's restatement AND against the state the unmodified reference class
    reaches with global_opt_every = 2 (fixture "glob_*": its optimize() over LinearSolverCholmod moved the cameras by 1-4 mm)."""
    sd = synth.make_marker_state_dict(0)
    seq = synth.make_slam_sequence(3, n_views=4, n_obj=6)
    G = np.load(os.path.join(golden_dir, "slam_seq.npz"))
    trk = slam.SlamTracker(marker_model, global_opt_every=2)
    st = sfo.State()
    for i, v in enumerate(seq["views"]):
        a = _view_args(seq, v)
        out = trk.process_view(*a)
        ref = sfo.process_view(st, sd, *a, global_opt_every=2)
        vid = v["view_id"]
        assert (out["global_stats"] is not None) == (ref["global_stats"] is not None) == (i in (1, 3))
        if i in (1, 3):
            print(f"[slam global view {i}] GPU {out['global_stats']}, oracle {ref['global_stats']}; moved the camera by "
                  f"{np.abs(G[f'clean_v{i}_cam'] - G[f'glob_v{i}_cam']).max():.2f} mm against the run without it")
            assert out["global_stats"]["culled"] == ref["global_stats"]["culled"] == [] and out["global_stats"]["behind"] == []
        for w in trk.cam_poses:                                   # the global step rewrites EVERY camera and object
            assert _rel(trk.cam_poses[w], st.cam_poses[w]) < 1e-3
        assert set(trk.obj_poses) == set(st.obj_poses)
        for o in st.obj_poses:
            assert _rel(trk.obj_poses[o], st.obj_poses[o]) < 1e-3, o
        _check_vs_reference_fixture(G, "glob", i, trk, vid)


def test_slam_sequence_in_sfm_mode(marker_model, golden_dir):
    """ObjectSLAM(sfm_mode=True): re-initialisation test over all views (:417), its = [10, 10, 40, 40] for the per-view solve too (:843-846,
    SUO_OPT_SLAM_SFM) and the full graph after EVERY view, the first included (:443) — tracker vs oracle vs the unmodified reference (sfm_*)."""
    sd = synth.make_marker_state_dict(0)
    seq = synth.make_slam_sequence(3, n_views=3, n_obj=6)
    G = np.load(os.path.join(golden_dir, "slam_seq.npz"))
    trk = slam.SlamTracker(marker_model, sfm_mode=True)
    st = sfo.State()
    try:
        for i, v in enumerate(seq["views"]):
            a = _view_args(seq, v)
            out = trk.process_view(*a)
            ref = sfo.process_view(st, sd, *a, sfm_mode=True)
            vid = v["view_id"]
            assert out["global_stats"] is not None and ref["global_stats"] is not None
            print(f"[slam sfm view {i}] curr_only status {out['status'][:6].tolist()} (oracle {ref['ba_stats']}), global {out['global_stats']} (oracle {ref['global_stats']})")
            for w in trk.cam_poses:
                assert _rel(trk.cam_poses[w], st.cam_poses[w]) < 1e-3
            assert set(trk.obj_poses) == set(st.obj_poses)
            _check_vs_reference_fixture(G, "sfm", i, trk, vid)
    finally:
        marker_model.context().set_option(_lib.SUO_OPT_SLAM_SFM, 0)          # (the model fixture is shared by the module's tests)


@pytest.mark.parametrize("name", ["allsym", "newnon", "cv"])
def test_slam_backup_camera_pose_paths(marker_model, golden_dir, name):
    """__backup_estimate_camera_pose (lib/object_slam.py:933-973) through the tracker: the bbox-centroid PnP (one suo_pnp_batch object) or the
    constant-velocity guess on the host, then suo_slam_frame with the pose given (cam_init_mode 1: before the passes, every crop symmetric;
    cam_init_mode 2: after a failed vote, the non-symmetric group's new objects stay out of the map) — against the oracle's restatement and the
    state the unmodified reference class reaches (fixtures allsym_* / newnon_* / cv_*)."""
    sd = synth.make_marker_state_dict(0)
    seq, present, how = {"allsym": (synth.make_slam_sequence(5, n_views=3, n_obj=6, n_sym=6), None, [None, "pnp", "pnp"]),
                         "newnon": (synth.make_slam_sequence(6, n_views=3, n_obj=8, n_sym=4), lambda i: range(4) if i == 0 else range(8), [None, "pnp", "pnp"]),
                         "cv": (synth.make_slam_sequence(7, n_views=3, n_obj=3, n_sym=3), None, [None, "last", "const_vel"])}[name]
    G = np.load(os.path.join(golden_dir, "slam_seq.npz"))
    objs = seq["objs"]
    trk = slam.SlamTracker(marker_model)
    st = sfo.State()
    for i, v in enumerate(seq["views"]):
        pr = list(range(len(objs))) if present is None else list(present(i))
        a = (v["view_id"], v["img"], seq["K"], [v["dets"][c]["obj_id"] for c in pr], np.stack([v["dets"][c]["bbox"] for c in pr]),
             np.stack([objs[c]["model_kps"] for c in pr]), np.stack([objs[c]["model_kps_mask"] for c in pr]),
             np.array([objs[c]["is_symmetric"] for c in pr]), np.array([objs[c]["diameter"] for c in pr]))
        out = trk.process_view(*a)
        ref = sfo.process_view(st, sd, *a)
        vid = v["view_id"]
        assert out["backup"] == ref["backup"] == how[i] and out["cam_ok"]
        print(f"[slam backup {name} view {i}] how = {out['backup']}, cam rel diff vs oracle {_rel(trk.cam_poses[vid], st.cam_poses[vid]):.2e}, map {sorted(trk.obj_poses)}, "
              f"status {out['status'][:6].tolist()}, reinit {out['reinit_ids']}, culled {out['culled']}")
        assert _rel(trk.cam_poses[vid], st.cam_poses[vid]) < 1e-3
        assert out["reinit_ids"] == ref["reinit"]
        assert set(trk.obj_poses) == set(st.obj_poses)
        for o in st.obj_poses:
            assert _rel(trk.obj_poses[o], st.obj_poses[o]) < 1e-3, o
        _check_vs_reference_fixture(G, name, i, trk, vid)


def _noisy_gt_cam(i, v):
    """The external camera poses of the "extcam" fixture (oracle/gen_golden_slam.py): ground truth perturbed by a few mm / mrad."""
    rng = np.random.default_rng(900 + i)
    w = rng.normal(scale=2e-3, size=3)
    Wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    T = np.array(v["T_GtoC"][:3], np.float64)
    T[:, :3] = (np.eye(3) + Wx + 0.5 * Wx @ Wx) @ T[:, :3]
    T[:, 3] += rng.normal(scale=3.0, size=3)
    return T


def test_slam_views_with_external_camera_poses(marker_model, golden_dir):
    """process_view's cam_pose argument (lib/object_slam.py:349-353): the pose is given (here the ground truth perturbed by a few mm / mrad), there
    is no vote, every crop is treated as symmetric (priors from the given pose) and the curr_only solve refines the pose — suo_slam_frame with
    cam_init_mode 1, against the oracle and the unmodified reference's states (fixture extcam_*)."""
    sd = synth.make_marker_state_dict(0)
    seq = synth.make_slam_sequence(3, n_views=3, n_obj=6)
    G = np.load(os.path.join(golden_dir, "slam_seq.npz"))
    trk = slam.SlamTracker(marker_model)
    st = sfo.State()
    for i, v in enumerate(seq["views"]):
        a = _view_args(seq, v)
        out = trk.process_view(*a, cam_pose=_noisy_gt_cam(i, v))
        sfo.process_view(st, sd, *a, cam_pose=_noisy_gt_cam(i, v))
        vid = v["view_id"]
        print(f"[slam extcam view {i}] cam rel diff vs oracle {_rel(trk.cam_poses[vid], st.cam_poses[vid]):.2e}, vs ground truth "
              f"{np.linalg.norm(trk.cam_poses[vid][:, 3] - v['T_GtoC'][:3, 3]):.2f} mm (given: {np.linalg.norm(_noisy_gt_cam(i, v)[:, 3] - v['T_GtoC'][:3, 3]):.2f} mm), "
              f"priors for {int(out['prior_mask'].any(1).sum())} crops")
        assert _rel(trk.cam_poses[vid], st.cam_poses[vid]) < 1e-3 and set(trk.obj_poses) == set(st.obj_poses)
        assert int(out["prior_mask"].any(1).sum()) == (0 if i == 0 else 6)
        _check_vs_reference_fixture(G, "extcam", i, trk, vid)


def test_slam_views_at_512_with_symmetric_priors(golden_dir):
    """BASELINE configs[4] shape: 512x512 crops -> 128x128 heat-maps (the CTA-per-map reduction kernel, the 48-channel stem fed by device-rendered
    priors), 4 objects of which 2 symmetric, 2 views — the same comparison as above at the T-LESS resolution and thresholds (evaluate.py:68-76)."""
    sd = synth.make_marker_state_dict(0)
    m = PkpNet(input_res=(512, 512), max_crops=4)
    m.load_state_dict(sd)
    m.cuda().eval()
    seq = synth.make_slam_sequence(11, n_views=2, n_obj=4, res=512, n_sym=2, radius=2 * synth.MARKER_RADIUS)
    tl = dict(kp_var_thresh=0.5, bbox_thresh=1.0, manual_kp_std=0.1, init_with_outliers=True)
    trk = slam.SlamTracker(m, **tl)
    st = sfo.State()
    for v in seq["views"]:
        a = _view_args(seq, v)
        out = trk.process_view(*a)
        ref = sfo.process_view(st, sd, *a, res=512, **tl)
        vid = v["view_id"]
        assert out["cam_ok"] and ref["cam_ok"]
        for o, d in st.detections[vid].items():
            g = trk.detections[vid][o]
            if np.array_equal(g["kp_mask"], d["kp_mask"]):
                np.testing.assert_allclose(g["uv_pred"], d["uv_pred"], atol=2e-4)
                if d["prior_uv"] is not None:
                    np.testing.assert_allclose(g["prior_uv"], d["prior_uv"], atol=1e-4)
        print(f"[slam 512 view {vid}] cam rel diff {_rel(trk.cam_poses[vid], st.cam_poses[vid]):.2e}, vs ground truth "
              f"{np.linalg.norm(trk.cam_poses[vid][:, 3] - v['T_GtoC'][:3, 3]):.2f} mm, status {out['status'][:6].tolist()}")
        assert _rel(trk.cam_poses[vid], st.cam_poses[vid]) < 1e-3
        assert np.linalg.norm(trk.cam_poses[vid][:, 3] - v["T_GtoC"][:3, 3]) < 15.0
        _check_vs_reference_fixture(np.load(os.path.join(golden_dir, "slam_seq.npz")), "c5", seq["views"].index(v), trk, vid)
    sym_ids = [o["obj_id"] for o in seq["objs"] if o["is_symmetric"]]
    assert all(trk.detections[seq["views"][1]["view_id"]][o]["prior_uv"] is not None for o in sym_ids)      # the second view's symmetric crops saw priors
