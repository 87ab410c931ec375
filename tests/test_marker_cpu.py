"""CPU: the synthetic fiducial ("marker") network and frames that make the frame path do real solver work
(suo_slam_b200/synth.py), checked through the CPU oracle; and the reference's own PnP Monte-Carlo assertion
(thirdparty/lambdatwist/test_pnp.cpp:68-147) on the oracle's RANSAC + refine restatement."""
import numpy as np
import torch

from oracle import frame_oracle, geom
from suo_slam_b200 import arch, frames, synth, weights


def test_marker_codes_are_separable_and_away_from_black():
    d = synth.marker_codes()
    assert d.shape == (41, 3) and np.allclose(np.linalg.norm(d, axis=1), 1.0)
    G = d @ d.T
    np.fill_diagonal(G, -1)
    # own colour projects to 0.5, the closest other colour to 0.5 * max cos: the threshold sits between them
    assert 0.5 * G.max() + 0.015 < synth.MARKER_T < 0.5 - 0.015
    # zero padding at the crop border (black) must stay below the threshold for every detector
    assert (-0.5 * d.sum(1)).max() < synth.MARKER_T - 0.1
    c = synth.marker_colors_u8()
    proj = (c / 255.0 - 0.5) @ d.T                                   # u8 rounding keeps the margins
    assert np.all(np.diag(proj) > synth.MARKER_T + 0.01)
    np.fill_diagonal(proj, -1)
    assert proj.max() < synth.MARKER_T - 0.01


def test_marker_state_dict_is_a_reference_format_state_dict():
    sd = synth.make_marker_state_dict(0)
    spec = dict(arch.state_dict_spec())
    assert set(sd) == set(spec)
    for k, shp in spec.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    info = weights.program_summary(weights.pack_state_dict(sd))
    assert info["n_convs"] == 185
    # dense: outside the hand-wired signal rows every conv keeps its seeded random weights
    w = sd["backbone.hourglass.1.up1_.0.conv2.weight"]
    assert float((w != 0).float().mean()) > 0.99


def test_marker_frame_pixels_to_poses_through_the_oracle():
    """One 8-crop frame: the gate passes real keypoints, every object gets a PnP pose and a non-empty BA, the poses are
    close to the frame's ground truth, and the heat-maps are peaky (top-2 logit margins far above the conv error)."""
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    sd = synth.make_marker_state_dict(0)
    fr = synth.make_marker_frame(1000)
    img = fr["img"].transpose(2, 0, 1).astype(np.float32)[None] / 255
    bb = np.stack([o["bbox"] for o in fr["objs"]]).astype(np.float32)
    bi = np.zeros(8, np.int32)
    mk = np.stack([o["model_kps"] for o in fr["objs"]])
    mm = np.stack([o["model_kps_mask"] for o in fr["objs"]])
    kb = frames.k_bbox_for(fr["K"], bb)
    res = frame_oracle.run_frames(sd, img, bb, bi, mk, mm, kb, np.full(8, 150.0))
    used = res["kp_used"]
    assert used.sum() >= 8 * 6 and (used.sum(1) >= 4).all(), used.sum(1)
    assert res["accepted"].sum() >= 7
    assert res["ba_inliers"].sum() >= 40
    uv_gt = np.stack([o["uv_gt"] for o in fr["objs"]])
    err = np.abs(res["uv"] - uv_gt).max(-1)[used]
    assert np.median(err) < 0.01, np.median(err)                     # < 1.3 crop pixels
    T_gt = np.stack([o["T_OtoC"] for o in fr["objs"]])
    terr = [np.linalg.norm(res["T_ba"][c][:, 3] - T_gt[c][:3, 3]) / np.linalg.norm(T_gt[c][:3, 3]) for c in np.nonzero(res["accepted"])[0]]
    assert np.median(terr) < 0.01, terr
    flat = np.sort(res["logits"].reshape(8, 41, -1), -1)
    margin = flat[..., -1] - flat[..., -2]
    assert (margin > 1e-3).mean() > 0.95, (margin > 1e-3).mean()      # decisive hard argmax (4 x the conv error of ~2e-4)
    assert flat[..., -1][used].min() > 20.0                           # gated keypoints sit on a real peak


def test_oracle_pnp_passes_the_reference_monte_carlo_assertion():
    """test_pnp.cpp:68-147 (250 points, 50 % outliers, sigma in {0, .25, .5, 1} px): failure = angle + |t| error > 0.05,
    fewer than 5 % failures per sigma.  100 experiments per sigma here (the reference runs 1000)."""
    for si, sigma in enumerate((0.0, 0.25, 0.5, 1.0)):
        errs = []
        for e in range(100):
            xs, ys, P = synth.make_pnp_benchmark(10_000 * si + e, 250, sigma, 0.5)
            T, _ = geom.lambdatwist_pnp(xs, ys, seed=0, obj_key=e)
            assert np.isfinite(T).all()
            errs.append(synth.pnp_benchmark_error(T, P))
        assert (np.array(errs) > 0.05).mean() < 0.05, (sigma, (np.array(errs) > 0.05).mean())


def test_slam_frame_oracle_tracks_a_marker_sequence():
    """The CPU restatement of ObjectSLAM.process_view in SLAM mode (two forwards per view, priors for the symmetric objects, camera-pose
    vote, curr_only LM) recovers the ground-truth camera motion of a synthetic sequence to a few millimetres over ~1 m."""
    from oracle import slam_frame_oracle as sfo
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    sd = synth.make_marker_state_dict(0)
    seq = synth.make_slam_sequence(3, n_views=3, n_obj=6)
    st = sfo.State()
    objs = seq["objs"]
    for v in seq["views"]:
        r = sfo.process_view(st, sd, v["view_id"], v["img"], seq["K"], [d["obj_id"] for d in v["dets"]], np.stack([d["bbox"] for d in v["dets"]]),
                             np.stack([o["model_kps"] for o in objs]), np.stack([o["model_kps_mask"] for o in objs]),
                             np.array([o["is_symmetric"] for o in objs]), np.array([o["diameter"] for o in objs]))
        assert r["cam_ok"] and r["reinit"] == []
        cam = st.cam_poses[v["view_id"]]
        assert np.linalg.norm(cam[:, 3] - v["T_GtoC"][:3, 3]) < 10.0 and np.abs(cam[:, :3] - v["T_GtoC"][:3, :3]).max() < 0.01
        sym_with_prior = [o["obj_id"] for o in objs if o["is_symmetric"] and st.detections[v["view_id"]][o["obj_id"]]["prior_uv"] is not None]
        assert (len(sym_with_prior) == 3) == (v["view_id"] != 100)        # priors exist once the symmetric objects are in the map
    assert len(st.obj_poses) == 6
