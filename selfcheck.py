"""smoke(): one small invocation of every hot-path stage on cuda:0, each checked against the
CPU oracle (oracle/ is test infrastructure; this module sits next to __graft_entry__.py, outside the
product package, and is only reached from __graft_entry__.smoke())."""
from __future__ import annotations

import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


def run_smoke(verbose: bool = False):
    from oracle import frame_oracle, geom
    from suo_slam_b200 import _lib, frames, synth
    from suo_slam_b200.pkpnet import PkpNet

    def say(*a):
        if verbose:
            print("[smoke]", *a, flush=True)

    g = np.load(os.path.join(ROOT, "tests", "golden", "net_small.npz"))
    sd = synth.make_synthetic_state_dict(0, peaky=4.0)
    m = PkpNet(input_res=(64, 64), max_crops=4)
    m.load_state_dict(sd)
    m.cuda(0).eval()
    out = m(torch.from_numpy(g["img"]).cuda(), [torch.from_numpy(g["boxes"]).cuda()], None)
    torch.cuda.synchronize()
    err = float(np.abs(out["prob_logits"].cpu().numpy() - g["logits"]).max())
    say(f"network (tcgen05, fp16x3 split math, A-halo / CTA-pair 3x3 convs) vs reference golden: max |dlogit| = {err:.2e}")
    assert err < 3e-3, err
    np.testing.assert_allclose(out["uv"].cpu().numpy(), g["uv"], atol=5e-5)
    np.testing.assert_allclose(out["cov"].cpu().numpy(), g["cov"], atol=5e-5)

    # keypoints -> poses on identical inputs (PnP + single-view BA) vs the oracle
    fr = synth.make_frame(7, n_obj=4)
    uv = np.stack([o["uv_meas"] for o in fr["objs"]]).astype(np.float32)
    cov = np.stack([o["cov"] for o in fr["objs"]]).astype(np.float32)
    km = np.full(uv.shape[:2], 0.9, np.float32)
    mk = np.stack([o["model_kps"] for o in fr["objs"]])
    mm = np.stack([o["model_kps_mask"] for o in fr["objs"]])
    kb = frames.k_bbox_for(fr["K"], [o["bbox"] for o in fr["objs"]])
    diam = np.full(4, 150.0)
    bi = np.zeros(4, np.int32)
    got = frames.solve_keypoints(m.context(), uv, cov, km, bi, mk, mm, kb, diam, seed=1)
    ref = frame_oracle.solve_from_keypoints(uv, cov, km, mk, mm, kb, diam, bi, seed=1)
    assert np.array_equal(got["kp_used"], ref["kp_used"])
    np.testing.assert_allclose(got["T_pnp"], ref["T_pnp"], rtol=1e-7, atol=1e-5)
    np.testing.assert_allclose(got["T_ba"], ref["T_ba"], rtol=1e-7, atol=1e-5)
    say(f"PnP + BA vs oracle: {int(ref['accepted'].sum())}/4 objects accepted, poses match to 1e-7; "
        f"{m.context().kernel_launches()} kernel launches")
    # pixels -> poses: one marker frame (4 objects) through the fused frame call, against the CPU frame oracle
    fr = synth.make_marker_frame(1000, n_obj=4)
    sdm = synth.make_marker_state_dict(0)
    mm2 = PkpNet(input_res=(256, 256), max_crops=4)
    mm2.load_state_dict(sdm)
    mm2.cuda(0).eval()
    bb = np.stack([o["bbox"] for o in fr["objs"]]).astype(np.float32)
    mk = np.stack([o["model_kps"] for o in fr["objs"]])
    msk = np.stack([o["model_kps_mask"] for o in fr["objs"]])
    kb = frames.k_bbox_for(fr["K"], bb)
    got = frames.FramePipeline(mm2).run(fr["img"][None], bb, bi, mk, msk, kb, diam)
    ref = frame_oracle.run_frames(sdm, fr["img"].transpose(2, 0, 1)[None].astype(np.float32) / 255, bb, bi, mk, msk, kb, diam)
    np.testing.assert_allclose(got["uv"], ref["uv"], atol=2e-4)
    same = np.array_equal(got["kp_used"], ref["kp_used"])
    acc = ref["accepted"] & np.all(got["kp_used"] == ref["kp_used"], axis=1)
    d = [np.linalg.norm(got["T_ba"][c] - ref["T_ba"][c]) / np.linalg.norm(ref["T_ba"][c]) for c in np.nonzero(acc)[0]]
    say(f"marker frame, pixels -> poses vs CPU oracle: {int(got['kp_used'].sum())} gated keypoints (same gating: {same}), "
        f"{int(ref['accepted'].sum())}/4 objects accepted, max rel pose diff {max(d) if d else float('nan'):.2e}")
    assert got["kp_used"].sum() >= 16 and ref["accepted"].sum() >= 2 and d and max(d) < 1e-2
    _ = geom
    _ = _lib
    say("ok")
