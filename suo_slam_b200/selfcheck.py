"""smoke(): one small invocation of every hot-path stage on cuda:0, each checked against the
CPU oracle (oracle/ is test infrastructure; this module is only reached from
__graft_entry__.smoke())."""
from __future__ import annotations

import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_smoke(verbose: bool = False):
    from oracle import frame_oracle, geom
    from . import _lib, frames, synth
    from .pkpnet import PkpNet

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
    _ = geom
    _ = _lib
    say("ok")
