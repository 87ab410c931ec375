"""GPU parity, round 2: pixels -> poses on marker frames against the CPU frame oracle, the decisive-argmax floor, the FP16
range guard through the frame call, converted checkpoints, the edge Jacobians against central differences through the C
ABI, the reference's PnP Monte-Carlo assertion (250 points), result records, the asynchronous frame batches and the g2o
drop-in's chi2 semantics."""
import os

import numpy as np
import pytest
import torch

from oracle import frame_oracle, geom
from suo_slam_b200 import _lib, ba, dist as sdist, frames, geometry, runtime, synth
from suo_slam_b200.pkpnet import PkpNet

pytestmark = pytest.mark.gpu


def _marker_batch(seed0, n_frames, crops=8, res=256):
    imgs, boxes, bi, mk, mm, kb, diam, uvgt, tgt = [], [], [], [], [], [], [], [], []
    for f in range(n_frames):
        fr = synth.make_marker_frame(seed0 + f, n_obj=crops, res=res, radius=synth.MARKER_RADIUS * res / 256)
        imgs.append(fr["img"])
        bb = [o["bbox"] for o in fr["objs"]]
        boxes += bb; bi += [f] * crops
        mk += [o["model_kps"] for o in fr["objs"]]; mm += [o["model_kps_mask"] for o in fr["objs"]]
        diam += [o["diameter"] for o in fr["objs"]]; uvgt += [o["uv_gt"] for o in fr["objs"]]; tgt += [o["T_OtoC"] for o in fr["objs"]]
        kb.append(frames.k_bbox_for(fr["K"], bb))
    return dict(img=np.stack(imgs), boxes=np.stack(boxes).astype(np.float32), bi=np.asarray(bi, np.int32), mk=np.stack(mk), mm=np.stack(mm),
                kb=np.concatenate(kb), diam=np.asarray(diam, np.float64), uv_gt=np.stack(uvgt), T_gt=np.stack(tgt))


@pytest.fixture(scope="module")
def marker_sd():
    return synth.make_marker_state_dict(0)


@pytest.fixture(scope="module")
def marker_model(marker_sd):
    m = PkpNet(input_res=(256, 256), max_crops=16)
    m.load_state_dict(marker_sd)
    m.cuda().eval()
    return m


def _decisive(ref_logits, got_logits):
    err = float(np.abs(ref_logits - got_logits).max())
    flat = ref_logits.reshape(ref_logits.shape[0], ref_logits.shape[1], -1)
    top2 = np.sort(flat, -1)[..., -2:]
    return (top2[..., 1] - top2[..., 0]) > 4 * err, flat.argmax(-1), err


def test_pixels_to_poses_on_marker_frames_vs_the_cpu_oracle(marker_sd, marker_model, golden_dir):
    """u8 frames -> suo_frames_u8 (crop, hourglass, reduction, gating, PnP, single-view BA) against the CPU frame oracle run on the
    same pixels (torch-CPU FP32 network + FP64 solver restatements): network outputs to the conv tolerance, identical gating
    away from the thresholds, decisive hard argmax bit-exact, and poses within the north-star bar wherever both sides solved
    the same keypoint set."""
    b = _marker_batch(2000, 2)
    got = frames.FramePipeline(marker_model).run(b["img"], b["boxes"], b["bi"], b["mk"], b["mm"], b["kb"], b["diam"])
    imgs = np.ascontiguousarray(b["img"].transpose(0, 3, 1, 2).astype(np.float32) / 255)
    ref = frame_oracle.run_frames(marker_sd, imgs, b["boxes"], b["bi"], b["mk"], b["mm"], b["kb"], b["diam"])
    np.testing.assert_allclose(got["uv"], ref["uv"], atol=2e-4)
    np.testing.assert_allclose(got["cov"], ref["cov"], atol=2e-4)
    std = np.sqrt(np.abs(ref["cov"][:, :, [0, 1], [0, 1]]))
    near = (np.abs(np.abs(ref["uv"]).max(-1) - 0.9) < 1e-3) | (np.abs(std - 0.4).min(-1) < 1e-3) | (np.abs(ref["kp_mask"] - 0.3) < 1e-3)
    assert np.array_equal(got["kp_used"][~near], ref["kp_used"][~near])
    assert got["kp_used"].sum() >= 16 * 6 and ref["accepted"].sum() >= 12
    # hard argmax: through the model call (the frame call does not return the maps)
    out = marker_model(torch.from_numpy(imgs).cuda(), [torch.from_numpy(b["boxes"][b["bi"] == f]).cuda() for f in range(2)])
    dec, ref_idx, err = _decisive(ref["logits"], out["prob_logits"].cpu().numpy())
    # (the marker network subtracts a threshold of 0.481 from a colour projection of ~0.5 and multiplies what is left by 1600: the FP32
    # rounding of EITHER side's stem sum shows up as ~5e-3 in the logits, so fewer maps are decisive here than with the random weights of
    # test_decisive_argmax_fraction_on_peaky_heatmaps)
    print(f"[marker 256] max|dlogit| = {err:.2e} on logits up to {np.abs(ref['logits']).max():.1f}; decisive argmax fraction = {dec.mean():.4f}")
    assert dec.mean() >= 0.5
    assert np.array_equal(out["argmax"].cpu().numpy()[dec], ref_idx[dec])
    # poses: objects whose gated keypoint set is identical on both sides (keypoints differ by the conv error, ~1e-5 NDC)
    same = ref["accepted"] & np.all(got["kp_used"] == ref["kp_used"], axis=1)
    rel = lambda a, c: np.linalg.norm(a - c) / np.linalg.norm(c)
    d_pnp = np.array([rel(got["T_pnp"][c][:3], ref["T_pnp"][c][:3]) for c in np.nonzero(same)[0]])
    d_ba = np.array([rel(got["T_ba"][c], ref["T_ba"][c]) for c in np.nonzero(same)[0]])
    print(f"[marker 256] {int(same.sum())} objects with identical gating: rel pose diff PnP median {np.median(d_pnp):.2e} max {d_pnp.max():.2e}, "
          f"BA median {np.median(d_ba):.2e} max {d_ba.max():.2e}")
    assert same.sum() >= 10
    assert np.median(d_ba) < 1e-4 and (d_ba < 1e-3).mean() >= 0.8      # an inlier flipping at the 1e-3 RANSAC threshold moves a pose by more
    t_err = np.array([rel(got["T_ba"][c][:, 3], b["T_gt"][c][:3, 3]) for c in np.nonzero(same)[0]])
    assert np.median(t_err) < 0.01
    # the same two frames through the UNMODIFIED reference ObjectSLAM(single_view_mode=True) (tests/golden/slam_seq.npz "sv_*", made by
    # oracle/gen_golden_slam.py: reference control flow + network, leaf solvers = the oracle)
    G = np.load(os.path.join(golden_dir, "slam_seq.npz"))
    g_used = np.concatenate([G[f"sv_f{f}_kp_used"] for f in range(2)]).astype(bool)
    g_inl = np.concatenate([G[f"sv_f{f}_ba_inliers"] for f in range(2)]).astype(bool)
    g_pnp, g_ba = np.concatenate([G[f"sv_f{f}_T_pnp"] for f in range(2)]), np.concatenate([G[f"sv_f{f}_T_ba"] for f in range(2)])
    same_g = np.all(got["kp_used"] == g_used, axis=1)
    d_pnp = np.array([rel(got["T_pnp"][c][:3], g_pnp[c]) for c in np.nonzero(same_g)[0]])
    d_ba = np.array([rel(got["T_ba"][c], g_ba[c]) for c in np.nonzero(same_g)[0]])
    inl_same = np.array([np.array_equal(got["ba_inliers"][c], g_inl[c]) for c in np.nonzero(same_g)[0]])
    print(f"[marker 256 vs the reference fixture] {int(same_g.sum())} of 16 objects with identical gating: rel pose diff PnP median {np.median(d_pnp):.2e} "
          f"max {d_pnp.max():.2e}, BA median {np.median(d_ba):.2e} max {d_ba.max():.2e}; identical BA inlier sets {int(inl_same.sum())}")
    assert same_g.sum() >= 10 and np.median(d_ba) < 1e-4 and (d_ba < 1e-3).mean() >= 0.8 and inl_same.mean() >= 0.8


@pytest.mark.parametrize("res", [64, 256, 512])
def test_decisive_argmax_fraction_on_peaky_heatmaps(golden_dir, res):
    """Hard argmax bit-exact wherever the reference's top-2 margin exceeds 4x the measured logit error — and with peaky heat-maps
    (random weights, last tmpOut conv x50: SURVEY.md §8d "peaky") that must be (nearly) every map: floor 0.95 at 64^2 -> 16^2,
    256^2 -> 64^2 and 512^2 -> 128^2."""
    from oracle import net_oracle
    sd = synth.make_synthetic_state_dict(0, peaky=50.0)
    m = PkpNet(input_res=(res, res), max_crops=4)
    m.load_state_dict(sd)
    m.cuda().eval()
    if res == 64:
        g = np.load(f"{golden_dir}/net_small.npz")
        imgs, boxes = torch.from_numpy(g["img"]), [torch.from_numpy(g["boxes"])]
    else:
        fr = synth.make_frame(2100 + res, n_obj=3)
        imgs = torch.from_numpy(np.ascontiguousarray(fr["img"].transpose(2, 0, 1)[None].astype(np.float32) / 255))
        boxes = [torch.from_numpy(np.stack([o["bbox"] for o in fr["objs"]]).astype(np.float32))]
    out = m(imgs.cuda(), [boxes[0].cuda()])
    ref = net_oracle.pkpnet_forward(sd, imgs, boxes, None, (res, res))
    dec, ref_idx, err = _decisive(ref["prob_logits"].numpy(), out["prob_logits"].cpu().numpy())
    print(f"[peaky x50, {res}] max|dlogit| = {err:.2e} on logits up to {float(ref['prob_logits'].abs().max()):.0f}; decisive argmax fraction = {dec.mean():.4f}")
    assert dec.mean() >= 0.95
    assert np.array_equal(out["argmax"].cpu().numpy()[dec], ref_idx[dec])
    np.testing.assert_allclose(out["uv"].cpu().numpy(), ref["uv"].numpy(), atol=2e-4)


def test_fp16_range_guard_through_the_frame_call_and_the_model_call():
    """An activation above the FP16 range (6e4) must surface as SUO_E_RANGE from the calls callers actually use — suo_frames with host
    pointers and PkpNet.forward on CUDA tensors — not as silently wrong keypoints (include/suo_b200.h, suo_check_range)."""
    sd = synth.make_synthetic_state_dict(3)
    sd["backbone.conv1_.weight"] = sd["backbone.conv1_.weight"] * 3e6
    m = PkpNet(input_res=(64, 64), max_crops=4)
    m.load_state_dict(sd)
    m.cuda().eval()
    fr = synth.make_frame(5, n_obj=2, H=120, W=160)
    bb = np.stack([o["bbox"] for o in fr["objs"]]).astype(np.float32)
    mk = np.stack([o["model_kps"] for o in fr["objs"]]); mm = np.stack([o["model_kps_mask"] for o in fr["objs"]])
    with pytest.raises(_lib.SuoError, match="FP16 range"):
        frames.FramePipeline(m).run(fr["img"][None], bb, np.zeros(2, np.int32), mk, mm, frames.k_bbox_for(fr["K"], bb), np.full(2, 150.0))
    img = torch.from_numpy(fr["img"].transpose(2, 0, 1)[None].astype(np.float32) / 255)
    with pytest.raises(_lib.SuoError, match="FP16 range"):
        m(img.cuda(), [torch.from_numpy(bb).cuda()])
    # the documented way out: tf32x3 math has the FP32 exponent range
    m.context().set_option(_lib.SUO_OPT_CONV_MATH, 0)
    out = m(img.cuda(), [torch.from_numpy(bb).cuda()])
    assert torch.isfinite(out["uv"]).all()


def test_converted_checkpoint_gives_the_same_forward(tmp_path, golden_dir):
    """Row f4: a checkpoint in the reference's format (train.py:173-181, 'module.' prefixes of DataParallelWrapper) -> .suo file ->
    PkpNet.load_packed -> forward identical to load_state_dict(checkpoint['model'])."""
    import argparse
    from suo_slam_b200 import checkpoint
    sd = synth.make_synthetic_state_dict(5, peaky=4.0)
    ck = tmp_path / "checkpoint-latest.pth.tar"
    torch.save({"args": argparse.Namespace(dataset="ycbv"), "epoch": 7, "model": {"module." + k: v for k, v in sd.items()}, "best_val": 0.1}, ck)
    out_path = tmp_path / "w.suo"
    checkpoint.convert(str(ck), str(out_path))
    g = np.load(f"{golden_dir}/net_small.npz")
    img, boxes = torch.from_numpy(g["img"]).cuda(), [torch.from_numpy(g["boxes"]).cuda()]
    a = PkpNet(input_res=(64, 64), max_crops=4)
    a.load_state_dict(sd)
    b = PkpNet(input_res=(64, 64), max_crops=4)
    b.load_packed(checkpoint.load_packed(str(out_path))[0])
    oa, ob = a.cuda()(img, boxes), b.cuda()(img, boxes)
    for k in ("uv", "cov", "prob_logits", "kp_mask", "argmax"):
        assert torch.equal(oa[k], ob[k]), k


def test_edge_jacobians_vs_central_differences_through_the_abi():
    """The recipe the reference left commented out (types_object_slam.cpp:108-122): analytic Jacobians of both edge types — as the LM
    kernels compute them (suo_edge_linearize runs the same device functions) — against central differences of the kernel's own
    error under T <- exp(d) T (g2o's numeric default, base_unary_edge.hpp:91-131), and against the oracle's analytic ones."""
    ctx = runtime.get_context()
    rng = np.random.default_rng(3)
    n = 64
    To = np.stack([np.c_[synth.random_rotation(rng), [rng.uniform(-100, 100), rng.uniform(-100, 100), rng.uniform(600, 1200)]] for _ in range(n)])
    Tc = np.stack([np.c_[synth.so3_exp(rng.normal(scale=0.2, size=3)), rng.normal(scale=20.0, size=3)] for _ in range(n)])
    k = np.tile([1066.778, -1067.487, 312.9869, 241.3109], (n, 1)) * rng.uniform(0.5, 1.5, (n, 1))      # negative fy as in bbox NDC
    p, uv = rng.uniform(-60, 60, (n, 3)), rng.uniform(0, 480, (n, 2))
    lib = _lib.lib()

    def lin(To_, Tc_, want_j=True):
        e, Ji, Jj = np.zeros((n, 2)), np.zeros((n, 2, 6)), np.zeros((n, 2, 6))
        ctx.check(lib.suo_edge_linearize(ctx.handle, n, _lib.ptr(np.ascontiguousarray(To_.reshape(n, 12))) if To_ is not None else None,
                                         _lib.ptr(np.ascontiguousarray(Tc_.reshape(n, 12))), _lib.ptr(k), _lib.ptr(p), _lib.ptr(uv), _lib.ptr(e),
                                         _lib.ptr(Ji) if want_j and To_ is not None else None, _lib.ptr(Jj) if want_j else None, 0, None))
        return e, Ji, Jj

    e0, Ji, Jj = lin(To, Tc)
    h = 1e-6
    for q in range(6):
        d = np.zeros(6); d[q] = h
        plus = lambda T, s: np.stack([geom.se3_oplus(T[i], s * d) for i in range(n)])
        fi = (lin(plus(To, 1), Tc, False)[0] - lin(plus(To, -1), Tc, False)[0]) / (2 * h)
        fj = (lin(To, plus(Tc, 1), False)[0] - lin(To, plus(Tc, -1), False)[0]) / (2 * h)
        np.testing.assert_allclose(Ji[:, :, q], fi, rtol=2e-6, atol=2e-5)
        np.testing.assert_allclose(Jj[:, :, q], fj, rtol=2e-6, atol=2e-5)
    for i in range(0, n, 7):      # the oracle's analytic Jacobians and error
        eo, Jio, Jjo = geom.edge_eval(To[i], Tc[i], k[i], p[i], uv[i])
        np.testing.assert_allclose(e0[i], eo, rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(Ji[i], Jio, rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(Jj[i], Jjo, rtol=1e-10, atol=1e-9)
    # unary edge (EdgeSE3ProjectFromFixedObject): T_obj == NULL, p is p_inG
    pw = np.stack([To[i][:, :3] @ p[i] + To[i][:, 3] for i in range(n)])
    e1, _, Jj1 = np.zeros((n, 2)), None, np.zeros((n, 2, 6))
    ctx.check(lib.suo_edge_linearize(ctx.handle, n, None, _lib.ptr(np.ascontiguousarray(Tc.reshape(n, 12))), _lib.ptr(k), _lib.ptr(np.ascontiguousarray(pw)),
                                     _lib.ptr(uv), _lib.ptr(e1), None, _lib.ptr(Jj1), 0, None))
    np.testing.assert_allclose(e1, e0, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(Jj1, Jj, rtol=1e-12, atol=1e-9)


def test_pnp_monte_carlo_250_points_half_outliers():
    """thirdparty/lambdatwist/test_pnp.cpp:68-147 on suo_pnp_batch: 250 points, 50 % outliers, sigma in {0, .25, .5, 1} px, failure =
    angle + |t| error > 0.05, fewer than 5 % failures per sigma (200 experiments per sigma in one launch), no NaN; and every
    experiment equal to the CPU oracle (same hypotheses, same winner, same refined pose)."""
    for si, sigma in enumerate((0.0, 0.25, 0.5, 1.0)):
        data = [synth.make_pnp_benchmark(10_000 * si + e, 250, sigma, 0.5) for e in range(200)]
        T, st = geometry.pnp_batch([d[0] for d in data], [d[1] for d in data], seed=0, return_stats=True)
        assert np.isfinite(T).all()
        errs = np.array([synth.pnp_benchmark_error(T[e], data[e][2]) for e in range(200)])
        fail = (errs > 0.05).mean()
        print(f"[pnp monte-carlo] sigma {sigma}: failures {fail:.3f}, median err {np.median(errs):.2e}, median RANSAC iterations {np.median(st[:, 2]):.0f}")
        assert fail < 0.05
        for e in range(0, 200, 20):
            To, so = geom.lambdatwist_pnp(data[e][0], data[e][1], seed=0, obj_key=e)
            assert st[e, 0] == so["best_inliers"] and st[e, 1] == so["best_iter"] and st[e, 2] == so["total_iters"]
            np.testing.assert_allclose(T[e], To, rtol=1e-8, atol=1e-8 * max(1.0, np.abs(To).max()))


def test_pnp_device_call_never_truncates():
    """More points than the device-pointer capacity (SUO_OPT_PNP_MAX_POINTS, default 64): the object fails loudly (identity, stats -1) —
    it is not solved on its first 64 points — and runs once the option is raised."""
    ctx = runtime.get_context()
    rng = np.random.default_rng(4)
    X = rng.uniform(-60, 60, (100, 3))
    R, t = synth.random_rotation(rng), np.array([10.0, -5.0, 700.0])
    pc = X @ R.T + t
    y = pc[:, :2] / pc[:, 2:3]
    dev = torch.device("cuda", ctx.device)
    xs, ys, off = torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), torch.tensor([0, 100], dtype=torch.int32, device=dev)
    T, st = torch.zeros((1, 16), dtype=torch.float64, device=dev), torch.zeros((1, 5), dtype=torch.int32, device=dev)
    call = lambda: ctx.check(_lib.lib().suo_pnp_batch(ctx.handle, _lib.ptr(xs), _lib.ptr(ys), _lib.ptr(off), 1, 0.001, 0, None, _lib.ptr(T), _lib.ptr(st), 1, None))
    call()
    torch.cuda.synchronize()
    assert np.array_equal(T.cpu().numpy().reshape(4, 4), np.eye(4)) and st.cpu().numpy()[0, 0] == -1
    ctx.set_option(_lib.SUO_OPT_PNP_MAX_POINTS, 128)
    call()
    torch.cuda.synchronize()
    ctx.set_option(_lib.SUO_OPT_PNP_MAX_POINTS, 64)
    Tg = T.cpu().numpy().reshape(4, 4)
    np.testing.assert_allclose(Tg[:3, :3], R, atol=1e-8)
    np.testing.assert_allclose(Tg[:3, 3], t, atol=1e-5)
    assert st.cpu().numpy()[0, 0] == 100


def test_result_records_device_packing_equals_the_host_layout(marker_model):
    """suo_pack_records (one device kernel) writes exactly the bytes dist.pack_records_host defines (the layout the CPU gloo test
    exchanges): poses, acceptance flag, counts, keypoints, covariances, per-keypoint flags."""
    b = _marker_batch(2300, 2)
    got = frames.FramePipeline(marker_model).run(b["img"], b["boxes"], b["bi"], b["mk"], b["mm"], b["kb"], b["diam"])
    L = len(b["bi"])
    ctx = marker_model.context()
    rb = int(_lib.lib().suo_record_bytes(41))
    assert rb == sdist.record_bytes(41)
    rec = np.zeros((L, rb), np.uint8)
    ids = np.arange(100, 100 + L, dtype=np.int32)
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    a = [c(got["T_pnp"], np.float64), c(got["T_ba"], np.float64), c(got["kp_used"], np.uint8), c(got["ba_inliers"], np.uint8), c(got["uv"], np.float32),
         c(got["cov"], np.float32)]                        # (kept alive across the call: the library reads them through raw pointers)
    ctx.check(_lib.lib().suo_pack_records(ctx.handle, _lib.ptr(ids), 0, *[_lib.ptr(v) for v in a], L, _lib.ptr(rec), 0, None))
    want = sdist.pack_records_host(ids, got["T_pnp"], got["T_ba"], got["kp_used"], got["ba_inliers"], got["uv"], got["cov"])
    have = np.frombuffer(rec.tobytes(), dtype=sdist.record_dtype(41))
    for name in want.dtype.names:
        assert np.array_equal(have[name], want[name]), name
    assert rec.tobytes() == want.tobytes()
    u = sdist.unpack_records(rec)
    assert np.array_equal(u["crop_id"], ids) and u["accepted"].sum() >= 10 and np.array_equal(u["n_used"], got["kp_used"].sum(1))


def test_asynchronous_frame_batches_equal_the_synchronous_call(marker_model):
    """suo_frames_u8_submit / suo_frames_wait on two slots (copies of batch i+1 overlapping batch i) return exactly what suo_frames_u8
    returns batch by batch, and the records packed on the device by the submit are the records of those results."""
    ctx = marker_model.context()
    lib, p = _lib.lib(), _lib.ptr
    batches = [_marker_batch(2400 + 10 * i, 2) for i in range(3)]
    pipe = frames.FramePipeline(marker_model)
    want = [pipe.run(b["img"], b["boxes"], b["bi"], b["mk"], b["mm"], b["kb"], b["diam"]) for b in batches]
    L, K = 16, 41
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    ins = [dict(img=pin(b["img"]), boxes=pin(b["boxes"]), bi=pin(b["bi"]), mk=pin(b["mk"]), mm=pin(b["mm"].astype(np.uint8)), kb=pin(b["kb"]), diam=pin(b["diam"])) for b in batches]
    outs = [dict(T_pnp=torch.zeros((L, 16), dtype=torch.float64).pin_memory(), T_ba=torch.zeros((L, 12), dtype=torch.float64).pin_memory(),
                 used=torch.zeros((L, K), dtype=torch.uint8).pin_memory(), bain=torch.zeros((L, K), dtype=torch.uint8).pin_memory(),
                 uv=torch.zeros((L, K, 2)).pin_memory(), cov=torch.zeros((L, K, 4)).pin_memory()) for _ in range(3)]
    rb = int(lib.suo_record_bytes(K))
    recs = [torch.zeros((L, rb), dtype=torch.uint8, device="cuda") for _ in range(3)]
    stream = torch.cuda.current_stream().cuda_stream

    def submit(i, slot=None, stream=stream):
        a, o = ins[i], outs[i]
        ctx.check(lib.suo_frames_u8_submit(ctx.handle, i % 2 if slot is None else slot, p(a["img"]), 2, 480, 640, p(a["boxes"]), p(a["bi"]), L, p(a["mk"]), p(a["mm"]), p(a["kb"]), p(a["diam"]),
                                           0.2, 0.9, 0, 1, p(o["T_pnp"]), p(o["T_ba"]), p(o["used"]), p(o["bain"]), p(o["uv"]), p(o["cov"]), p(recs[i]), 1000 * i, stream))
    submit(0)
    other = torch.cuda.Stream()
    with pytest.raises(_lib.SuoError, match="same stream"):
        submit(1, stream=other.cuda_stream)           # the two slots share the executor: one stream only
    submit(1)
    with pytest.raises(_lib.SuoError, match="pending"):
        submit(2)                                     # slot 0 has not been waited for
    ctx.check(lib.suo_frames_wait(ctx.handle, 0))
    submit(2)
    ctx.check(lib.suo_frames_wait(ctx.handle, 1))
    ctx.check(lib.suo_frames_wait(ctx.handle, 0))
    torch.cuda.synchronize()
    for i in range(3):
        o, w = outs[i], want[i]
        assert np.array_equal(o["T_pnp"].numpy().reshape(L, 4, 4), w["T_pnp"]) and np.array_equal(o["T_ba"].numpy().reshape(L, 3, 4), w["T_ba"])
        assert np.array_equal(o["used"].numpy().astype(bool), w["kp_used"]) and np.array_equal(o["bain"].numpy().astype(bool), w["ba_inliers"])
        assert np.array_equal(o["uv"].numpy(), w["uv"]) and np.array_equal(o["cov"].numpy().reshape(L, K, 2, 2), w["cov"])
        host = sdist.pack_records_host(np.arange(1000 * i, 1000 * i + L), w["T_pnp"], w["T_ba"], w["kp_used"], w["ba_inliers"], w["uv"], w["cov"])
        assert recs[i].cpu().numpy().tobytes() == host.tobytes()


def test_frame_pipeline_stream_equals_run(marker_model):
    """FramePipeline.stream (the Python face of suo_frames_u8_submit / suo_frames_wait) yields, in order, exactly what run() returns
    batch by batch — for an odd and an even number of batches, a batch size that changes, and a consumer that stops early."""
    pipe = frames.FramePipeline(marker_model)
    batches = [_marker_batch(2600 + 10 * i, 1 + i % 2) for i in range(5)]
    keys = ("img", "boxes", "bi", "mk", "mm", "kb", "diam")
    names = ("images", "boxes", "box_img", "model_kps", "model_mask", "K_bbox", "diameter")
    as_kw = lambda b: {n: b[k] for n, k in zip(names, keys)}
    want = [pipe.run(**as_kw(b)) for b in batches]
    for n in (5, 4, 1):
        got = list(pipe.stream(as_kw(b) for b in batches[:n]))
        assert len(got) == n
        for g, w in zip(got, want):
            for k in w:
                assert np.array_equal(g[k], w[k]), k
    it = pipe.stream(as_kw(b) for b in batches)
    first = next(it)
    it.close()                                         # two batches are in flight: the generator's cleanup waits for them
    assert np.array_equal(first["T_ba"], want[0]["T_ba"])
    again = pipe.run(**as_kw(batches[0]))              # the context is usable (no slot left pending)
    assert np.array_equal(again["T_pnp"], want[0]["T_pnp"])


def test_g2o_dropin_chi2_is_the_error_the_kernel_left_in_the_edges():
    """ObjectSLAM.optimize() re-reads e.chi2() of inlier edges WITHOUT recomputing (lib/object_slam.py:881-883), so after a rejected LM
    trial it classifies with the rejected state's error.  The g2o drop-in copies the kernel's leftover errors into the edges
    (suo_ba_last_errors): the reference's 4-round flow through the drop-in must therefore classify and converge EXACTLY like the
    packed rounds of suo_ba_batch (which the oracle tests pin against the g2o restatement)."""
    from suo_slam_b200 import g2o
    from tests.test_gpu_geom import _reference_optimize_loop
    n_obj, n_kp = 6, 12
    for seed, its in ((31, [10, 10, 10, 10]), (32, [10, 10, 40, 40]), (33, [2, 2, 2, 2])):
        pr = synth.make_ba_problem(seed, n_obj, n_kp, noise_px=1.0, outlier_frac=0.2)
        opt = g2o.SparseOptimizer()
        opt.set_algorithm(g2o.OptimizationAlgorithmLevenberg(g2o.BlockSolverSE3(g2o.LinearSolverCholmodSE3())))
        obj_v = []
        for j in range(n_obj):
            v = g2o.VertexSE3Expmap(); v.set_id(j); v.set_estimate(g2o.SE3Quat(pr["T_init"][j][:, :3], pr["T_init"][j][:, 3]))
            opt.add_vertex(v); obj_v.append(v)
        cam = g2o.VertexSE3Expmap(); cam.set_id(n_obj); cam.set_estimate(g2o.SE3Quat(np.eye(3), np.zeros(3))); cam.set_fixed(True)
        opt.add_vertex(cam)
        edges = []
        for j in range(n_obj):
            for q in range(n_kp):
                e = g2o.EdgeSE3ProjectFromObject(pr["cam_k"], pr["p_O"][j, q])
                e.set_vertex(0, obj_v[j]); e.set_vertex(1, cam); e.set_measurement(pr["uv"][j, q]); e.set_information(pr["info"][j, q])
                e.set_robust_kernel(g2o.RobustKernelHuber(np.sqrt(5.991))); e.set_level(0)
                edges.append(e); opt.add_edge(e)
        inl = _reference_optimize_loop(opt, edges, its)
        got = np.stack([v.estimate().matrix()[:3] for v in obj_v])
        poses = np.concatenate([pr["T_init"], np.hstack([np.eye(3), np.zeros((3, 1))])[None]], 0)
        fixed = np.zeros(n_obj + 1, np.uint8); fixed[n_obj] = 1
        P, io, _ = ba.ba_batch([0, n_obj + 1], [0, n_obj * n_kp], poses, fixed, np.repeat(np.arange(n_obj), n_kp), np.full(n_obj * n_kp, n_obj),
                               np.tile(pr["cam_k"], (n_obj * n_kp, 1)), pr["p_O"], pr["uv"], pr["info"], np.ones(n_obj * n_kp), its)
        assert np.array_equal(inl, io), (seed, its)
        # (same classification; the poses agree to the size of the last LM steps: a trial whose chi2 gain is pure rounding is accepted in one
        # packing of the edges and rejected in the other)
        np.testing.assert_allclose(got, P[:n_obj], rtol=1e-5, atol=1e-5)
