"""GPU parity of the whole network forward (rows a1-a4) through the reference call surface
``model(images, boxes, prior_kp)``, against fixtures made by the UNMODIFIED reference."""
import numpy as np
import pytest
import torch

from oracle import net_oracle
from suo_slam_b200 import _lib, synth
from suo_slam_b200.pkpnet import PkpNet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sd():
    return synth.make_synthetic_state_dict(seed=0, peaky=4.0)


def _model(sd, backend, passes, res=64, max_crops=8):
    """backend 0 = SIMT FP32, 1 = tcgen05 TF32 (passes 1|3), 2 = tcgen05 FP16x3"""
    m = PkpNet(input_res=(res, res), max_crops=max_crops)
    m.load_state_dict(sd)
    m.cuda().eval()
    m.context().set_option(_lib.SUO_OPT_CONV_MATH, 1 if backend == 2 else 0)
    backend = min(backend, 1)
    m.context().set_option(_lib.SUO_OPT_CONV_BACKEND, backend)
    m.context().set_option(_lib.SUO_OPT_TF32_PASSES, passes)
    return m


def _margin_ok(logits, argmax, err):
    """indices must agree wherever the reference's top-2 margin exceeds the logit error bound"""
    flat = logits.reshape(logits.shape[0], logits.shape[1], -1)
    top2 = np.sort(flat, -1)[..., -2:]
    ref_idx = flat.argmax(-1)
    decisive = (top2[..., 1] - top2[..., 0]) > 4 * err
    return decisive, ref_idx


@pytest.mark.parametrize("backend,passes,tol_logit,tol_uv", [(0, 3, 2e-4, 2e-5), (1, 3, 3e-4, 5e-5), (1, 1, 0.3, 2e-2), (2, 3, 3e-4, 5e-5)])
@pytest.mark.parametrize("name", ["net_small", "net_prior"])
def test_forward_vs_reference_golden(golden_dir, sd, name, backend, passes, tol_logit, tol_uv):
    g = np.load(f"{golden_dir}/{name}.npz")
    m = _model(sd, backend, passes)
    prior = None if g["prior"].size == 0 else [torch.from_numpy(g["prior"]).cuda()]
    out = m(torch.from_numpy(g["img"]).cuda(), [torch.from_numpy(g["boxes"]).cuda()], prior)
    torch.cuda.synchronize()
    logits = out["prob_logits"].cpu().numpy()
    err = np.abs(logits - g["logits"]).max()
    print(f"[{name} backend={backend} passes={passes}] max|dlogit|={err:.3e} max|duv|={np.abs(out['uv'].cpu().numpy() - g['uv']).max():.3e} "
          f"max|dcov|={np.abs(out['cov'].cpu().numpy() - g['cov']).max():.3e}")
    assert err < tol_logit * max(1.0, np.abs(g["logits"]).max() / 10), f"logit err {err}"
    np.testing.assert_allclose(out["uv"].cpu().numpy(), g["uv"], atol=tol_uv)
    np.testing.assert_allclose(out["cov"].cpu().numpy(), g["cov"], atol=tol_uv)
    np.testing.assert_allclose(out["kp_mask"].cpu().numpy(), g["kp_mask"], atol=max(tol_uv, 1e-4))
    decisive, ref_idx = _margin_ok(g["logits"], None, err)
    got_idx = out["argmax"].cpu().numpy()
    if passes == 3:
        assert decisive.mean() > 0.5
    assert np.array_equal(got_idx[decisive], ref_idx[decisive])          # bit-exact hard argmax where decisive
    assert set(out.keys()) >= {"uv", "cov", "prob_logits", "prob", "kp_mask_logits", "kp_mask"}
    assert out["uv"].shape == (3, 41, 2) and out["cov"].shape == (3, 41, 2, 2) and out["prob"].shape == (3, 41, 16, 16)


def test_forward_host_tensors_and_graph_replay(sd, golden_dir):
    """CPU tensors in -> CPU tensors out (host-pointer ABI path); second call replays the CUDA graph."""
    g = np.load(f"{golden_dir}/net_small.npz")
    m = _model(sd, 2, 3)
    a = m(torch.from_numpy(g["img"]), [torch.from_numpy(g["boxes"])], None)
    b = m(torch.from_numpy(g["img"]), [torch.from_numpy(g["boxes"])], None)
    assert a["uv"].device.type == "cpu"
    np.testing.assert_array_equal(a["prob_logits"].numpy(), b["prob_logits"].numpy())     # deterministic
    np.testing.assert_allclose(a["uv"].numpy(), g["uv"], atol=5e-5)
    assert m.context().kernel_launches() > 400


@pytest.mark.parametrize("backend", [1, 2])
def test_forward_256_vs_oracle(sd, backend):
    """BASELINE config-2 shape: 8 crops of a 640x480 frame at 256x256 -> 64x64 heat-maps."""
    fr = synth.make_frame(1)
    img = torch.from_numpy(fr["img"].transpose(2, 0, 1).astype(np.float32) / 255)[None]
    boxes = torch.from_numpy(np.stack([o["bbox"] for o in fr["objs"]]))
    m = _model(sd, backend, 3, res=256, max_crops=8)
    out = m(img.cuda(), [boxes.cuda()], None)
    torch.cuda.synchronize()
    ref = net_oracle.pkpnet_forward(sd, img, [boxes], None, (256, 256))
    lg, lr = out["prob_logits"].cpu().numpy(), ref["prob_logits"].numpy()
    err = np.abs(lg - lr).max()
    print(f"[256 backend={backend}] max|dlogit|={err:.3e} max|duv|={np.abs(out['uv'].cpu().numpy() - ref['uv'].numpy()).max():.3e}")
    assert err < 5e-4 * max(1.0, np.abs(lr).max() / 10), err
    np.testing.assert_allclose(out["uv"].cpu().numpy(), ref["uv"].numpy(), atol=5e-5)
    np.testing.assert_allclose(out["cov"].cpu().numpy(), ref["cov"].numpy(), atol=5e-5)
    decisive, ref_idx = _margin_ok(lr, None, err)
    assert np.array_equal(out["argmax"].cpu().numpy()[decisive], ref_idx[decisive])


def test_forward_tless_shape_with_priors(sd):
    """BASELINE config 5 shape: 512x512 crops -> 128x128 heat-maps, mixed objects with / without prior
    planes (symmetric objects get a rendered prior, lib/object_slam.py:486-514)."""
    rng = np.random.default_rng(5)
    img = torch.from_numpy(rng.random((1, 3, 480, 640), dtype=np.float32))
    boxes = torch.tensor([[40.0, 30.0, 420.0, 400.0], [300.0, 100.0, 630.0, 470.0]])
    prior = torch.zeros(2, 41, 512, 512)
    yy, xx = torch.meshgrid(torch.arange(512.0), torch.arange(512.0), indexing="ij")
    for k in range(0, 41, 3):                       # object 1 "symmetric": Gaussian blobs on a third of the planes
        cy, cx = rng.uniform(60, 450, 2)
        prior[1, k] = torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * 14.0 ** 2))
    m = _model(sd, 2, 3, res=512, max_crops=2)
    out = m(img.cuda(), [boxes.cuda()], [prior.cuda()])
    torch.cuda.synchronize()
    ref = net_oracle.pkpnet_forward(sd, img, [boxes], [prior], (512, 512))
    assert out["prob_logits"].shape == (2, 41, 128, 128)
    lr = ref["prob_logits"].numpy()
    err = np.abs(out["prob_logits"].cpu().numpy() - lr).max()
    assert err < 5e-4 * max(1.0, np.abs(lr).max() / 10), err
    np.testing.assert_allclose(out["uv"].cpu().numpy(), ref["uv"].numpy(), atol=5e-5)
    np.testing.assert_allclose(out["cov"].cpu().numpy(), ref["cov"].numpy(), atol=5e-5)
    decisive, ref_idx = _margin_ok(lr, None, err)
    assert np.array_equal(out["argmax"].cpu().numpy()[decisive], ref_idx[decisive])


def test_forward_errors(sd):
    m = PkpNet(input_res=(64, 64), max_crops=2)
    with pytest.raises(_lib.SuoError):
        m(torch.zeros(1, 3, 32, 32), [torch.tensor([[0.0, 0.0, 8.0, 8.0]])])           # no weights loaded
    m.load_state_dict(sd)
    with pytest.raises(_lib.SuoError):
        m(torch.zeros(1, 3, 32, 32), [torch.tensor([[0.0, 0.0, 8.0, 8.0]] * 3)])       # more crops than max_crops
    with pytest.raises(AssertionError):
        m(torch.zeros(2, 3, 32, 32), [torch.tensor([[0.0, 0.0, 8.0, 8.0]])])           # len(boxes) != batch (pkpnet.py:91)


def test_forward_with_keypoint_priors_equals_dense_priors(sd):
    """Row f2 fused: model(images, boxes, prior_uv=[(uv, mask)]) stamps the prior planes into the network input on
    the device; the result must be IDENTICAL to feeding the reference's dense planes through prior_kp."""
    from oracle import prior_oracle
    rng = np.random.default_rng(12)
    img = torch.from_numpy(rng.random((2, 3, 120, 160), dtype=np.float32))
    boxes = [torch.tensor([[10.0, 8.0, 90.0, 100.0], [40.5, 20.25, 150.0, 110.0]]), torch.tensor([[30.0, 30.0, 112.0, 95.0]])]
    uv = rng.uniform(-1.1, 1.1, size=(3, 41, 2)).astype(np.float32)
    mask = rng.random((3, 41)) < 0.4
    mask[1] = False                                  # a non-symmetric object in the same batch: all-zero prior planes
    planes = np.stack([prior_oracle.make_prior_kp_input(uv[i], mask[i], (64, 64)) for i in range(3)])
    m = _model(sd, 2, 3)
    dense = m(img.cuda(), [b.cuda() for b in boxes], [torch.from_numpy(planes[:2]).cuda(), torch.from_numpy(planes[2:]).cuda()])
    fused = m(img.cuda(), [b.cuda() for b in boxes], prior_uv=[(uv[:2], mask[:2]), (uv[2:], mask[2:])])
    host = m(img, boxes, prior_uv=[(uv[:2], mask[:2]), (uv[2:], mask[2:])])          # host tensors: staged by the C ABI
    torch.cuda.synchronize()
    for k in ("prob_logits", "uv", "cov", "kp_mask", "argmax"):
        assert torch.equal(dense[k], fused[k]), k
        assert torch.equal(dense[k].cpu(), host[k]), k
    ref = net_oracle.pkpnet_forward(sd, img, boxes, [torch.from_numpy(planes[:2]), torch.from_numpy(planes[2:])], (64, 64))
    np.testing.assert_allclose(fused["uv"].cpu().numpy(), ref["uv"].numpy(), atol=5e-5)
    with pytest.raises(ValueError):
        m(img.cuda(), [b.cuda() for b in boxes], [torch.from_numpy(planes[:2]).cuda(), torch.from_numpy(planes[2:]).cuda()],
          prior_uv=[(uv[:2], mask[:2]), (uv[2:], mask[2:])])


def test_frame_pipeline_fused_call_vs_oracle_and_u8_frames(sd):
    """suo_frames / suo_frames_u8 (network -> gating -> PnP -> single-view BA in one call, rows a1-a8) against
    (1) the CPU frame oracle, (2) the same stages called one by one, (3) itself on the camera's u8 frames."""
    from oracle import frame_oracle
    from suo_slam_b200 import frames
    n_frames, n_obj = 3, 4
    imgs_u8, boxes, bi, mk, mm, kb, diam = [], [], [], [], [], [], []
    for f in range(n_frames):
        fr = synth.make_frame(40 + f, n_obj=n_obj, H=120, W=160)
        imgs_u8.append(fr["img"])
        bb = [o["bbox"] for o in fr["objs"]]
        boxes += bb; bi += [f] * n_obj
        mk += [o["model_kps"] for o in fr["objs"]]; mm += [o["model_kps_mask"] for o in fr["objs"]]; diam += [o["diameter"] for o in fr["objs"]]
        kb.append(frames.k_bbox_for(fr["K"], bb))
    imgs_u8 = np.stack(imgs_u8)
    imgs = np.ascontiguousarray(imgs_u8.transpose(0, 3, 1, 2).astype(np.float32) / 255)       # lib/object_slam.py:1092
    boxes, bi = np.stack(boxes).astype(np.float32), np.asarray(bi, np.int32)
    mk, mm, kb, diam = np.stack(mk), np.stack(mm), np.concatenate(kb), np.asarray(diam, np.float64)
    m = _model(sd, 2, 3, max_crops=16)
    pipe = frames.FramePipeline(m, kp_var_thresh=0.5, bbox_thresh=0.95, seed=2)
    a = pipe.run(imgs, boxes, bi, mk, mm, kb, diam)
    b = pipe.run(imgs_u8, boxes, bi, mk, mm, kb, diam)
    for k in a:                                             # u8 frames: the same crops, hence the same everything
        assert np.array_equal(a[k], b[k]), k
    # stage by stage on the GPU
    out = m(torch.from_numpy(imgs), [torch.from_numpy(boxes[bi == f]) for f in range(n_frames)])
    assert np.array_equal(out["uv"].numpy(), a["uv"]) and np.array_equal(out["cov"].numpy(), a["cov"])
    st = frames.solve_keypoints(m.context(), a["uv"], a["cov"], out["kp_mask"].numpy(), bi, mk, mm, kb, diam, kp_var_thresh=0.5,
                                bbox_thresh=0.95, seed=2)
    for k in ("T_pnp", "T_ba", "kp_used", "ba_inliers"):
        assert np.array_equal(st[k], a[k]), k
    # the CPU oracle (FP32 torch network + FP64 solvers): network outputs within the conv tolerance; on the oracle's
    # own keypoints the solver stages are compared in test_gpu_geom.py, here the gating decisions must agree wherever
    # they are not within the network tolerance of a threshold
    ref = frame_oracle.run_frames(sd, imgs, boxes, bi, mk, mm, kb, diam, input_res=(64, 64), kp_var_thresh=0.5, bbox_thresh=0.95, seed=2)
    np.testing.assert_allclose(a["uv"], ref["uv"], atol=5e-5)
    np.testing.assert_allclose(a["cov"], ref["cov"], atol=5e-5)
    std = np.sqrt(np.abs(ref["cov"][:, :, [0, 1], [0, 1]]))
    near = (np.abs(ref["kp_mask"] - 0.3) < 1e-3) | (np.abs(np.abs(ref["uv"]).max(-1) - 0.95) < 1e-3) | (np.abs(std - 1.0).min(-1) < 1e-3)
    assert np.array_equal(a["kp_used"][~near], ref["kp_used"][~near])


def test_cta_pair_conv_kernel_equals_single_cta_kernels(sd):
    """The 3x3 convs as CTA pairs (csrc/conv_pair.cu, the default) vs one CTA per tile: identical network output."""
    rng = np.random.default_rng(22)
    img = torch.from_numpy(rng.random((1, 3, 240, 320), dtype=np.float32)).cuda()
    boxes = [torch.tensor([[10.0, 20.0, 200.0, 230.0], [100.0, 5.0, 310.0, 200.0], [50.0, 50.0, 120.0, 140.0]]).cuda()]
    m = _model(sd, 2, 3, res=256, max_crops=3)
    m.context().set_option(_lib.SUO_OPT_CONV_HALO, 0)      # (see above)
    outs = {}
    for pair in (1, 0):
        m.context().set_option(_lib.SUO_OPT_CONV_PAIR, pair)
        n0 = m.context().kernel_launches()
        o = m(img, boxes)
        torch.cuda.synchronize()
        outs[pair] = (o, m.context().kernel_launches() - n0)
    assert outs[1][1] == outs[0][1]
    for k in ("prob_logits", "uv", "cov", "kp_mask", "argmax"):
        assert torch.equal(outs[1][0][k], outs[0][0][k]), k


def test_forward_ragged_and_empty_box_lists(sd):
    """Ragged per-image box lists (one image without any box) and a call with no boxes at all: same crops, same order as the
    reference (torchvision roi_align over the list, lib/models/pkpnet.py:93), empty tensors when there is nothing to do."""
    rng = np.random.default_rng(23)
    img = torch.from_numpy(rng.random((3, 3, 96, 128), dtype=np.float32)).cuda()
    boxes = [torch.tensor([[4.0, 6.0, 90.0, 80.0], [30.0, 10.0, 120.0, 70.0]]).cuda(), torch.zeros((0, 4)).cuda(),
             torch.tensor([[10.0, 20.0, 60.0, 90.0]]).cuda()]
    m = _model(sd, 2, 3, res=64, max_crops=4)
    out = m(img, boxes)
    torch.cuda.synchronize()
    ref = net_oracle.pkpnet_forward(sd, img.cpu(), [b.cpu() for b in boxes], None, (64, 64))
    lr = ref["prob_logits"].numpy()
    assert out["prob_logits"].shape == lr.shape == (3, 41, 16, 16)
    assert np.abs(out["prob_logits"].cpu().numpy() - lr).max() < 5e-4 * max(1.0, np.abs(lr).max() / 10)
    assert np.abs(out["uv"].cpu().numpy() - ref["uv"].numpy()).max() < 5e-5
    empty = m(img, [torch.zeros((0, 4)).cuda()] * 3)
    assert empty["uv"].shape == (0, 41, 2) and empty["cov"].shape == (0, 41, 2, 2) and empty["prob_logits"].shape == (0, 41, 16, 16)
    assert empty["uv"].device == img.device


def test_frame_pipeline_ragged_frames(sd):
    """Frames with different numbers of detections, one frame with none (its BA graph is empty), and a batch without any
    detection.  A frame without detections must not disturb the others: the batch gives exactly the results of the same
    detections with that frame removed (the RANSAC stream is keyed by the detection's index in the batch, so the comparison
    keeps the order); an empty batch gives empty arrays."""
    from suo_slam_b200 import frames
    counts = [3, 0, 1, 4]
    per_frame, imgs_u8 = [], []
    for f, n in enumerate(counts):
        fr = synth.make_frame(60 + f, n_obj=max(n, 1), H=120, W=160)
        imgs_u8.append(fr["img"])
        objs = fr["objs"][:n]
        bb = [o["bbox"] for o in objs]
        per_frame.append(dict(boxes=np.asarray(bb, np.float32).reshape(-1, 4), mk=np.asarray([o["model_kps"] for o in objs], np.float64).reshape(-1, 41, 3),
                              mm=np.asarray([o["model_kps_mask"] for o in objs]).reshape(-1, 41), kb=frames.k_bbox_for(fr["K"], bb).reshape(-1, 3, 3),
                              diam=np.asarray([o["diameter"] for o in objs], np.float64)))
    imgs_u8 = np.stack(imgs_u8)
    cat = lambda k: np.concatenate([p[k] for p in per_frame])
    bi = np.concatenate([np.full(n, f, np.int32) for f, n in enumerate(counts)])
    m = _model(sd, 2, 3, max_crops=16)
    pipe = frames.FramePipeline(m, kp_var_thresh=0.5, bbox_thresh=0.95, seed=2)
    full = pipe.run(imgs_u8, cat("boxes"), bi, cat("mk"), cat("mm"), cat("kb"), cat("diam"))
    assert full["T_pnp"].shape == (sum(counts), 4, 4)
    keep = [f for f, n in enumerate(counts) if n]
    bi_compact = np.concatenate([np.full(counts[f], i, np.int32) for i, f in enumerate(keep)])
    compact = pipe.run(imgs_u8[keep], cat("boxes"), bi_compact, cat("mk"), cat("mm"), cat("kb"), cat("diam"))
    for k in ("uv", "cov", "kp_used", "T_pnp", "T_ba", "ba_inliers"):
        assert np.array_equal(compact[k], full[k]), k
    # the first frame alone (detection indices 0..2 in both runs): identical, i.e. frames do not leak into each other
    p = per_frame[0]
    alone = pipe.run(imgs_u8[:1], p["boxes"], np.zeros(counts[0], np.int32), p["mk"], p["mm"], p["kb"], p["diam"])
    for k in ("uv", "cov", "kp_used", "T_pnp", "T_ba", "ba_inliers"):
        assert np.array_equal(alone[k], full[k][:counts[0]]), k
    none = pipe.run(imgs_u8, np.zeros((0, 4), np.float32), np.zeros(0, np.int32), np.zeros((0, 41, 3)), np.zeros((0, 41), bool),
                    np.zeros((0, 3, 3)), np.zeros(0))
    assert none["T_pnp"].shape == (0, 4, 4) and none["kp_used"].shape == (0, 41) and none["kp_used"].dtype == bool


@pytest.mark.parametrize("grid_cap", [0, 3])
def test_halo_conv_kernel_network_output(sd, grid_cap, monkeypatch):
    """The 3x3 convs on the A-halo kernel (csrc/conv_halo.cu, opt-in): the network output agrees with the default kernels to FP32
    rounding (another accumulation order) and with the oracle within the usual tolerance; grid_cap = 3: many tiles per cluster."""
    if grid_cap:
        monkeypatch.setenv("SUO_GRID_CAP", str(grid_cap))
    rng = np.random.default_rng(24)
    img = torch.from_numpy(rng.random((1, 3, 240, 320), dtype=np.float32)).cuda()
    boxes = [torch.tensor([[10.0, 20.0, 200.0, 230.0], [100.0, 5.0, 310.0, 200.0], [50.0, 50.0, 120.0, 140.0]]).cuda()]
    m = _model(sd, 2, 3, res=256, max_crops=3)
    outs = {}
    for halo in (1, 0):
        m.context().set_option(_lib.SUO_OPT_CONV_HALO, halo)
        o = m(img, boxes)
        torch.cuda.synchronize()
        outs[halo] = {k: v.cpu().numpy() for k, v in o.items()}
    ref = net_oracle.pkpnet_forward(sd, img.cpu(), [b.cpu() for b in boxes], None, (256, 256))
    lr = ref["prob_logits"].numpy()
    tol = 5e-4 * max(1.0, np.abs(lr).max() / 10)
    assert np.abs(outs[1]["prob_logits"] - lr).max() < tol
    assert np.abs(outs[1]["prob_logits"] - outs[0]["prob_logits"]).max() < tol
    assert np.abs(outs[1]["uv"] - ref["uv"].numpy()).max() < 5e-5
