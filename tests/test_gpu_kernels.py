"""GPU parity: each CUDA kernel, through the C ABI, against the CPU oracle on the same inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import net_oracle
from suo_slam_b200 import _lib, pkpnet, synth

pytestmark = pytest.mark.gpu


def _conv_ref(x_nhwc, w_ohwi, bias, ksize, stride, pre, residual, relu):
    x = torch.from_numpy(x_nhwc).double().permute(0, 3, 1, 2)
    if pre is not None:
        x = F.relu(x * torch.from_numpy(pre[0]).double()[None, :, None, None] + torch.from_numpy(pre[1]).double()[None, :, None, None])
    w = torch.from_numpy(w_ohwi).double().permute(0, 3, 1, 2)
    y = F.conv2d(x, w, None if bias is None else torch.from_numpy(bias).double(), stride=stride, padding=(ksize - 1) // 2)
    if relu:
        y = F.relu(y)
    y = y.permute(0, 2, 3, 1).numpy()
    if residual is not None:
        y = y + residual
    return y


CONV_CASES = [
    # B, H, W, Cin, Cout, ksize, stride, pre, residual, relu
    (2, 16, 16, 256, 128, 1, 1, True, False, True),     # conv1 of a bottleneck
    (2, 16, 16, 128, 128, 3, 1, False, False, True),    # conv2
    (2, 16, 16, 128, 256, 1, 1, False, True, False),    # conv3 + skip
    (3, 4, 4, 256, 256, 1, 1, False, True, False),      # M = 48 < one tile
    (1, 8, 8, 64, 64, 3, 1, False, False, True),        # r1.conv2
    (2, 32, 32, 4, 64, 7, 2, False, False, True),       # stem, RGB-only layout
    (1, 32, 32, 48, 64, 7, 2, False, False, True),      # stem, 44+4 channel layout
    (2, 16, 16, 256, 41, 1, 1, False, False, False),    # tmpOut (N = 41 -> padded tile)
    (2, 16, 16, 64, 256, 1, 1, False, True, False),     # tmpOut_ (K = 64)
    (3, 32, 32, 64, 64, 1, 1, True, False, True),       # r1.conv1: 64-column tiles, K = 64, prologue
    (3, 16, 16, 128, 64, 1, 1, True, False, True),      # r4.conv1: 64-column tiles, two K chunks
    (3, 32, 32, 64, 128, 1, 1, False, False, False),    # r1.conv4 (skip projection, FP32 in / FP32 out)
    (3, 1, 1, 256, 128, 1, 1, True, False, True),       # deepest hourglass level of a 64x64 crop: M = 3
    (3, 2, 2, 128, 256, 1, 1, False, True, False),      # M = 12
    (40, 32, 32, 256, 128, 1, 1, True, False, True),    # 320 tiles: several tiles per CTA (persistent loop, ring wrap)
    (40, 32, 32, 128, 256, 1, 1, False, True, False),   # 640 tiles, skip tensor streamed
    (24, 32, 32, 128, 128, 3, 1, False, False, True),   # 192 tiles of the 3x3
    (40, 32, 32, 128, 256, 1, 1, False, False, False),  # two column tiles, several tiles per CTA, no skip
    (40, 32, 32, 128, 128, 1, 1, False, True, False),   # one column tile, skip tensor
]


# tolerances are relative to max|ref|.  FP32 SIMT: 2e-6.  3xTF32: the tensor core's FP32 accumulate truncates,
# error grows ~1.4e-8 per accumulation step (K/8 steps): 8e-6 covers K = 2464 (stem).  Plain TF32: 2^-11 operands.
# FP16x3 (backend 2): 22-bit operands, error ~2^-22 per product plus the same accumulate truncation.
# backends 3/4/5: A operand by TMA from FP16 planes / output written as FP16 planes / both (plain 1x1 and 3x3 only).
@pytest.mark.parametrize("backend,passes,tol", [(0, 3, 2e-6), (1, 3, 8e-6), (1, 1, 3e-3), (2, 3, 8e-6), (3, 3, 8e-6), (4, 3, 8e-6), (5, 3, 8e-6)])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_engine_vs_fp64_reference(gpu_ctx, case, backend, passes, tol):
    B, H, W, Cin, Cout, ks, st, use_pre, use_res, relu = case
    if backend in (3, 5) and (use_pre or ks == 7 or Cin % 64):
        pytest.skip("TMA-fed A: plain 1x1 / 3x3 convs with Cin % 64 == 0 only")
    rng = np.random.default_rng(hash(case) % 2 ** 31)
    x = rng.normal(size=(B, H, W, Cin)).astype(np.float32)
    w = (rng.normal(size=(Cout, ks, ks, Cin)) / np.sqrt(ks * ks * Cin)).astype(np.float32)
    b = rng.normal(size=Cout).astype(np.float32)
    pre = (rng.uniform(0.5, 1.5, Cin).astype(np.float32), rng.normal(scale=0.1, size=Cin).astype(np.float32)) if use_pre else None
    Ho, Wo = (H // 2, W // 2) if st == 2 else (H, W)
    res = rng.normal(size=(B, Ho, Wo, Cout)).astype(np.float32) if use_res else None
    if Cout % 4:   # NHWC store needs Cout % 4 == 0: the network pads 41 -> 64 (zero weights); do the same here
        pad = (-Cout) % 4
        w = np.concatenate([w, np.zeros((pad,) + w.shape[1:], np.float32)])
        b = np.concatenate([b, np.zeros(pad, np.float32)])
        Cout += pad
    got = pkpnet.conv2d(gpu_ctx, x, w, b, ks, st, pre, res, relu, backend=backend, tf32_passes=passes)
    ref = _conv_ref(x, w, b, ks, st, pre, res, relu)
    scale = np.abs(ref).max()
    err = np.abs(got - ref).max() / scale
    assert err < tol, f"rel-to-max error {err:.3e} (backend {backend}, passes {passes})"


PAIR_CASES = [
    # B, H, W, Cin — 3x3, Cout = 128, ReLU, FP16-plane input and output (the bottleneck's conv2)
    (2, 16, 16, 128),     # 4 tiles = 2 pair tiles
    (24, 32, 32, 128),    # 192 tiles: several pair tiles per cluster (ring wrap, both accumulators recycled)
    (3, 64, 64, 128),     # tile = 2 image rows of 64
    (5, 8, 8, 128),       # 2.5 tiles: odd tile count, the last tile half outside the tensor
    (3, 4, 4, 128),       # M = 48: the second CTA of the only pair has no pixels at all
    (2, 16, 16, 64),      # one K chunk per tap
    (1, 128, 128, 128),   # tile = one image row
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_pair_kernel_equals_single_cta_kernel(gpu_ctx, case):
    """csrc/conv_pair.cu (two CTAs, one tcgen05.mma.cta_group::2 of M = 256, each CTA loads half of the weight rows) forms the
    same products in the same order as the single-CTA kernel: outputs must be IDENTICAL, and right against FP64."""
    B, H, W, Cin = case
    rng = np.random.default_rng(hash(case) % 2 ** 31)
    x = rng.normal(size=(B, H, W, Cin)).astype(np.float32)
    w = (rng.normal(size=(128, 3, 3, Cin)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.normal(size=128).astype(np.float32)
    n0 = gpu_ctx.kernel_launches()
    pair = pkpnet.conv2d(gpu_ctx, x, w, b, 3, 1, None, None, True, backend=6)
    single = pkpnet.conv2d(gpu_ctx, x, w, b, 3, 1, None, None, True, backend=5)
    assert gpu_ctx.kernel_launches() - n0 == 2
    assert np.array_equal(pair, single)
    ref = _conv_ref(x, w, b, 3, 1, None, None, True)
    assert np.abs(pair - ref).max() / np.abs(ref).max() < 8e-6


HALO_CASES = [(2, 16, 16, 128), (24, 32, 32, 128), (3, 64, 64, 128), (2, 16, 16, 64), (1, 64, 64, 64), (9, 16, 16, 128)]


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv_halo_kernel_vs_fp64_and_single_cta_kernel(gpu_ctx, case):
    """csrc/conv_halo.cu fetches the activations once per column shift (three x-shifted variants of rows + 2 image rows, the dy taps are
    descriptor offsets into the same shared-memory image): same products as the other 3x3 kernels in another accumulation order."""
    B, H, W, Cin = case
    rng = np.random.default_rng(hash(case) % 2 ** 31)
    x = rng.normal(size=(B, H, W, Cin)).astype(np.float32)
    w = (rng.normal(size=(128, 3, 3, Cin)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.normal(size=128).astype(np.float32)
    n0 = gpu_ctx.kernel_launches()
    halo = pkpnet.conv2d(gpu_ctx, x, w, b, 3, 1, None, None, True, backend=7)
    single = pkpnet.conv2d(gpu_ctx, x, w, b, 3, 1, None, None, True, backend=5)
    assert gpu_ctx.kernel_launches() - n0 == 2
    ref = _conv_ref(x, w, b, 3, 1, None, None, True)
    scale = np.abs(ref).max()
    assert np.abs(halo - ref).max() / scale < 8e-6
    assert np.abs(halo - single).max() / scale < 2e-6


def test_heatmap_reduce_vs_reference_golden(gpu_ctx, golden_dir):
    g = np.load(f"{golden_dir}/reduce.npz")
    out = pkpnet.heatmap_reduce(gpu_ctx, g["logits"])
    assert np.array_equal(out["argmax"], g["logits"].reshape(2, 41, -1).argmax(-1))      # bit-exact indices
    np.testing.assert_allclose(out["uv"], g["uv"], atol=1e-6)
    np.testing.assert_allclose(out["cov"], g["cov"], atol=1e-6)
    np.testing.assert_allclose(out["prob"], g["prob"], atol=1e-7, rtol=1e-5)


@pytest.mark.parametrize("shape", [(8, 41, 64, 64), (2, 41, 128, 128), (1, 3, 4, 4)])
def test_heatmap_reduce_vs_oracle_shapes(gpu_ctx, shape):
    rng = np.random.default_rng(shape[2])
    logits = rng.normal(scale=2.0, size=shape).astype(np.float32)
    K = shape[1]
    cw, cb = rng.normal(size=(K, K)).astype(np.float32), rng.normal(size=K).astype(np.float32)
    ctx = gpu_ctx if K == 41 else _lib.Context(0, 1, 64, K)
    out = pkpnet.heatmap_reduce(ctx, logits, cw, cb)
    ref = net_oracle.heatmap_reduce(torch.from_numpy(logits), torch.from_numpy(cw), torch.from_numpy(cb))
    assert np.array_equal(out["argmax"], ref["argmax"].numpy())
    np.testing.assert_allclose(out["uv"], ref["uv"].numpy(), atol=2e-6)
    np.testing.assert_allclose(out["cov"], ref["cov"].numpy(), atol=2e-6)
    np.testing.assert_allclose(out["kp_mask"], ref["kp_mask"].numpy(), atol=1e-5)
    np.testing.assert_allclose(out["kp_mask_logits"], ref["kp_mask_logits"].numpy(), atol=1e-4)


def test_heatmap_ties_pick_first_index(gpu_ctx):
    logits = np.zeros((1, 41, 8, 8), np.float32)
    logits[0, 3, 2, 5] = logits[0, 3, 6, 1] = 4.0
    out = pkpnet.heatmap_reduce(gpu_ctx, logits)
    assert out["argmax"][0, 3] == 2 * 8 + 5 and out["argmax"][0, 0] == 0


@pytest.mark.parametrize("with_prior", [False, True])
def test_crop_concat_vs_torchvision(gpu_ctx, with_prior):
    import torchvision
    rng = np.random.default_rng(9)
    H, W, R = 120, 160, 64
    img = rng.random((2, 3, H, W), dtype=np.float32)
    boxes = np.array([[10.0, 8.0, 90.0, 100.0], [40.5, 20.25, 150.0, 110.0], [100.0, 30.0, 112.0, 45.0],
                      [-20.0, -10.0, 60.0, 50.0], [100.0, 60.0, 200.0, 140.0], [5.0, 5.0, 5.5, 5.5]], np.float32)
    box_img = np.array([0, 0, 0, 1, 1, 1], np.int32)
    L = len(boxes)
    prior = rng.random((L, 41, R, R), dtype=np.float32) if with_prior else None
    oc = 48 if with_prior else 4
    out = np.zeros((L, R, R, oc), np.float32)
    gpu_ctx.check(_lib.lib().suo_crop_concat(gpu_ctx.handle, _lib.ptr(img), 2, H, W, _lib.ptr(boxes), _lib.ptr(box_img), L,
                                              _lib.ptr(prior), R, _lib.ptr(out), oc, 0, None))
    bl = [torch.from_numpy(boxes[box_img == i]) for i in range(2)]
    ref = torchvision.ops.roi_align(torch.from_numpy(img), bl, output_size=(R, R)).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(out[..., :3], ref, atol=2e-6)
    if with_prior:
        np.testing.assert_array_equal(out[..., 3:44], prior.transpose(0, 2, 3, 1))
        assert not out[..., 44:].any()
    else:
        assert not out[..., 3].any()


def test_render_priors_vs_reference_golden(gpu_ctx, golden_dir):
    """Row f2: prior planes rendered on the GPU == planes produced by the unmodified reference
    utils.make_prior_kp_input (tests/golden/prior.npz), bit for bit — .5 ties, clamping, NaN/inf, masked
    keypoints, windows clipped by every border, NDC and pixel input."""
    import os
    from suo_slam_b200 import utils
    g = np.load(os.path.join(golden_dir, "prior.npz"))
    for name in "abc":
        uv, mask, shape, ndc = g[name + "_uv"], g[name + "_mask"], tuple(int(v) for v in g[name + "_shape"]), bool(g[name + "_ndc"])
        got = utils.make_prior_kp_input_batch(uv, mask, shape, ndc=ndc, ctx=gpu_ctx)
        assert got.dtype == np.float32 and np.array_equal(got, g[name + "_planes"]), name
    one = utils.make_prior_kp_input(g["a_uv"][1], g["a_mask"][1], (256, 256), ctx=gpu_ctx)      # the reference's single-object form
    assert np.array_equal(one, g["a_planes"][1])


def test_render_priors_vs_oracle_full_size(gpu_ctx):
    """A whole frame's worth at the reference's size: 8 objects x 41 keypoints x 256 x 256."""
    from oracle import prior_oracle
    from suo_slam_b200 import utils
    rng = np.random.default_rng(3)
    uv = rng.uniform(-1.1, 1.1, size=(8, 41, 2)).astype(np.float32)
    mask = rng.random((8, 41)) < 0.5
    got = utils.make_prior_kp_input_batch(uv, mask, (256, 256), ctx=gpu_ctx)
    ref = np.stack([prior_oracle.make_prior_kp_input(uv[i], mask[i], (256, 256)) for i in range(8)])
    assert np.array_equal(got, ref)
    assert not got[~mask].any() and (got[mask].reshape(mask.sum(), -1).max(1) > 0.99).all()   # centre may sit on the clipped row 256
