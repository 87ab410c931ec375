"""CPU: the net oracle (oracle/net_oracle.py) against fixtures produced by the UNMODIFIED
reference modules (oracle/gen_golden_net.py), and live against the reference when mounted."""
import os

import numpy as np
import pytest
import torch

from oracle import net_oracle, ref_shims
from suo_slam_b200 import arch, synth


@pytest.fixture(scope="module")
def sd():
    return synth.make_synthetic_state_dict(seed=0, peaky=4.0)


@pytest.mark.parametrize("name", ["net_small", "net_prior"])
def test_oracle_reproduces_reference_golden(golden_dir, sd, name):
    g = np.load(f"{golden_dir}/{name}.npz")
    prior = None if g["prior"].size == 0 else [torch.from_numpy(g["prior"])]
    out = net_oracle.pkpnet_forward(sd, torch.from_numpy(g["img"]), [torch.from_numpy(g["boxes"])], prior, (64, 64))
    # same torch ops in the same order as the reference => essentially bit-equal; allow thread-count reassociation
    np.testing.assert_allclose(out["prob_logits"].numpy(), g["logits"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(out["uv"].numpy(), g["uv"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out["cov"].numpy(), g["cov"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out["kp_mask"].numpy(), g["kp_mask"], rtol=0, atol=1e-5)
    assert np.array_equal(out["argmax"].numpy(), g["logits"].reshape(3, 41, -1).argmax(-1))


def test_reduce_golden(golden_dir):
    g = np.load(f"{golden_dir}/reduce.npz")
    out = net_oracle.heatmap_reduce(torch.from_numpy(g["logits"]))
    np.testing.assert_allclose(out["uv"].numpy(), g["uv"], atol=1e-6)
    np.testing.assert_allclose(out["cov"].numpy(), g["cov"], atol=1e-6)
    np.testing.assert_allclose(out["prob"].numpy(), g["prob"], atol=1e-7)
    # the flat map: uv = 0 and cov = variance of the pixel-centre grid (1 - 1/H^2)/3
    assert abs(out["uv"][1, 5]).max() < 1e-6
    np.testing.assert_allclose(out["cov"][1, 5].numpy(), np.eye(2) * (1 - 1 / 32 ** 2) / 3, atol=1e-5)


def test_grid_is_transposed():
    """uv[...,0] varies along ROWS, uv[...,1] = -(column coordinate)."""
    raw = torch.full((1, 1, 8, 8), -50.0)
    raw[0, 0, 1, 6] = 50.0   # row 1, col 6
    out = net_oracle.heatmap_reduce(raw)
    r = (np.arange(0.5, 8, 1) / 4 - 1)
    np.testing.assert_allclose(out["uv"][0, 0].numpy(), [r[1], -r[6]], atol=1e-6)
    assert int(out["argmax"][0, 0]) == 1 * 8 + 6


def test_kbbox_golden(golden_dir):
    g = np.load(f"{golden_dir}/kbbox.npz")
    for K, bb, ref in zip(g["K"], g["bbox"], g["K_bbox"]):
        np.testing.assert_allclose(synth.fix_K_for_bbox_ndc(K, bb), ref, rtol=1e-14, atol=1e-12)


def test_kbbox_float32_golden(golden_dir):
    """utils.fix_K_for_bbox_ndc on float32 bboxes with fractional coordinates (what the reference's loader produces): the host function and
    the SLAM oracle's copy reproduce the reference's result — including its float32 `x2 - x1` and `2.0 / w` — to the last bits, and exactly
    after the float32 staging of lib/object_slam.py:1082-1086."""
    from oracle import slam_frame_oracle as sfo
    from suo_slam_b200 import frames
    g = np.load(f"{golden_dir}/kbbox_f32.npz")
    for bb, raw, f32 in zip(g["bbox"], g["K_bbox"], g["K_bbox_f32"]):
        np.testing.assert_allclose(synth.fix_K_for_bbox_ndc(g["K"], bb), raw, rtol=2e-15, atol=1e-13)
        np.testing.assert_allclose(sfo.fix_K_for_bbox_ndc(g["K"], bb, f32=False), raw, rtol=2e-15, atol=1e-13)
        assert np.array_equal(sfo.fix_K_for_bbox_ndc(g["K"], bb), f32.astype(np.float64))
    assert np.array_equal(frames.k_bbox_for(g["K"], g["bbox"]), g["K_bbox_f32"].astype(np.float64))
    # the float64 arithmetic the restatements used before differs in the scale factors (this is what the float32 path is for)
    d = [abs(2.0 / (float(bb[2]) - float(bb[0])) - raw[0, 0] / g["K"][0, 0]) for bb, raw in zip(g["bbox"], g["K_bbox"])]
    assert max(d) > 1e-10


def test_state_dict_spec_counts():
    spec = arch.state_dict_spec()
    assert len(spec) == 1274                         # reference PkpNet().state_dict() (SURVEY §2.2 probe)
    n_param = sum(int(np.prod(s)) for k, s in spec if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert n_param == 12_726_732


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not mounted")
def test_oracle_vs_live_reference(sd):
    ref = ref_shims.import_reference_pkpnet()
    net = ref.PkpNet(input_res=(64, 64))
    net.load_state_dict(sd, strict=True)
    net.eval()
    rng = np.random.default_rng(3)
    img = torch.from_numpy(rng.random((2, 3, 96, 128), dtype=np.float32))
    boxes = [torch.tensor([[5.0, 6.0, 70.0, 80.0]]), torch.tensor([[20.0, 10.0, 100.0, 90.0], [0.0, 0.0, 127.0, 95.0]])]
    with torch.no_grad():
        a = net(img, boxes, None)
    b = net_oracle.pkpnet_forward(sd, img, boxes, None, (64, 64))
    for k in ("uv", "cov", "kp_mask", "prob_logits"):
        np.testing.assert_allclose(b[k].numpy(), a[k].numpy(), atol=2e-4 if k == "prob_logits" else 1e-5)


def test_prior_oracle_reproduces_reference_golden(golden_dir):
    """Row f2: the restated make_prior_kp_input against planes produced by the unmodified reference function."""
    from oracle import prior_oracle
    g = np.load(os.path.join(golden_dir, "prior.npz"))
    for name in "abc":
        uv, mask, shape, ndc = g[name + "_uv"], g[name + "_mask"], tuple(g[name + "_shape"]), bool(g[name + "_ndc"])
        got = np.stack([prior_oracle.make_prior_kp_input(uv[i], mask[i], shape, ndc=ndc) for i in range(len(uv))])
        assert np.array_equal(got, g[name + "_planes"]), name


def test_prior_stamp_table_is_the_reference_stamp():
    """The committed device table (csrc/prior_gauss_table.inc) == the stamp OpenCV produces for the reference call."""
    from oracle import prior_oracle
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "suo_slam_b200", "csrc", "prior_gauss_table.inc")
    words = []
    for line in open(path):
        if line.startswith("0x"):
            words += [int(w.strip().rstrip("u"), 16) for w in line.strip().rstrip(",").split(",")]
    q = np.array(words, np.uint32).view(np.float32).reshape(46, 46)
    g = prior_oracle.gaussian_2d(91)
    idx = np.minimum(np.arange(91), 90 - np.arange(91))
    assert np.array_equal(q[np.ix_(idx, idx)], g)
    assert g[45, 45] == 1.0 and g[0, 45] > g[1, 45]        # BORDER_REFLECT_101 doubles the outermost tap


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not mounted")
def test_prior_oracle_vs_live_reference():
    from oracle import prior_oracle
    ref_utils = ref_shims.import_reference_utils()
    rng = np.random.default_rng(1)
    for dt in (np.float32, np.float64):
        uv = rng.uniform(-1.2, 1.2, size=(41, 2)).astype(dt)
        m = rng.random(41) < 0.7
        assert np.array_equal(prior_oracle.make_prior_kp_input(uv, m, (256, 256)), ref_utils.make_prior_kp_input(uv, m, (256, 256)))
