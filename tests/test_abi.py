"""CPU: the C-ABI library builds, loads, and exports every symbol include/suo_b200.h declares;
host-side packing logic.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from suo_slam_b200 import _lib, arch, synth, weights


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(os.path.dirname(_lib.HERE), "include", "suo_b200.h")).read()
    declared = set(re.findall(r"\b(suo_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("suo_ctx")
    assert declared == set(_lib.exported_symbols())
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), name


def test_create_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        return
    try:
        _lib.Context(device=0)
    except _lib.SuoError as e:
        assert "no CPU path" in str(e)
    else:
        raise AssertionError("Context() must raise without a GPU: there is no CPU fallback")


def test_pack_state_dict_program():
    sd = synth.make_synthetic_state_dict(0)
    blob = weights.pack_state_dict(sd)
    info = weights.program_summary(blob)
    # 186 reference convs + the second (RGB-only) stem variant - 2: ll_, tmpOut_ and the first stack's tmpOut are folded into one 1x1 conv
    assert info["n_convs"] == 185
    # 9 max-pools (hg.py:16,71) and 8 up-sample+adds (hg.py:56-58) in two depth-4 hourglasses
    assert info["n_ops"] == 185 + 9 + 8
    h = np.frombuffer(blob[:64], np.int32)
    assert h[0] == weights.MAGIC and h[2] == arch.NUM_KP


def test_bn_folding_matches_unfused_math():
    """conv -> BN(eval) == folded conv, checked on one bottleneck with torch CPU."""
    import torch
    import torch.nn.functional as F
    sd = synth.make_synthetic_state_dict(1)
    B = weights._Builder(sd, 41)
    p = "backbone.r4"
    w, b = B.conv_wb(p + ".conv1", p + ".bn1")
    x = torch.randn(2, 128, 8, 8, dtype=torch.float64)
    ref = F.batch_norm(F.conv2d(x, sd[p + ".conv1.weight"].double(), sd[p + ".conv1.bias"].double()),
                       sd[p + ".bn1.running_mean"].double(), sd[p + ".bn1.running_var"].double(),
                       sd[p + ".bn1.weight"].double(), sd[p + ".bn1.bias"].double(), False, 0.0, 1e-5)
    got = F.conv2d(x, torch.from_numpy(w), torch.from_numpy(b))
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-12)


def test_pkpnet_rejects_bad_state_dict():
    import pytest
    from suo_slam_b200.pkpnet import PkpNet
    sd = synth.make_synthetic_state_dict(0)
    bad = dict(sd)
    bad.pop("backbone.r1.conv1.weight")
    with pytest.raises(RuntimeError):
        PkpNet().load_state_dict(bad)
    PkpNet().load_state_dict(sd)        # packs without a GPU


def test_checkpoint_converter_roundtrip(tmp_path):
    """Row f4: a checkpoint in the reference's format (train.py:173-181: {'model', 'epoch', 'args'}, keys prefixed
    with 'module.' when trained under DataParallelWrapper) -> packed weights file -> the same blob
    PkpNet.load_state_dict would have built."""
    import argparse
    import torch
    from suo_slam_b200 import checkpoint, synth, weights
    from suo_slam_b200.pkpnet import PkpNet
    sd = synth.make_synthetic_state_dict(seed=3)
    ck = tmp_path / "checkpoint-7.pth.tar"
    torch.save({"model": {"module." + k: v for k, v in sd.items()}, "epoch": 7, "best_result": 0.5,
                "args": argparse.Namespace(dataset="ycbv", lr=1e-3)}, ck)
    sd2, epoch, args = checkpoint.load_checkpoint(str(ck))
    assert epoch == 7 and args.dataset == "ycbv" and set(sd2) == set(sd)
    meta = checkpoint.convert(str(ck), str(tmp_path / "m.suo"))
    blob, meta2 = checkpoint.load_packed(str(tmp_path / "m.suo"))
    assert meta2 == meta and meta["epoch"] == 7 and meta["n_convs"] == 185
    assert blob == weights.pack_state_dict(sd)
    m = PkpNet().load_packed(blob)
    assert m._blob == blob
    with pytest.raises(ValueError):
        torch.save({"epoch": 1}, tmp_path / "bad.pth.tar")
        checkpoint.load_checkpoint(str(tmp_path / "bad.pth.tar"))
    with pytest.raises(RuntimeError):
        PkpNet().load_packed(b"\\0" * 128)


def test_option_constants_match_the_header():
    """suo_slam_b200/_lib.py mirrors the SUO_OPT_* enum of include/suo_b200.h by hand: keep them equal."""
    import re
    from suo_slam_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(_lib.HERE), "include", "suo_b200.h")).read()
    enum = dict(re.findall(r"(SUO_OPT_[A-Z0-9_]+)\s*=\s*(\d+)", hdr))
    assert len(enum) >= 10
    for name, val in enum.items():
        assert getattr(_lib, name) == int(val), name


@pytest.mark.parametrize("cin_store", [4, 48])
def test_stem_weight_layout_matches_the_gather_order(cin_store):
    """The packed stem weights [Cout_pad][K] and the implicit-GEMM gather order of csrc/conv_gather.cuh (K index = kernel row ky *
    (chunks_per_row * 32) + kx * Cin_store + ci, starting at input pixel (2 oy + ky - 3, 2 ox - 3)) reproduce the 7x7/2 pad-3 conv: emulate the
    gather in numpy and compare with torch.  RGB-only layout: one 32-float chunk per kernel row and an eighth all-zero row (K = 256)."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(cin_store)
    Cin = 3 if cin_store == 4 else 44
    w = rng.normal(size=(64, Cin, 7, 7))
    b = rng.normal(size=64)
    B = weights._Builder({}, 41)
    i0, o0 = B.buf(1, cin_store), B.buf(2, 64)
    B.conv(i0, o0, w, b, weights.CONV_STEM7, relu=1, variant=0)
    op = B.ops[0]
    Cout_pad, K, cpr, w_off, b_off = op[8], op[9], op[10], op[13], op[14]
    pool = np.concatenate(B.pool)
    wk = pool[w_off:w_off + Cout_pad * K].reshape(Cout_pad, K).astype(np.float64)
    assert (K, cpr) == ((256, 1) if cin_store == 4 else (7 * 12 * 32, 12))
    H = W = 16
    x = np.zeros((H, W, cin_store))
    x[:, :, :Cin] = rng.normal(size=(H, W, Cin))
    Ho, Wo = H // 2, W // 2
    A = np.zeros((Ho * Wo, K))
    for oy in range(Ho):
        for ox in range(Wo):
            for ky in range(K // (cpr * 32)):
                iy = 2 * oy + ky - 3
                for e in range(cpr * 32):                      # float index inside the kernel row window
                    pix, ci = divmod(e, cin_store)
                    ix = 2 * ox - 3 + pix
                    if pix < 7 and 0 <= iy < H and 0 <= ix < W:
                        A[oy * Wo + ox, ky * cpr * 32 + e] = x[iy, ix, ci]
    got = np.maximum(A @ wk[:64].T + pool[b_off:b_off + 64], 0).reshape(Ho, Wo, 64)
    ref = F.relu(F.conv2d(torch.from_numpy(x[:, :, :Cin]).permute(2, 0, 1)[None], torch.from_numpy(w), torch.from_numpy(b), stride=2, padding=3))
    np.testing.assert_allclose(got, ref[0].permute(1, 2, 0).numpy(), atol=2e-5)      # pool is float32


def test_dropin_puts_the_library_under_the_reference_module_names():
    """suo_slam_b200.dropin.install(): the three imports through which the reference reaches its hot path (lib/object_slam.py:9-10,14)
    resolve to this package — with the reference's own lib/ package untouched when it is mounted."""
    import importlib
    import sys
    from suo_slam_b200 import dropin, g2o as our_g2o, lambdatwist as our_lt
    from suo_slam_b200.pkpnet import PkpNet
    saved = {k: sys.modules.get(k) for k in ("g2o", "lambdatwist", "lib.models.pkpnet")}
    try:
        dropin.install()
        assert importlib.import_module("g2o") is our_g2o and importlib.import_module("lambdatwist") is our_lt
        for name in ("SparseOptimizer", "BlockSolverSE3", "LinearSolverDenseSE3", "LinearSolverCholmodSE3", "OptimizationAlgorithmLevenberg", "SE3Quat",
                     "VertexSE3Expmap", "EdgeSE3ProjectFromObject", "EdgeSE3ProjectFromFixedObject", "RobustKernelHuber"):   # lib/object_slam.py:706-831
            assert hasattr(our_g2o, name), name
        assert callable(our_lt.pnp)
        ref = "/root/reference"
        if os.path.isdir(os.path.join(ref, "lib", "models")):
            sys.path.insert(0, ref)
            try:
                for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
                    if k != "lib.models.pkpnet":
                        sys.modules.pop(k)
                mod = importlib.import_module("lib.models.pkpnet")          # what `from .models.pkpnet import PkpNet` resolves to
                assert mod.PkpNet is PkpNet
                hg = importlib.import_module("lib.models.hg")               # the rest of the reference package still imports from its own tree
                assert hg.__file__.startswith(ref)
            finally:
                sys.path.remove(ref)
                for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
                    sys.modules.pop(k)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
