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
    # 186 reference convs + the second (RGB-only) stem variant
    assert info["n_convs"] == 187
    # 9 max-pools (hg.py:16,71) and 8 up-sample+adds (hg.py:56-58) in two depth-4 hourglasses
    assert info["n_ops"] == 187 + 9 + 8
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
    assert meta2 == meta and meta["epoch"] == 7 and meta["n_convs"] == 187
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
