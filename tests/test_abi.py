"""CPU: the C-ABI library builds, loads, and exports every symbol include/suo_b200.h declares;
host-side packing logic.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np

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
