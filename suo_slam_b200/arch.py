"""Architecture table of the keypoint network, restated as data.

The reference builds the network out of nn.Modules (lib/models/hg.py:60-119,
lib/models/hg.py:6-58, lib/models/layers/Residual.py:3-35,
lib/models/pkpnet.py:65-78).  Here the same topology is a flat *program* of
conv ops that the weight packer (weights.py), the CUDA executor
(csrc/net_exec.cu) and the CPU oracle (oracle/net_oracle.py) all walk.  The
state-dict key names are the reference's so a reference
checkpoint's ``checkpoint['model']`` loads unchanged.
"""
from __future__ import annotations

NUM_KP = 41                 # lib/labeling/kp_config.py:82-94 (fixed global vocabulary)
N_STACK = 2                 # hg.py:61 nStack
N_MODULES = 2               # hg.py:61 nModules
N_FEATS = 256               # hg.py:61 nFeats
HG_DEPTH = 4                # hg.py:79 Hourglass(4, ...)


def residual_keys(prefix: str, cin: int, cout: int):
    """(key, shape) of one pre-activation bottleneck (Residual.py:4-18)."""
    mid = cout // 2
    out = []

    def bn(name, c):
        for s in ("weight", "bias", "running_mean", "running_var"):
            out.append((f"{prefix}.{name}.{s}", (c,)))
        out.append((f"{prefix}.{name}.num_batches_tracked", ()))

    def conv(name, co, ci, k):
        out.append((f"{prefix}.{name}.weight", (co, ci, k, k)))
        out.append((f"{prefix}.{name}.bias", (co,)))

    bn("bn", cin)
    conv("conv1", mid, cin, 1)
    bn("bn1", mid)
    conv("conv2", mid, mid, 3)
    bn("bn2", mid)
    conv("conv3", cout, mid, 1)
    if cin != cout:
        conv("conv4", cout, cin, 1)
    return out


def hourglass_residual_prefixes(prefix: str, n: int):
    """Residual prefixes of Hourglass(n) in nn.Module registration order
    (hg.py:6-35: low2 | low2_, up1_, low1_, low3_)."""
    out = []
    if n > 1:
        out += hourglass_residual_prefixes(f"{prefix}.low2", n - 1)
    else:
        out += [f"{prefix}.low2_.{j}" for j in range(N_MODULES)]
    out += [f"{prefix}.up1_.{j}" for j in range(N_MODULES)]
    out += [f"{prefix}.low1_.{j}" for j in range(N_MODULES)]
    out += [f"{prefix}.low3_.{j}" for j in range(N_MODULES)]
    return out


def state_dict_spec(num_kp: int = NUM_KP):
    """Ordered (key, shape) list identical to reference PkpNet().state_dict()."""
    cin = 3 + num_kp
    spec = [("backbone.conv1_.weight", (64, cin, 7, 7)), ("backbone.conv1_.bias", (64,))]
    for s in ("weight", "bias", "running_mean", "running_var"):
        spec.append((f"backbone.bn1.{s}", (64,)))
    spec.append(("backbone.bn1.num_batches_tracked", ()))
    spec += residual_keys("backbone.r1", 64, 128)
    spec += residual_keys("backbone.r4", 128, 128)
    spec += residual_keys("backbone.r5", 128, N_FEATS)
    for i in range(N_STACK):
        for p in hourglass_residual_prefixes(f"backbone.hourglass.{i}", HG_DEPTH):
            spec += residual_keys(p, N_FEATS, N_FEATS)
    for j in range(N_STACK * N_MODULES):
        spec += residual_keys(f"backbone.Residual.{j}", N_FEATS, N_FEATS)
    for i in range(N_STACK):
        spec.append((f"backbone.lin_.{i}.0.weight", (N_FEATS, N_FEATS, 1, 1)))
        spec.append((f"backbone.lin_.{i}.0.bias", (N_FEATS,)))
        for s in ("weight", "bias", "running_mean", "running_var"):
            spec.append((f"backbone.lin_.{i}.1.{s}", (N_FEATS,)))
        spec.append((f"backbone.lin_.{i}.1.num_batches_tracked", ()))
    for i in range(N_STACK):
        spec.append((f"backbone.tmpOut.{i}.weight", (num_kp, N_FEATS, 1, 1)))
        spec.append((f"backbone.tmpOut.{i}.bias", (num_kp,)))
    for i in range(N_STACK - 1):
        spec.append((f"backbone.ll_.{i}.weight", (N_FEATS, N_FEATS, 1, 1)))
        spec.append((f"backbone.ll_.{i}.bias", (N_FEATS,)))
    for i in range(N_STACK - 1):
        spec.append((f"backbone.tmpOut_.{i}.weight", (N_FEATS, num_kp, 1, 1)))
        spec.append((f"backbone.tmpOut_.{i}.bias", (N_FEATS,)))
    spec.append(("classifier.2.weight", (num_kp, num_kp)))
    spec.append(("classifier.2.bias", (num_kp,)))
    return spec
