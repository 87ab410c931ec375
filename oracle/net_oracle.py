"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference keypoint
network forward (rows a1-a4 of SURVEY.md §8).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this; the product
path (suo_slam_b200/) never does.

It is a *functional* restatement over a plain state dict (no nn.Module tree):
each function cites the reference lines it follows.  Arithmetic is torch CPU
fp32 (the reference's own library: MKL-DNN convs, torchvision roi_align), i.e.
the same third-party ops the reference calls, composed by our own code.

PARITY PIN: tests/test_oracle_net.py checks this file against fixtures in
tests/golden/net_*.npz that were produced by the UNMODIFIED reference modules
(lib/models/pkpnet.py, hg.py, layers/Residual.py imported from /root/reference
by oracle/gen_golden_net.py) and, when /root/reference is present, live
against the reference module itself.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
import torchvision

BN_EPS = 1e-5  # nn.BatchNorm2d default used at hg.py:68, Residual.py:8-14


def _bn(x, sd, p):
    # eval-mode BatchNorm2d (running stats), Residual.py:22 / hg.py:97
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _conv(x, sd, p, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def residual(x, sd, p):
    """Pre-activation bottleneck, layers/Residual.py:20-35."""
    out = F.relu(_bn(x, sd, p + ".bn"))
    out = _conv(out, sd, p + ".conv1")
    out = F.relu(_bn(out, sd, p + ".bn1"))
    out = _conv(out, sd, p + ".conv2", padding=1)
    out = F.relu(_bn(out, sd, p + ".bn2"))
    out = _conv(out, sd, p + ".conv3")
    skip = x
    if (p + ".conv4.weight") in sd:
        skip = _conv(x, sd, p + ".conv4")
    return out + skip


def hourglass(x, sd, p, n, n_modules=2):
    """hg.py:37-58."""
    up1 = x
    for j in range(n_modules):
        up1 = residual(up1, sd, f"{p}.up1_.{j}")
    low1 = F.max_pool2d(x, 2, 2)
    for j in range(n_modules):
        low1 = residual(low1, sd, f"{p}.low1_.{j}")
    if n > 1:
        low2 = hourglass(low1, sd, p + ".low2", n - 1, n_modules)
    else:
        low2 = low1
        for j in range(n_modules):
            low2 = residual(low2, sd, f"{p}.low2_.{j}")
    low3 = low2
    for j in range(n_modules):
        low3 = residual(low3, sd, f"{p}.low3_.{j}")
    up2 = F.interpolate(low3, scale_factor=2)  # nearest, hg.py:56
    return up1 + up2


def backbone(x, sd, n_stack=2, n_modules=2, depth=4):
    """HourglassNet.forward, hg.py:95-119 (returns the last stack's heat-maps)."""
    p = "backbone"
    x = _conv(x, sd, p + ".conv1_", stride=2, padding=3)
    x = F.relu(_bn(x, sd, p + ".bn1"))
    x = residual(x, sd, p + ".r1")
    x = F.max_pool2d(x, 2, 2)
    x = residual(x, sd, p + ".r4")
    x = residual(x, sd, p + ".r5")
    out = None
    for i in range(n_stack):
        ll = hourglass(x, sd, f"{p}.hourglass.{i}", depth, n_modules)
        for j in range(n_modules):
            ll = residual(ll, sd, f"{p}.Residual.{i * n_modules + j}")
        ll = F.relu(_bn(_conv(ll, sd, f"{p}.lin_.{i}.0"), sd, f"{p}.lin_.{i}.1"))
        out = _conv(ll, sd, f"{p}.tmpOut.{i}")
        if i < n_stack - 1:
            x = x + _conv(ll, sd, f"{p}.ll_.{i}") + _conv(out, sd, f"{p}.tmpOut_.{i}")
    return out


def heatmap_reduce(raw, classifier_w=None, classifier_b=None):
    """spatial_softmax + post_process_kp(calc_sigma=True) + classifier head,
    pkpnet.py:13-63,74-78,106-118.  Grid is TRANSPOSED exactly as in the
    reference's mesh_grid (pkpnet.py:19-26, torch.meshgrid default 'ij'):
    xx[h,w] = r[h], yy[h,w] = -r[w], r[i] = (i+0.5)/(H/2) - 1.
    Adds the builder-defined hard argmax: flat index h*W+w of
    the max logit, first occurrence on ties."""
    B, K, H, W = raw.shape
    assert H == W
    prob = F.softmax(raw.reshape(B, K, H * W), dim=-1).reshape(B, K, H, W)
    r = (torch.arange(0.5, H, 1) / (H / 2) - 1).to(torch.float32)
    xx = r[:, None].expand(H, W)
    yy = (-r)[None, :].expand(H, W)
    sx = torch.sum(prob * xx, [2, 3])
    sy = torch.sum(prob * yy, [2, 3])
    uv = torch.stack([sx, sy], -1)
    res = torch.stack([xx, yy], -1)[None, None] - uv.reshape(B, K, 1, 1, 2)
    cov = torch.sum(prob[..., None, None] * (res[..., :, None] * res[..., None, :]), [2, 3])
    ret = {"uv": uv, "cov": cov, "prob_logits": raw, "prob": prob,
           "argmax": torch.argmax(raw.reshape(B, K, H * W), dim=-1).to(torch.int32)}
    if classifier_w is not None:
        pooled = raw.mean(3).mean(2)
        logits = F.linear(F.relu(pooled), classifier_w, classifier_b)  # Dropout is identity in eval
        ret["kp_mask_logits"] = logits
        ret["kp_mask"] = torch.sigmoid(logits)
    return ret


def crop_and_concat(images, boxes, prior_kp, input_res, num_kp):
    """pkpnet.py:91-101: roi_align(defaults) then concat prior planes."""
    crops = torchvision.ops.roi_align(images, boxes, output_size=tuple(input_res))
    if prior_kp is None:
        prior = torch.zeros((crops.shape[0], num_kp, crops.shape[2], crops.shape[3]))
    else:
        prior = torch.cat(prior_kp)
    return torch.cat([crops, prior], 1)


@torch.no_grad()
def pkpnet_forward(sd, images, boxes, prior_kp=None, input_res=(256, 256)):
    """PkpNet.forward (eval mode), pkpnet.py:80-119."""
    num_kp = sd["classifier.2.bias"].shape[0]
    x = crop_and_concat(images, boxes, prior_kp, input_res, num_kp)
    raw = backbone(x, sd)
    return heatmap_reduce(raw, sd["classifier.2.weight"], sd["classifier.2.bias"])


def gate_keypoints(uv, cov, kp_mask, model_mask, bbox_thresh=0.9, kp_var_thresh=0.2):
    """Keypoint gating, lib/object_slam.py:1100-1115 (numpy in / numpy out)."""
    import numpy as np
    m = (kp_mask > 0.3) & model_mask
    m = m & (np.min(uv, -1) > -bbox_thresh) & (np.max(uv, -1) < bbox_thresh)
    std = np.sqrt(cov[..., [0, 1], [0, 1]])
    m = m & np.all(std < 2 * kp_var_thresh, axis=-1)
    return m
