"""state_dict -> packed weight blob + execution program for libsuo_b200.

Replaces ``PkpNet.load_state_dict(checkpoint['model'])`` + the nn.Module tree of the
reference (lib/object_slam.py:92-97, lib/models/hg.py:60-119, lib/models/hg.py:6-58,
lib/models/layers/Residual.py:3-35).  The reference keeps BatchNorm layers as separate
modules; here every BN that directly follows a conv is folded into that conv (in
float64, then rounded once to fp32) and the pre-activation BN+ReLU in front of each
bottleneck becomes a per-input-channel affine prologue of its first 1x1 conv.

The blob also carries the *program*: a buffer table and an op list that the C++
executor (csrc/api.cu) replays; the network topology therefore lives only here.
Blob layout (little endian): BlobHeader (16 x i32) | n_bufs x {div, C, kind} |
n_ops x OpDesc (16 x i32) | float pool.
"""
from __future__ import annotations

import numpy as np

from . import arch

MAGIC = 0x574F5553  # 'SUOW'
OP_CONV, OP_MAXPOOL, OP_UPADD = 0, 1, 2
CONV_1x1, CONV_3x3, CONV_STEM7 = 0, 1, 2
BN_EPS = 1e-5


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64) if hasattr(t, "detach") else np.asarray(t, dtype=np.float64)


class _Builder:
    def __init__(self, sd, num_kp):
        self.sd = sd
        self.num_kp = num_kp
        self.bufs = []      # (div, C)
        self.ops = []       # 16 ints each
        self.pool = []      # list of float32 arrays
        self.pool_len = 0

    # ---- pool / buffers -----------------------------------------------------------
    def put(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32).ravel()
        pad = (-self.pool_len) % 4          # keep every array 16-byte aligned (float4 loads)
        if pad:
            self.pool.append(np.zeros(pad, np.float32))
            self.pool_len += pad
        off = self.pool_len
        self.pool.append(a)
        self.pool_len += a.size
        return off

    def buf(self, div, C, kind=0):
        """kind 1: consumed only by plain (no prologue) 1x1/3x3 convs -> the executor may keep it as two FP16 planes
        (hi, lo') that the consumer loads by TMA (fp16x3 tensor-core path)."""
        self.bufs.append((div, C, kind))
        return len(self.bufs) - 1

    # ---- folding ------------------------------------------------------------------
    def bn_affine(self, p):
        g, b = _np(self.sd[p + ".weight"]), _np(self.sd[p + ".bias"])
        m, v = _np(self.sd[p + ".running_mean"]), _np(self.sd[p + ".running_var"])
        s = g / np.sqrt(v + BN_EPS)
        return s, b - m * s

    def conv_wb(self, p, post_bn=None):
        w, b = _np(self.sd[p + ".weight"]), _np(self.sd[p + ".bias"])
        if post_bn is not None:
            s, t = self.bn_affine(post_bn)
            w = w * s[:, None, None, None]
            b = b * s + t
        return w, b

    # ---- ops ----------------------------------------------------------------------
    def conv(self, in_buf, out_buf, w, b, mode, *, pre=None, relu=0, res=-1, variant=2, out_nchw=0,
             cin_store=None, cout_store=None):
        """w: [Cout, Cin, kh, kw] float64 (already folded)."""
        Cout, Cin = w.shape[0], w.shape[1]
        cin_store = cin_store or self.bufs[in_buf][1]
        Cout_pad = -(-Cout // 64) * 64
        if mode == CONV_1x1:
            K, cpr = cin_store, 0
            wk = np.zeros((Cout_pad, K))
            wk[:Cout, :Cin] = w[:, :, 0, 0]
        elif mode == CONV_3x3:
            K, cpr = 9 * cin_store, 0
            wk = np.zeros((Cout_pad, 9, cin_store))
            wk[:Cout, :, :Cin] = w.transpose(0, 2, 3, 1).reshape(Cout, 9, Cin)
            wk = wk.reshape(Cout_pad, K)
        else:
            if 7 * cin_store <= 32:
                # RGB-only layout (4 floats / pixel): a kernel row is 28 floats = ONE 32-float chunk, so a 64-wide FP16 chunk
                # holds two kernel rows; an eighth, all-zero row pads K to 256 (was 7 x 64 = 448: 43 % less producer + MMA work)
                cpr, nrow = 1, 8
            else:
                cpr, nrow = 2 * -(-7 * cin_store // 64), 7    # kernel-row stride: a multiple of 64 floats (TF32 and FP16 chunk widths)
            K = nrow * cpr * 32
            wk = np.zeros((Cout_pad, nrow, cpr * 32))
            row = np.zeros((Cout, 7, 7, cin_store))
            row[:, :, :, :Cin] = w.transpose(0, 2, 3, 1)
            wk[:Cout, :7, :7 * cin_store] = row.reshape(Cout, 7, 7 * cin_store)
            wk = wk.reshape(Cout_pad, K)
        assert K % 64 == 0, (K, mode)
        bk = np.zeros(Cout_pad)
        bk[:Cout] = b
        w_off, b_off = self.put(wk), self.put(bk)
        pre_off = -1
        if pre is not None:
            s, t = pre
            assert len(s) == cin_store
            pre_off = self.put(np.concatenate([s, t]))
        n_store = cout_store if cout_store is not None else Cout
        self.ops.append([OP_CONV, variant, in_buf, out_buf, res, mode, cin_store, n_store, Cout_pad, K, cpr,
                         relu, out_nchw, w_off, b_off, pre_off])
        return out_buf

    def residual(self, p, x, cin, cout, div):
        """layers/Residual.py:20-35 as 3 (or 4) fused conv ops."""
        mid = cout // 2
        w1, b1 = self.conv_wb(p + ".conv1", p + ".bn1")
        t1 = self.conv(x, self.buf(div, mid, 1), w1, b1, CONV_1x1, pre=self.bn_affine(p + ".bn"), relu=1)
        w2, b2 = self.conv_wb(p + ".conv2", p + ".bn2")
        t2 = self.conv(t1, self.buf(div, mid, 1), w2, b2, CONV_3x3, relu=1)
        skip = x
        if cin != cout:
            w4, b4 = self.conv_wb(p + ".conv4")
            skip = self.conv(x, self.buf(div, cout), w4, b4, CONV_1x1)
        w3, b3 = self.conv_wb(p + ".conv3")
        return self.conv(t2, self.buf(div, cout), w3, b3, CONV_1x1, res=skip)

    def maxpool(self, x, div, C):
        out = self.buf(div * 2, C)
        self.ops.append([OP_MAXPOOL, 2, x, out, -1] + [0] * 11)
        return out

    def upadd(self, up1, low, div, C):
        out = self.buf(div, C)
        self.ops.append([OP_UPADD, 2, up1, out, low] + [0] * 11)
        return out

    def hourglass(self, p, x, n, div):
        """hg.py:37-58."""
        F = arch.N_FEATS
        up1 = x
        for j in range(arch.N_MODULES):
            up1 = self.residual(f"{p}.up1_.{j}", up1, F, F, div)
        low1 = self.maxpool(x, div, F)
        for j in range(arch.N_MODULES):
            low1 = self.residual(f"{p}.low1_.{j}", low1, F, F, div * 2)
        if n > 1:
            low2 = self.hourglass(p + ".low2", low1, n - 1, div * 2)
        else:
            low2 = low1
            for j in range(arch.N_MODULES):
                low2 = self.residual(f"{p}.low2_.{j}", low2, F, F, div * 2)
        low3 = low2
        for j in range(arch.N_MODULES):
            low3 = self.residual(f"{p}.low3_.{j}", low3, F, F, div * 2)
        return self.upadd(up1, low3, div, F)


def pack_state_dict(sd, num_kp: int = arch.NUM_KP) -> bytes:
    """Fold, reorder and serialise a reference-format state dict (see module docstring)."""
    missing = [k for k, _ in arch.state_dict_spec(num_kp) if k not in sd]
    if missing:
        raise KeyError(f"state dict is missing {len(missing)} keys, e.g. {missing[:3]}")
    B = _Builder(sd, num_kp)
    F = arch.N_FEATS
    p = "backbone"
    # stem: 7x7/2 conv + bn1 + relu (hg.py:96-98); two input layouts, one output
    w, b = B.conv_wb(p + ".conv1_", p + ".bn1")
    in4, in48 = B.buf(1, 4), B.buf(1, 48)
    x = B.buf(2, 64)
    B.conv(in4, x, w[:, :3], b, CONV_STEM7, relu=1, variant=0)       # priors == None: prior planes are zero
    B.conv(in48, x, w, b, CONV_STEM7, relu=1, variant=1)
    x = B.residual(p + ".r1", x, 64, 128, 2)
    x = B.maxpool(x, 2, 128)
    x = B.residual(p + ".r4", x, 128, 128, 4)
    x = B.residual(p + ".r5", x, 128, F, 4)
    logits = -1
    for i in range(arch.N_STACK):
        ll = B.hourglass(f"{p}.hourglass.{i}", x, arch.HG_DEPTH, 4)
        for j in range(arch.N_MODULES):
            ll = B.residual(f"{p}.Residual.{i * arch.N_MODULES + j}", ll, F, F, 4)
        wl, bl = B.conv_wb(f"{p}.lin_.{i}.0", f"{p}.lin_.{i}.1")
        ll = B.conv(ll, B.buf(4, F, 1), wl, bl, CONV_1x1, relu=1)
        wt, bt = B.conv_wb(f"{p}.tmpOut.{i}")
        if i < arch.N_STACK - 1:
            # x = x + ll_(ll) + tmpOut_(tmpOut(ll)) (hg.py:113-117).  The intermediate heat-maps of this stack are not returned (hg.py:119 keeps
            # out[-1] only) and nothing non-linear sits between tmpOut and tmpOut_, so the two 1x1 convs and ll_ are ONE 1x1 conv on ll:
            #   W = W_ll + W_tmpOut_ W_tmpOut,  b = b_ll + W_tmpOut_ b_tmpOut + b_tmpOut_      (folded in float64, rounded once like BN)
            # — one launch and one pass over the 64x64x256 tensor instead of three (the reference's association differs by FP32 rounding only).
            wll, bll = B.conv_wb(f"{p}.ll_.{i}")
            wto, bto = B.conv_wb(f"{p}.tmpOut_.{i}")
            w_eff = wll[:, :, 0, 0] + wto[:, :, 0, 0] @ wt[:, :, 0, 0]
            b_eff = bll + wto[:, :, 0, 0] @ bt + bto
            x = B.conv(ll, B.buf(4, F), w_eff[:, :, None, None], b_eff, CONV_1x1, res=x)
        else:
            logits = B.conv(ll, B.buf(4, num_kp), wt, bt, CONV_1x1, out_nchw=1)   # NCHW heat-map logits
    cls_w = B.put(_np(sd["classifier.2.weight"]))
    cls_b = B.put(_np(sd["classifier.2.bias"]))
    pool = np.concatenate(B.pool).astype(np.float32)
    header = np.zeros(16, np.int32)
    header[:12] = [MAGIC, 2, num_kp, len(B.bufs), len(B.ops), pool.size, in4, in48, logits, cls_w, cls_b, 4]
    return b"".join([header.tobytes(), np.asarray(B.bufs, np.int32).tobytes(),
                     np.asarray(B.ops, np.int32).tobytes(), pool.tobytes()])


def program_summary(blob: bytes):
    """(n_bufs, n_ops, n_convs, pool floats) of a packed blob — for tests / docs."""
    h = np.frombuffer(blob[:64], np.int32)
    ops = np.frombuffer(blob[64 + 12 * h[3]: 64 + 12 * h[3] + 64 * h[4]], np.int32).reshape(-1, 16)
    return dict(n_bufs=int(h[3]), n_ops=int(h[4]), n_convs=int((ops[:, 0] == OP_CONV).sum()), n_floats=int(h[5]))
