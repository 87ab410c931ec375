"""Drop-in for the reference's ``lib.models.pkpnet.PkpNet`` (lib/models/pkpnet.py:65-119).

Same constructor arguments, same ``load_state_dict`` key names, same
``model(images, boxes, prior_kp)`` call and the same output dict — but the forward
is one call into libsuo_b200 (crop + concat, hourglass on tcgen05 tensor cores,
heat-map reduction, classifier).  Tensor plumbing only on the Python side.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, arch, weights


class PkpNet:
    def __init__(self, input_res=(256, 256), calc_cov=True, max_crops: int = 64):
        assert input_res[0] == input_res[1], "Only support square images for now"   # pkpnet.py:23
        self.input_res = tuple(input_res)
        self.calc_cov = calc_cov
        self.num_kp = arch.NUM_KP
        self.max_crops = max_crops
        self.training = False
        self.return_prob = True          # the reference always returns "prob"; turn off to save 2x heat-map traffic
        # fp16x3 conv math is range-guarded (|activation| <= 6e4): forwards on CUDA tensors poll the device flag after the call
        # (one stream synchronisation — the reference's caller synchronises right after anyway, lib/object_slam.py:1100-1109);
        # set to True to skip the poll and call check_range() yourself at the next synchronisation point
        self.defer_range_check = False
        self._sd = None
        self._blob = None
        self._ctx = None
        self._device = None

    # ---- nn.Module-like surface used by lib/object_slam.py:92-99 ---------------------
    def load_state_dict(self, state_dict, strict: bool = True):
        spec = dict(arch.state_dict_spec(self.num_kp))
        missing = [k for k in spec if k not in state_dict]
        unexpected = [k for k in state_dict if k not in spec]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for PkpNet: missing {missing[:4]}..., "
                               f"unexpected {unexpected[:4]}...")
        for k, shp in spec.items():
            if k in state_dict and tuple(state_dict[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(state_dict[k].shape)} vs {shp}")
        self._sd = {k: v.detach().cpu().clone() for k, v in state_dict.items()}
        self._blob = weights.pack_state_dict(self._sd, self.num_kp)
        self._ctx = None
        return self

    def load_packed(self, blob: bytes):
        """Weights already folded and packed (suo_slam_b200.checkpoint.convert / weights.pack_state_dict)."""
        h = np.frombuffer(blob[:64], np.int32)
        if h[0] != weights.MAGIC or h[2] != self.num_kp:
            raise RuntimeError("not a packed PkpNet weight blob for this keypoint vocabulary")
        self._sd, self._blob, self._ctx = None, bytes(blob), None
        return self

    def state_dict(self):
        return dict(self._sd) if self._sd is not None else {}

    def cuda(self, device=None):
        self._device = torch.cuda.current_device() if device is None else torch.device(device).index or 0
        return self

    def to(self, device):
        d = torch.device(device)
        if d.type != "cuda":
            raise RuntimeError("suo_slam_b200.PkpNet runs on a B200 only (no CPU path)")
        return self.cuda(d)

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise RuntimeError("suo_slam_b200.PkpNet is inference-only (training is out of scope, SURVEY.md §2 #13)")
        return self

    def parameters(self):
        return iter(())

    # ---- context ---------------------------------------------------------------------
    def context(self) -> _lib.Context:
        if self._ctx is None:
            if self._blob is None:
                raise _lib.SuoError("PkpNet.forward before load_state_dict")
            dev = self._device if self._device is not None else 0
            self._ctx = _lib.Context(device=dev, max_crops=self.max_crops, crop_res=self.input_res[0], num_kp=self.num_kp)
            self._ctx.load_weights(self._blob)
            self._device = dev
        return self._ctx

    # ---- forward ---------------------------------------------------------------------
    def forward(self, images, boxes, prior_kp=None, prior_uv=None):
        """images [B,3,H,W] f32; boxes: list (len B) of [L_i,4] xyxy; prior_kp: list of
        [L_i,41,R,R] or None.  Returns the reference's dict of torch tensors on images.device.

        Extension: instead of the dense ``prior_kp`` planes the caller may pass
        ``prior_uv`` = list (len B) of ``(uv [L_i,41,2] f32 NDC, mask [L_i,41] bool)`` — what ObjectSLAM has
        in hand before it calls utils.make_prior_kp_input (lib/object_slam.py:510-514); the planes are then
        stamped into the network input on the device (bit-identical to the reference's planes)."""
        assert type(boxes) == list and len(boxes) == images.shape[0]          # pkpnet.py:91
        ctx = self.context()
        dev = images.device
        on_dev = dev.type == "cuda"
        if on_dev and dev.index != ctx.device:
            raise _lib.SuoError(f"images on {dev} but the model context is on cuda:{ctx.device}")
        f32 = dict(dtype=torch.float32, device=dev)
        images = images.to(torch.float32).contiguous()
        box_img = torch.cat([torch.full((len(b),), i, dtype=torch.int32) for i, b in enumerate(boxes)]).to(dev)
        boxes_t = torch.cat([b.reshape(-1, 4) for b in boxes]).to(**f32).contiguous()
        L = boxes_t.shape[0]
        priors = None
        if prior_kp is not None:
            priors = torch.cat(list(prior_kp)).to(**f32).contiguous()
            assert priors.shape == (L, self.num_kp, *self.input_res), priors.shape
        puv = pmask = None
        if prior_uv is not None:
            if prior_kp is not None:
                raise ValueError("give prior_kp (planes) or prior_uv (keypoints), not both")
            assert len(prior_uv) == len(boxes)
            puv = torch.cat([torch.as_tensor(u).reshape(-1, self.num_kp, 2) for u, _ in prior_uv]).to(**f32).contiguous()
            pmask = torch.cat([torch.as_tensor(m).reshape(-1, self.num_kp) for _, m in prior_uv]).to(dtype=torch.uint8, device=dev).contiguous()
            assert puv.shape[0] == L and pmask.shape[0] == L
        K, HM = self.num_kp, self.input_res[0] // 4
        if L == 0:
            # no boxes at all: the reference's roi_align returns an empty batch and every output is an empty tensor
            # (lib/models/pkpnet.py:93-119 on K = 0 crops); nothing to launch
            out = {"uv": torch.empty((0, K, 2), **f32), "prob_logits": torch.empty((0, K, HM, HM), **f32),
                   "kp_mask_logits": torch.empty((0, K), **f32), "kp_mask": torch.empty((0, K), **f32),
                   "argmax": torch.empty((0, K), dtype=torch.int32, device=dev)}
            if self.calc_cov:
                out["cov"] = torch.empty((0, K, 2, 2), **f32)
            if self.return_prob:
                out["prob"] = torch.empty((0, K, HM, HM), **f32)
            return out
        out = {
            "uv": torch.empty((L, K, 2), **f32),
            "prob_logits": torch.empty((L, K, HM, HM), **f32),
            "kp_mask_logits": torch.empty((L, K), **f32),
            "kp_mask": torch.empty((L, K), **f32),
            "argmax": torch.empty((L, K), dtype=torch.int32, device=dev),
        }
        cov = torch.empty((L, K, 2, 2), **f32) if self.calc_cov else None
        prob = torch.empty((L, K, HM, HM), **f32) if self.return_prob else None
        stream = torch.cuda.current_stream(dev).cuda_stream if on_dev else None
        B, _, H, W = images.shape
        outs = (_lib.ptr(out["uv"]), _lib.ptr(cov), _lib.ptr(out["prob_logits"]), _lib.ptr(prob),
                _lib.ptr(out["kp_mask_logits"]), _lib.ptr(out["kp_mask"]), _lib.ptr(out["argmax"]), 1 if on_dev else 0, stream)
        if puv is not None:
            ctx.check(_lib.lib().suo_forward_kp_priors(ctx.handle, _lib.ptr(images), B, H, W, _lib.ptr(boxes_t), _lib.ptr(box_img), L,
                                                       _lib.ptr(puv), _lib.ptr(pmask), *outs))
        else:
            ctx.check(_lib.lib().suo_forward(ctx.handle, _lib.ptr(images), B, H, W, _lib.ptr(boxes_t), _lib.ptr(box_img), L,
                                             _lib.ptr(priors), *outs))
        if on_dev and not self.defer_range_check:
            self.check_range()
        if cov is not None:
            out["cov"] = cov
        if prob is not None:
            out["prob"] = prob
        return out

    __call__ = forward

    def check_range(self):
        """Raises SuoError (SUO_E_RANGE) when an activation left the FP16 range of the fp16x3 conv math since the last check:
        the outputs of those forwards are invalid; switch the context to SUO_OPT_CONV_MATH = 0 (tf32x3)."""
        ctx = self.context()
        ctx.check(_lib.lib().suo_check_range(ctx.handle))


def heatmap_reduce(ctx: _lib.Context, logits, cls_w=None, cls_b=None, want_prob=True):
    """Stand-alone a3+a4 on host numpy arrays (tests)."""
    logits = np.ascontiguousarray(logits, np.float32)
    B, K, H, W = logits.shape
    out = dict(uv=np.zeros((B, K, 2), np.float32), cov=np.zeros((B, K, 2, 2), np.float32),
               argmax=np.zeros((B, K), np.int32), kp_mask=np.zeros((B, K), np.float32),
               kp_mask_logits=np.zeros((B, K), np.float32))
    prob = np.zeros_like(logits) if want_prob else None
    cw = None if cls_w is None else np.ascontiguousarray(cls_w, np.float32)
    cb = None if cls_b is None else np.ascontiguousarray(cls_b, np.float32)
    ctx.check(_lib.lib().suo_heatmap_reduce(ctx.handle, _lib.ptr(logits), B, K, H, W, _lib.ptr(cw), _lib.ptr(cb),
                                             _lib.ptr(out["uv"]), _lib.ptr(out["cov"]), _lib.ptr(prob),
                                             _lib.ptr(out["kp_mask_logits"]), _lib.ptr(out["kp_mask"]),
                                             _lib.ptr(out["argmax"]), 0, None))
    if want_prob:
        out["prob"] = prob
    return out


def conv2d(ctx: _lib.Context, x_nhwc, w_ohwi, bias=None, ksize=1, stride=1, pre=None, residual=None, relu=False,
           backend=1, tf32_passes=3):
    """One conv through the engine (tests/bench hook). x [B,H,W,Cin], w [Cout,kh,kw,Cin] numpy."""
    x = np.ascontiguousarray(x_nhwc, np.float32)
    w = np.ascontiguousarray(w_ohwi, np.float32)
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    Ho, Wo = (H // 2, W // 2) if stride == 2 else (H, W)
    out = np.zeros((B, Ho, Wo, Cout), np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    ps = pt = None
    if pre is not None:
        ps, pt = np.ascontiguousarray(pre[0], np.float32), np.ascontiguousarray(pre[1], np.float32)
    r = None if residual is None else np.ascontiguousarray(residual, np.float32)
    ctx.check(_lib.lib().suo_conv2d(ctx.handle, _lib.ptr(x), B, H, W, Cin, _lib.ptr(w), _lib.ptr(b), Cout, ksize, stride,
                                    _lib.ptr(ps), _lib.ptr(pt), _lib.ptr(r), int(relu), _lib.ptr(out), backend,
                                    tf32_passes, 0, None))
    return out
