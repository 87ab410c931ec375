"""Single-view frame pipeline (the path BASELINE config 2 times): the three downward calls
ObjectSLAM makes per frame — model(...) (lib/object_slam.py:1099), pnp(...) per object (:1144)
and optimize() (:443-451) — plus the NumPy glue between them (:1100-1165), executed for a BATCH
of independent frames by one ``suo_frames`` call with everything resident on the GPU."""
from __future__ import annotations

import numpy as np

from . import _lib, synth
from .pkpnet import PkpNet


def k_bbox_for(K, bboxes):
    """utils.fix_K_for_bbox_ndc per box, rounded through float32 like the reference's
    ``K_bbox_np`` (float32 array, lib/object_slam.py:1082,1086) before .astype(float64) (:1140)."""
    out = np.zeros((len(bboxes), 3, 3), dtype=np.float32)
    for i, bb in enumerate(bboxes):
        out[i] = synth.fix_K_for_bbox_ndc(K, np.asarray(bb, np.float32))       # (the boxes of the C ABI are float32, as the reference's are)
    return out.astype(np.float64)


class FramePipeline:
    def __init__(self, model: PkpNet, kp_var_thresh: float = 0.2, bbox_thresh: float = 0.9, seed: int = 0):
        self.model = model
        self.kp_var_thresh, self.bbox_thresh, self.seed = kp_var_thresh, bbox_thresh, seed   # evaluate.py:58-76 (YCBV)

    def run(self, images, boxes, box_img, model_kps, model_mask, K_bbox, diameter, priors=None, run_ba=True):
        """Host arrays in (numpy or pinned torch CPU tensors), host numpy out.
        images [n_img,3,H,W] f32 in [0,1] (what model(...) takes) or [n_img,H,W,3] uint8 (what process_view receives,
        lib/object_slam.py:327-328; converted per tap on the GPU); boxes [L,4]; box_img [L] sorted; model_kps [L,K,3] f64;
        model_mask [L,K] bool; K_bbox [L,3,3] f64; diameter [L]."""
        ctx = self.model.context()
        u8 = str(images.dtype).endswith("uint8")
        if u8:
            n_img, H, W, _ = images.shape
        else:
            n_img, _, H, W = images.shape
        L, K = model_mask.shape
        c = lambda a, dt: a if (hasattr(a, "data_ptr") and not isinstance(a, np.ndarray)) else np.ascontiguousarray(a, dtype=dt)
        images = c(images, np.uint8 if u8 else np.float32)
        boxes, box_img = c(boxes, np.float32), c(box_img, np.int32)
        model_kps, K_bbox, diameter = c(model_kps, np.float64), c(K_bbox, np.float64), c(diameter, np.float64)
        model_mask = np.ascontiguousarray(model_mask, dtype=np.uint8)
        pri = None if priors is None else c(priors, np.float32)
        out = dict(T_pnp=np.zeros((L, 4, 4)), T_ba=np.zeros((L, 3, 4)), kp_used=np.zeros((L, K), np.uint8),
                   ba_inliers=np.zeros((L, K), np.uint8), uv=np.zeros((L, K, 2), np.float32),
                   cov=np.zeros((L, K, 2, 2), np.float32))
        if L == 0:                                  # frames without a single detection: nothing to run (process_view returns early)
            out["kp_used"], out["ba_inliers"] = out["kp_used"].astype(bool), out["ba_inliers"].astype(bool)
            return out
        entry = _lib.lib().suo_frames_u8 if u8 else _lib.lib().suo_frames
        ctx.check(entry(
            ctx.handle, _lib.ptr(images), n_img, H, W, _lib.ptr(boxes), _lib.ptr(box_img), L, _lib.ptr(pri),
            _lib.ptr(model_kps), _lib.ptr(model_mask), _lib.ptr(K_bbox), _lib.ptr(diameter),
            float(self.kp_var_thresh), float(self.bbox_thresh), int(self.seed), int(run_ba),
            _lib.ptr(out["T_pnp"]), _lib.ptr(out["T_ba"]), _lib.ptr(out["kp_used"]), _lib.ptr(out["ba_inliers"]),
            _lib.ptr(out["uv"]), _lib.ptr(out["cov"]), 0, None))
        out["kp_used"] = out["kp_used"].astype(bool)
        out["ba_inliers"] = out["ba_inliers"].astype(bool)
        return out

    def stream(self, batches, run_ba=True):
        """Generator: ``for out in pipe.stream(batches)`` yields, in order, what ``run(**batch)`` returns for every batch of an
        iterable of dicts (keys: images [n_img,H,W,3] uint8, boxes, box_img, model_kps, model_mask, K_bbox, diameter; no priors).
        Double-buffered through ``suo_frames_u8_submit`` / ``suo_frames_wait``: the batch is packed into the slot's pinned staging
        tensors and submitted, and while it runs the frames of the next batch are already crossing PCIe; a result is yielded when
        its slot is needed again (two batches later) or the input ends."""
        import torch
        ctx, lib, p = self.model.context(), _lib.lib(), _lib.ptr
        stream = torch.cuda.current_stream().cuda_stream
        slots = [dict(pending=False, bufs={}) for _ in range(2)]

        def pinned(sl, name, shape, dtype):
            t = sl["bufs"].get(name)
            n = int(np.prod(shape))
            if t is None or t.dtype != dtype or t.numel() < n:
                t = torch.empty(max(n, 1), dtype=dtype).pin_memory()
                sl["bufs"][name] = t
            return t[:n].view(*shape)

        def finish(sl):
            ctx.check(lib.suo_frames_wait(ctx.handle, sl["slot"]))
            sl["pending"] = False
            o, (L, K) = sl["out"], sl["LK"]
            return dict(T_pnp=o["T_pnp"].numpy().reshape(L, 4, 4).copy(), T_ba=o["T_ba"].numpy().reshape(L, 3, 4).copy(),
                        kp_used=o["used"].numpy().astype(bool), ba_inliers=o["bain"].numpy().astype(bool),
                        uv=o["uv"].numpy().copy(), cov=o["cov"].numpy().reshape(L, K, 2, 2).copy())

        try:
            for i, b in enumerate(batches):
                sl = slots[i % 2]
                sl["slot"] = i % 2
                if sl["pending"]:
                    yield finish(sl)
                images = b["images"]
                if not str(images.dtype).endswith("uint8"):
                    raise ValueError("FramePipeline.stream takes uint8 [n_img,H,W,3] frames")
                n_img, H, W, _ = images.shape
                L, K = b["model_mask"].shape
                if L == 0:
                    raise ValueError("FramePipeline.stream: a batch without crops (use run())")
                ins = {}
                for name, a, dt, tdt in (("images", images, np.uint8, torch.uint8), ("boxes", b["boxes"], np.float32, torch.float32),
                                         ("box_img", b["box_img"], np.int32, torch.int32), ("model_kps", b["model_kps"], np.float64, torch.float64),
                                         ("model_mask", b["model_mask"], np.uint8, torch.uint8), ("K_bbox", b["K_bbox"], np.float64, torch.float64),
                                         ("diameter", b["diameter"], np.float64, torch.float64)):
                    a = np.ascontiguousarray(a.numpy() if hasattr(a, "data_ptr") and not isinstance(a, np.ndarray) else a, dtype=dt)
                    t = pinned(sl, name, a.shape, tdt)
                    t.numpy()[...] = a
                    ins[name] = t
                o = dict(T_pnp=pinned(sl, "T_pnp", (L, 16), torch.float64), T_ba=pinned(sl, "T_ba", (L, 12), torch.float64),
                         used=pinned(sl, "used", (L, K), torch.uint8), bain=pinned(sl, "bain", (L, K), torch.uint8),
                         uv=pinned(sl, "uv", (L, K, 2), torch.float32), cov=pinned(sl, "cov", (L, K, 4), torch.float32))
                if not run_ba:
                    o["T_ba"].zero_(); o["bain"].zero_()
                sl["out"], sl["LK"] = o, (L, K)
                ctx.check(lib.suo_frames_u8_submit(
                    ctx.handle, sl["slot"], p(ins["images"]), n_img, H, W, p(ins["boxes"]), p(ins["box_img"]), L, p(ins["model_kps"]),
                    p(ins["model_mask"]), p(ins["K_bbox"]), p(ins["diameter"]), float(self.kp_var_thresh), float(self.bbox_thresh),
                    int(self.seed), int(run_ba), p(o["T_pnp"]), p(o["T_ba"]), p(o["used"]), p(o["bain"]), p(o["uv"]), p(o["cov"]),
                    None, 0, stream))
                sl["pending"] = True
                last = i
            if any(s_["pending"] for s_ in slots):
                for sl in (slots[(last + 1) % 2], slots[last % 2]):      # oldest first
                    if sl["pending"]:
                        yield finish(sl)
        finally:
            for sl in slots:                                              # a consumer that stops early must not leave a slot pending
                if sl["pending"]:
                    lib.suo_frames_wait(ctx.handle, sl["slot"])
                    sl["pending"] = False


def solve_keypoints(ctx, uv, cov, kp_mask, box_img, model_kps, model_mask, K_bbox, diameter, kp_var_thresh=0.2,
                    bbox_thresh=0.9, seed=0, run_ba=True):
    """Rows a4'..a8 on given keypoints (host numpy in/out): gating -> PnP -> single-view BA."""
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    uv, cov, kp_mask = c(uv, np.float32), c(cov, np.float32), c(kp_mask, np.float32)
    box_img = c(box_img, np.int32)
    L, K = kp_mask.shape
    n_img = int(box_img.max()) + 1 if L else 0
    model_kps, K_bbox, diameter = c(model_kps, np.float64), c(K_bbox, np.float64), c(diameter, np.float64)
    model_mask = c(model_mask, np.uint8)
    out = dict(T_pnp=np.zeros((L, 4, 4)), T_ba=np.zeros((L, 3, 4)), kp_used=np.zeros((L, K), np.uint8),
               ba_inliers=np.zeros((L, K), np.uint8))
    if L == 0:
        out["kp_used"], out["ba_inliers"] = out["kp_used"].astype(bool), out["ba_inliers"].astype(bool)
        return out
    ctx.check(_lib.lib().suo_solve_keypoints(
        ctx.handle, _lib.ptr(uv), _lib.ptr(cov), _lib.ptr(kp_mask), _lib.ptr(box_img), n_img, L, _lib.ptr(model_kps),
        _lib.ptr(model_mask), _lib.ptr(K_bbox), _lib.ptr(diameter), float(kp_var_thresh), float(bbox_thresh), int(seed),
        int(run_ba), _lib.ptr(out["T_pnp"]), _lib.ptr(out["T_ba"]), _lib.ptr(out["kp_used"]), _lib.ptr(out["ba_inliers"]),
        0, None))
    out["kp_used"] = out["kp_used"].astype(bool)
    out["ba_inliers"] = out["ba_inliers"].astype(bool)
    return out
