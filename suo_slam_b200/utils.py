"""Host-side mirror of the two reference helpers that feed the hot path (reference lib/utils/utils.py), same names
and argument meaning, so a caller written against ``lib.utils.utils`` can switch the import:

* ``make_prior_kp_input(kp_uv, kp_uv_mask, img_shape, ndc=True)`` (utils.py:398-411) — rendered on the GPU by
  ``suo_render_priors`` (csrc/prior.cu); bit-identical planes.  ``make_prior_kp_input_batch`` does a whole
  frame's objects in one launch.
* ``fix_K_for_bbox_ndc(K, bbox)`` (utils.py:416-429) — nine FP64 numbers per object: plain numpy.
"""
from __future__ import annotations

import numpy as np

from . import _lib, runtime
from .synth import fix_K_for_bbox_ndc  # noqa: F401  (same function, kept in one place)


def make_prior_kp_input_batch(kp_uv, kp_uv_mask, img_shape, ndc=True, ctx=None):
    """kp_uv [L,N,2], kp_uv_mask [L,N] -> [L,N,height,width] float32."""
    ctx = ctx or runtime.get_context()
    uv = np.ascontiguousarray(kp_uv, np.float32)
    assert uv.ndim == 3 and uv.shape[2] >= 2
    uv = np.ascontiguousarray(uv[:, :, :2])
    mask = np.ascontiguousarray(np.asarray(kp_uv_mask).astype(bool), np.uint8)
    L, N = mask.shape
    h, w = int(img_shape[0]), int(img_shape[1])
    out = np.empty((L, N, h, w), np.float32)
    ctx.check(_lib.lib().suo_render_priors(ctx.handle, _lib.ptr(uv), _lib.ptr(mask), L, N, h, w, int(bool(ndc)), _lib.ptr(out), 0, None))
    return out


def make_prior_kp_input(kp_uv, kp_uv_mask, img_shape, ndc=True, ctx=None):
    """Drop-in for utils.make_prior_kp_input: kp_uv [N,2(+)], kp_uv_mask [N] -> [N,height,width] float32.
    The pixel arithmetic is FP32 (the reference gets FP32 too for ObjectSLAM's float32 prior_uv_full,
    lib/object_slam.py:510; a float64 input differs only on exact .5 ties after FP32 rounding)."""
    return make_prior_kp_input_batch(np.asarray(kp_uv)[None], np.asarray(kp_uv_mask)[None], img_shape, ndc, ctx)[0]
