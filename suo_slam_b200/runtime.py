"""Process-wide default context (one per GPU) for the geometry entry points."""
from __future__ import annotations

import os

from . import _lib

_ctxs = {}


def default_device() -> int:
    return int(os.environ.get("LOCAL_RANK", "0")) if "LOCAL_RANK" in os.environ else 0


def get_context(device: int | None = None) -> _lib.Context:
    d = default_device() if device is None else device
    if d not in _ctxs:
        _ctxs[d] = _lib.Context(device=d, max_crops=1, crop_res=256, num_kp=41)
    return _ctxs[d]


def current_stream_ptr(device: int):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
