"""Developer probe: ba_kernel / pnp_kernel time on EMPTY inputs (nothing gated) next to real inputs, CUDA events.
usage (GPU box): python tools/ba_empty_probe.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from suo_slam_b200 import _lib, frames, synth  # noqa: E402

L, K, F = 256, 41, 32
ctx = _lib.Context(device=0, max_crops=1, crop_res=256, num_kp=K)
lib, p = _lib.lib(), _lib.ptr
dev = torch.device("cuda", 0)
uv, cov, mk, mm, kb, bi = [], [], [], [], [], []
for f in range(F):
    fr = synth.make_frame(100 + f, n_obj=8)
    for o in fr["objs"]:
        uv.append(o["uv_meas"]); cov.append(o["cov"]); mk.append(o["model_kps"]); mm.append(o["model_kps_mask"]); bi.append(f)
    kb.append(frames.k_bbox_for(fr["K"], [o["bbox"] for o in fr["objs"]]))
t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
d = dict(uv=t(np.stack(uv), np.float32), cov=t(np.stack(cov), np.float32), bi=t(bi, np.int32), mk=t(np.stack(mk), np.float64), mm=t(np.stack(mm), np.uint8),
         kb=t(np.concatenate(kb), np.float64), diam=t(np.full(L, 150.0), np.float64))
o = dict(Tp=torch.zeros((L, 16), dtype=torch.float64, device=dev), Tb=torch.zeros((L, 12), dtype=torch.float64, device=dev),
         u=torch.zeros((L, K), dtype=torch.uint8, device=dev), b=torch.zeros((L, K), dtype=torch.uint8, device=dev))
s = torch.cuda.current_stream().cuda_stream
for name, mask_val in (("real keypoints (every object solved)", 0.9), ("empty (kp_mask = 0: nothing gated)", 0.0)):
    km = torch.full((L, K), mask_val, device=dev)
    for run_ba in (0, 1):
        def call():
            ctx.check(lib.suo_solve_keypoints(ctx.handle, p(d["uv"]), p(d["cov"]), p(km), p(d["bi"]), F, L, p(d["mk"]), p(d["mm"]), p(d["kb"]), p(d["diam"]), 0.2, 0.9, 0, run_ba,
                                              p(o["Tp"]), p(o["Tb"]), p(o["u"]), p(o["b"]), 1, s))
        call(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record(); torch.cuda.synchronize()
        print(f"{name:45s} run_ba={run_ba}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us per call (256 objects, 32 frames)")
