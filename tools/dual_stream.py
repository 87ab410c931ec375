"""Developer experiment: S independent contexts (own buffers, own CUDA stream), each taking 1/S of the frames of a step,
persistent conv kernels capped at SUO_GRID_CAP CTAs — do HBM-bound and tensor-bound layers of different streams overlap?"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from suo_slam_b200 import _lib, synth  # noqa: E402
from suo_slam_b200.pkpnet import PkpNet  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
F_total = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = 6
F = F_total // S
L = F * bench.CROPS
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.lib()
ctxs = []
for k in range(S):
    m = PkpNet(input_res=(256, 256), max_crops=L)
    m.load_state_dict(synth.make_synthetic_state_dict(0, peaky=4.0))
    m.cuda(0)
    st = torch.cuda.Stream(dev)
    b = {kk: torch.from_numpy(v).to(dev) for kk, v in bench.make_batch(100 * k, F).items()}
    o = dict(T_pnp=torch.zeros((L, 16), dtype=torch.float64, device=dev), T_ba=torch.zeros((L, 12), dtype=torch.float64, device=dev),
             used=torch.zeros((L, 41), dtype=torch.uint8, device=dev), bain=torch.zeros((L, 41), dtype=torch.uint8, device=dev))
    ctxs.append((m, m.context(), st, b, o))


def step():
    p = _lib.ptr
    for m, ctx, st, b, o in ctxs:
        ctx.check(lib.suo_frames(ctx.handle, p(b["images"]), F, bench.H, bench.W, p(b["boxes"]), p(b["box_img"]), L, None, p(b["model_kps"]),
                                 p(b["model_mask"]), p(b["K_bbox"]), p(b["diameter"]), 0.2, 0.9, 0, 1,
                                 p(o["T_pnp"]), p(o["T_ba"]), p(o["used"]), p(o["bain"]), None, None, 1, st.cuda_stream))


for _ in range(3):
    step()
torch.cuda.synchronize()
main = torch.cuda.current_stream(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(main)
for _, _, st, _, _ in ctxs:
    st.wait_stream(main)
for _ in range(steps):
    step()
for _, _, st, _, _ in ctxs:
    main.wait_stream(st)
e1.record(main)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
for c in ctxs:
    c[1].close()
print(f"streams={S} grid_cap={os.environ.get('SUO_GRID_CAP', '-')} frames/step={F_total}: {ms:.2f} ms/step  {F_total / ms * 1e3:.1f} frames/s", flush=True)
