"""Developer probe: one large 3x3 layer through the CTA-pair kernel (backend 6) and the single-CTA kernel (backend 5).
Run under ncu to compare the two kernels on the same input:
    ncu --set full --clock-control none -k regex:conv -o gpurun_out/pair_probe python tools/pair_probe.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
from suo_slam_b200 import _lib, pkpnet

B, H, W, Cin = int(os.environ.get("PROBE_B", "64")), 64, 64, 128
rng = np.random.default_rng(0)
x = rng.normal(size=(B, H, W, Cin)).astype(np.float32)
w = (rng.normal(size=(128, 3, 3, Cin)) / np.sqrt(9 * Cin)).astype(np.float32)
b = rng.normal(size=128).astype(np.float32)
ctx = _lib.Context(device=0, max_crops=8, crop_res=64, num_kp=41)
outs = {}
backends = tuple(int(v) for v in os.environ.get("PROBE_BACKENDS", "6,5").split(","))     # 7 = A-halo, 6 = CTA pair, 5 = single CTA
for rep in range(int(os.environ.get("PROBE_REPS", "2"))):
    for backend in backends:
        outs[backend] = pkpnet.conv2d(ctx, x, w, b, 3, 1, None, None, True, backend=backend)
ref = outs[backends[-1]]
for k in backends[:-1]:
    print(f"backend {k} vs {backends[-1]}: identical {np.array_equal(outs[k], ref)}, max rel diff {np.abs(outs[k] - ref).max() / np.abs(ref).max():.2e}")
ctx.close()
