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
for rep in range(2):
    for backend in (6, 5):
        outs[backend] = pkpnet.conv2d(ctx, x, w, b, 3, 1, None, None, True, backend=backend)
print("identical:", np.array_equal(outs[5], outs[6]))
ctx.close()
