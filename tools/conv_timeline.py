"""Per-chunk / per-tile clock64 timeline of CTA 0 of the persistent tcgen05 conv kernel (developer tool).
Shapes: the three convs of a 256-channel bottleneck at 64x64 (Residual.py:20-35) as the network runs them."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.csv"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
os.environ["SUO_CONV_TIMELINE"] = out
from suo_slam_b200 import _lib, pkpnet  # noqa: E402

ctx = _lib.Context(0, 1, 64, 41)
rng = np.random.default_rng(0)
# (Cin, Cout, ksize, pre, residual, backend): backend 2 = fp16x3 register-fed, 3 = TMA-fed, 4 = split out, 5 = both
cases = [(256, 128, 1, True, False, 4), (128, 128, 3, False, False, 5), (128, 256, 1, False, True, 3)]
for (Cin, Cout, ks, pre, res, backend) in cases:
    x = rng.normal(size=(B, 64, 64, Cin)).astype(np.float32)
    w = (rng.normal(size=(Cout, ks, ks, Cin)) / np.sqrt(ks * ks * Cin)).astype(np.float32)
    prm = (rng.uniform(0.5, 1.5, Cin).astype(np.float32), rng.normal(size=Cin).astype(np.float32)) if pre else None
    r = rng.normal(size=(B, 64, 64, Cout)).astype(np.float32) if res else None
    for _ in range(2):
        pkpnet.conv2d(ctx, x, w, None, ks, 1, prm, r, not res, backend=backend, tf32_passes=3)
print("wrote", out)
