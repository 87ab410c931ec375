"""Per-chunk clock64 timeline of CTA 0 of the persistent tcgen05 conv kernel (developer tool)."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.csv"
os.environ["SUO_CONV_TIMELINE"] = out
from suo_slam_b200 import _lib, pkpnet  # noqa: E402

ctx = _lib.Context(0, 1, 64, 41)
rng = np.random.default_rng(0)
for (B, H, W, Cin, Cout, ks) in [(64, 4, 4, 128, 128, 3), (64, 64, 64, 128, 128, 3), (64, 64, 64, 256, 128, 1), (64, 64, 64, 128, 256, 1)]:
    x = rng.normal(size=(B, H, W, Cin)).astype(np.float32)
    w = (rng.normal(size=(Cout, ks, ks, Cin)) / np.sqrt(ks * ks * Cin)).astype(np.float32)
    for backend, passes in ((2, 3), (1, 3)):
        pkpnet.conv2d(ctx, x, w, None, ks, 1, None, None, False, backend=backend, tf32_passes=passes)
        pkpnet.conv2d(ctx, x, w, None, ks, 1, None, None, False, backend=backend, tf32_passes=passes)
print("wrote", out)
