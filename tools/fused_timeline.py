"""Developer tool: clock64 timeline of CTA 0 of the fused conv2+conv3 kernel on the first fused layer of the network
(r5 at 64x64) with the bench's crop count.  usage: python tools/fused_timeline.py gpurun_out/fused_timeline.csv [crops]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/fused_timeline.csv"
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
os.environ["SUO_FUSED_TIMELINE"] = out
from suo_slam_b200 import _lib, synth  # noqa: E402
from suo_slam_b200.pkpnet import PkpNet  # noqa: E402

m = PkpNet(max_crops=L)
m.load_state_dict(synth.make_synthetic_state_dict(0, peaky=4.0))
m.cuda().eval()
m.context().set_option(_lib.SUO_OPT_USE_GRAPH, 0)
rng = np.random.default_rng(0)
img = torch.from_numpy(rng.random((1, 3, 480, 640), dtype=np.float32)).cuda()
boxes = torch.tensor(np.tile([[50.0, 40.0, 400.0, 380.0]], (L, 1)).astype(np.float32)).cuda()
m(img, [boxes])
torch.cuda.synchronize()
d = np.loadtxt(out, delimiter=",", skiprows=1, comments="#")
print("".join(l for l in open(out) if l.startswith("#")))
print("tiles", len(d))
names = open(out).readline().strip().split(",")
np.set_printoptions(linewidth=250, suppress=True)
print(names)
print(d[:12].astype(np.int64))
per_tile = np.diff(d[:, 1])
print("cycles per tile (mma_start deltas): median", np.median(per_tile))
for a, b in [(1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (8, 9), (10, 11), (12, 13)]:
    print(f"{names[a]} -> {names[b]}: median {np.median(d[2:, b] - d[2:, a]):.0f} cycles")
print("acc2_full -> next tile's acc2_full:", np.median(np.diff(d[:, 8])))
