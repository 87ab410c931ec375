"""Developer tool: enable the raw-TMA A path for one conv op at a time and report the logit change vs. all-off."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from suo_slam_b200 import _lib, synth  # noqa: E402
from suo_slam_b200.pkpnet import PkpNet  # noqa: E402

g = np.load("tests/golden/net_small.npz")
sd = synth.make_synthetic_state_dict(seed=0, peaky=4.0)


def run():
    m = PkpNet(input_res=(64, 64), max_crops=8)
    m.load_state_dict(sd)
    m.cuda().eval()
    m.context().set_option(_lib.SUO_OPT_USE_GRAPH, 0)
    out = m(torch.from_numpy(g["img"]).cuda(), [torch.from_numpy(g["boxes"]).cuda()], None)
    torch.cuda.synchronize()
    return out["prob_logits"].cpu().numpy()


os.environ["SUO_RAW_TMA"] = "0"
base = run()
print("raw off: err vs golden", np.abs(base - g["logits"]).max())
os.environ["SUO_RAW_TMA"] = "1"
for op in range(0, 206):
    os.environ["SUO_RAW_ONLY_OP"] = str(op)
    d = np.abs(run() - base).max()
    if d > 1e-3:
        print("op", op, "diff", d, flush=True)
os.environ["SUO_RAW_ONLY_OP"] = "-1"
print("raw all: err vs golden", np.abs(run() - g["logits"]).max())
