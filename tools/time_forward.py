"""Quick device-side timing of the network forward for each conv backend (not the bench)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from suo_slam_b200 import _lib, synth  # noqa: E402
from suo_slam_b200.pkpnet import PkpNet  # noqa: E402


def main():
    crops = [int(a) for a in sys.argv[1:]] or [8, 64]
    sd = synth.make_synthetic_state_dict(0, peaky=4.0)
    for L in crops:
        n_img = L // 8
        imgs = torch.rand(n_img, 3, 480, 640, device="cuda")
        boxes = [torch.tensor([[50.0 + 10 * i, 40.0 + 5 * i, 250.0 + 20 * i, 300.0 + 10 * i] for i in range(8)], device="cuda") for _ in range(n_img)]
        for backend, passes in ((2, 3), (1, 3), (1, 1)):
            m = PkpNet(max_crops=L)
            m.return_prob = False
            m.load_state_dict(sd)
            m.cuda()
            ctx = m.context()
            ctx.set_option(_lib.SUO_OPT_CONV_MATH, 1 if backend == 2 else 0)
            ctx.set_option(_lib.SUO_OPT_CONV_BACKEND, min(backend, 1))
            ctx.set_option(_lib.SUO_OPT_TF32_PASSES, passes)
            for _ in range(3):
                m(imgs, boxes)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5 if backend >= 1 else 2
            e0.record()
            for _ in range(n):
                m(imgs, boxes)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            print(f"crops={L} backend={backend} passes={passes}: {ms:.3f} ms/forward  {n_img / ms * 1e3:.1f} frames/s  "
                  f"{31.495 * L / ms:.1f} GFLOP/ms-eq => {31.495e9 * L / (ms * 1e-3) / 1e12:.1f} TFLOP/s", flush=True)
            del m
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
