"""Developer probe: where the 0.2 ms of one drop-in pnp() call go — kernel time of a single-object suo_pnp_batch launch (device pointers, CUDA
events), the host-pointer C call, and the Python wrapper.  usage (GPU box): python tools/pnp_latency_probe.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from suo_slam_b200 import _lib, geometry, runtime, synth  # noqa: E402


def main():
    fr = synth.make_frame(5, n_obj=8)
    o = fr["objs"][0]
    m = o["model_kps_mask"]
    xs = np.ascontiguousarray(o["model_kps"][m], np.float64)
    Kb = o["K_bbox"]
    KinvT = np.linalg.inv(Kb).T
    ys = np.ascontiguousarray(o["uv_meas"][m] @ KinvT[:2, :2] + KinvT[2:3, :2], np.float64)
    ctx = runtime.get_context()
    lib, p = _lib.lib(), _lib.ptr
    n = len(xs)
    print("points:", n)
    # (1) kernel alone: device pointers, CUDA events
    d = lambda a: torch.from_numpy(a).cuda()
    dxs, dys, doff = d(xs), d(ys), d(np.array([0, n], np.int32))
    dT, dst = torch.zeros(16, dtype=torch.float64, device="cuda"), torch.zeros(5, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(10):
        ctx.check(lib.suo_pnp_batch(ctx.handle, p(dxs), p(dys), p(doff), 1, 0.001, 0, None, p(dT), p(dst), 1, s))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        ctx.check(lib.suo_pnp_batch(ctx.handle, p(dxs), p(dys), p(doff), 1, 0.001, 0, None, p(dT), p(dst), 1, s))
    e1.record()
    torch.cuda.synchronize()
    print(f"kernel (device pointers, back to back): {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per launch; stats {dst.cpu().numpy().tolist()}")
    # (2) the host-pointer C call
    off = np.array([0, n], np.int32)
    T, st = np.zeros((1, 4, 4)), np.zeros((1, 5), np.int32)
    for _ in range(10):
        ctx.check(lib.suo_pnp_batch(ctx.handle, p(xs), p(ys), p(off), 1, 0.001, 0, None, p(T), p(st), 0, None))
    t0 = time.perf_counter()
    for _ in range(200):
        ctx.check(lib.suo_pnp_batch(ctx.handle, p(xs), p(ys), p(off), 1, 0.001, 0, None, p(T), p(st), 0, None))
    print(f"host-pointer C call: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
    # (3) the Python wrappers
    t0 = time.perf_counter()
    for _ in range(200):
        geometry.lambdatwist_pnp(xs, ys)
    print(f"lambdatwist.pnp drop-in: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
    t0 = time.perf_counter()
    for _ in range(200):
        geometry.pnp(xs, o["uv_meas"][m].astype(np.float64), Kb)
    print(f"pnp() of lib/object_slam.py:25-41: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")


if __name__ == "__main__":
    main()
