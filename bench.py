"""bench.py — frames/s of the SUO-SLAM per-frame hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                    # this repo's CUDA path, workload c2 (BASELINE configs[1])
    python bench.py --impl reference --steps K --warmup W            # the reference's CPU path (oracle port) on the same workload
    python bench.py --workload ba512|c4|c5|latency ...               # the other BASELINE configs / the drop-in call pattern

Workload c2 = BASELINE.json configs[1]: synthetic YCBV-shape stream, 640x480 frames, 8 object crops per frame, 256x256
crops -> 41-channel 64x64 heat-maps; one *step* is one pass of the whole single-view frame path (crop -> hourglass ->
soft-argmax/cov -> gating -> per-object PnP -> single-view LM-BA) over a batch of `--frames-per-step` independent frames
per GPU.  Weights are the fiducial ("marker") network of suo_slam_b200/synth.py and every frame carries one coloured disc
per model keypoint, so the gate passes real keypoints, every object runs RANSAC + refine and every frame a non-empty
4-round BA — on this arm and on the CPU arm alike (`timed_work` in the JSON line records the counts).
Frames of a single-view stream are independent, so ranks take disjoint frames and exchange the per-crop
result records (poses, flags, keypoints, covariances; 1240 B / crop) with ONE ncclAllGather per step.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same path through the
host-pointer C ABI (pinned host buffers in, host results out, copies inside the timed region; double-buffered with
suo_frames_u8_submit / suo_frames_wait so the copies of batch i+1 overlap batch i).  Timing: CUDA events on the launching
stream, barrier + synchronize on both sides, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, NUM_KP = 480, 640, 41
GFLOP_PER_CROP_256 = 31.495          # 15.7475 GMAC per 256x256 crop (convs only); x4 at 512x512
METRIC = "frames/sec (8 obj-crops/frame, 640x480)"
KP_VAR_THRESH, BBOX_THRESH = 0.2, 0.9      # evaluate.py:58-76 (YCBV)

WORKLOADS = {
    "c2": dict(crops=8, res=256, frames=32, desc="configs[1]: 640x480 frames, 8 crops/frame, 256x256 -> 41x64x64, net+reduce+gating+PnP+single-view BA"),
    "c4": dict(crops=8, res=256, frames=64, desc="configs[3]: fixed 64-frame sequence, 8 crops/frame, frames sharded round-robin over the GPUs, one all-gather of the result records"),
    "c5": dict(crops=16, res=512, frames=8, desc="configs[4]: T-LESS-shape frames, 16 crops/frame, 512x512 -> 41x128x128, half the objects symmetric (device-rendered keypoint priors, second dependent forward)"),
    "latency": dict(crops=8, res=256, frames=1, desc="configs[1] one frame at a time through the drop-in calls: model(img,[bboxes],[priors]) + pnp() per object + optimize() (lib/object_slam.py:1099,1144,443-451)"),
    "ba512": dict(crops=0, res=0, frames=0, desc="configs[2]: batched LM bundle adjustment only, 512 objects x 12 keypoints, 20 LM iterations"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames-per-step", type=int, default=0, help="independent frames per GPU per step (0 = the workload's default)")
    ap.add_argument("--input-sets", type=int, default=3, help="distinct input batches rotated between steps")
    ap.add_argument("--tf32-passes", type=int, default=3, choices=[1, 3])
    ap.add_argument("--conv-math", default="fp16x3", choices=["fp16x3", "tf32"])
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-native-allgather", action="store_true", help="exchange the records with torch.distributed instead of suo_allgather_results")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# inputs
def make_batch(seed0, n_frames, crops=8, res=256, want_f32=False):
    """n_frames synthetic marker frames -> flat per-crop arrays (host)."""
    from suo_slam_b200 import frames, synth
    imgs, imgs_u8, boxes, bi, mk, mm, kb, diam, tgt, uvgt = [], [], [], [], [], [], [], [], [], []
    for f in range(n_frames):
        fr = synth.make_marker_frame(seed0 + f, n_obj=crops, H=H, W=W, res=res, radius=synth.MARKER_RADIUS * res / 256)
        imgs_u8.append(fr["img"])                                                # [H,W,3] u8, what process_view receives
        if want_f32:
            imgs.append(fr["img"].transpose(2, 0, 1).astype(np.float32) / 255.0)     # object_slam.py:1092 (CPU reference arm)
        bb = [o["bbox"] for o in fr["objs"]]
        boxes += bb
        bi += [f] * crops
        mk += [o["model_kps"] for o in fr["objs"]]
        mm += [o["model_kps_mask"] for o in fr["objs"]]
        diam += [o["diameter"] for o in fr["objs"]]
        tgt += [o["T_OtoC"] for o in fr["objs"]]
        uvgt += [o["uv_gt"] for o in fr["objs"]]
        kb.append(frames.k_bbox_for(fr["K"], bb))
    out = dict(images_u8=np.ascontiguousarray(np.stack(imgs_u8)), boxes=np.stack(boxes).astype(np.float32), box_img=np.asarray(bi, np.int32),
               model_kps=np.stack(mk), model_mask=np.stack(mm).astype(np.uint8), K_bbox=np.concatenate(kb),
               diameter=np.asarray(diam, np.float64))
    extra = dict(T_gt=np.stack(tgt), uv_gt=np.stack(uvgt))
    if want_f32:
        extra["images"] = np.ascontiguousarray(np.stack(imgs))
    return out, extra


def frame_config(workload, F, world):
    """The `config` object of the JSON line — identical for the native and the reference arm."""
    wl = WORKLOADS[workload]
    return {"workload": wl["desc"], "frames_per_step": F, "crops_per_frame": wl["crops"], "crop_res": wl["res"],
            "weights": "synthetic fiducial network (reference architecture and key names; 41 trunk channels carry a colour-marker detector "
                       "through the skip connections, all other weights seeded random)",
            "frames": "uniform-noise background + one coloured disc per model keypoint, seeded"}


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def host_cores():
    """Cores this process may actually use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def load_peaks():
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(pk_path)) if os.path.exists(pk_path) else {}


def work_counts(used, T_pnp, bain, diam):
    """What the solver stages of one step actually did (host arrays of the step's outputs)."""
    used, bain = np.asarray(used).astype(bool), np.asarray(bain).astype(bool)
    T = np.asarray(T_pnp).reshape(len(used), 4, 4)
    solved = ~np.all(np.isclose(T, np.eye(4)), axis=(1, 2))
    accepted = solved & (used.sum(1) >= 4) & (T[:, 2, 3] > 0.5 * np.asarray(diam))
    return {"crops": int(len(used)), "gated_kp_per_crop": float(used.sum() / max(len(used), 1)),
            "pnp_objects_run": int((used.sum(1) >= 4).sum()), "pnp_objects_solved": int(solved.sum()),
            "objects_accepted": int(accepted.sum()), "ba_edges": int(used[accepted].sum()), "ba_inlier_edges": int(bain.sum())}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU path (oracle port) on this box's host cores
def cpu_frames_per_s(workload, seconds, steps=None, warmup=1):
    """torch-CPU net + restated Lambda-Twist/Ceres PnP + g2o BA (oracle/), one frame of the workload per step, with every
    usable host core (torch intra-op threads; the solvers are single-threaded like the reference's, object_slam.py:440-442)."""
    import torch
    from oracle import frame_oracle
    from suo_slam_b200 import synth
    wl = WORKLOADS[workload if workload in ("c2", "c4", "c5", "latency") else "c2"]
    sd = synth.make_marker_state_dict(0)
    b, x = make_batch(1000, 1, wl["crops"], wl["res"], want_f32=True)
    cores = host_cores()
    torch.set_num_threads(cores)
    times, work = [], None
    t_end = time.perf_counter() + seconds
    i = 0
    while True:
        t0 = time.perf_counter()
        r = frame_oracle.run_frames(sd, x["images"], b["boxes"], b["box_img"], b["model_kps"], b["model_mask"].astype(bool),
                                    b["K_bbox"], b["diameter"], input_res=(wl["res"], wl["res"]),
                                    kp_var_thresh=KP_VAR_THRESH, bbox_thresh=BBOX_THRESH)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        work = work_counts(r["kp_used"], r["T_pnp"], r["ba_inliers"], b["diameter"])
        i += 1
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.perf_counter() > t_end and len(times) >= 2:
            break
    return 1.0 / float(np.median(times)), len(times), cores, work


def cpu_ba512_objects_per_s(seconds, steps=None, n_obj=64):
    """oracle g2o-LM restatement on a bounded sample of configs[2] (n_obj of the 512 objects per step), single-threaded."""
    from oracle import geom
    from suo_slam_b200 import synth
    pr = synth.make_ba_problem(0, n_obj=n_obj, n_kp=12)
    times = []
    t_end = time.perf_counter() + seconds
    while True:
        t0 = time.perf_counter()
        for o in range(n_obj):
            poses = np.zeros((2, 3, 4)); poses[0] = pr["T_init"][o]; poses[1, :, :3] = np.eye(3)
            geom.ba_optimize(poses, np.array([0, 1], np.uint8), np.zeros(12, np.int32), np.ones(12, np.int32), np.tile(pr["cam_k"], (12, 1)),
                             pr["p_O"][o], pr["uv"][o], pr["info"][o].reshape(12, 4), np.ones(12), [20])
        times.append(time.perf_counter() - t0)
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.perf_counter() > t_end and len(times) >= 2:
            break
    return n_obj / float(np.median(times)), len(times)


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = args.workload
    if wl == "ba512":
        ops, n = cpu_ba512_objects_per_s(0, steps=max(1, args.steps))
        print(json.dumps({
            "impl": "reference", "metric": "objects/sec (LM-BA, 512 objects x 12 keypoints, 20 iterations)", "value": ops, "unit": "objects/s",
            "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": 64e3 / ops, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": ba512_config(),
            "cpu_baseline": {"value": ops, "unit": "objects/s", "cores": 1, "kind": "port",
                             "sample": f"{n} steps of 64 of the 512 objects each through oracle/ (g2o LM restatement, single-threaded like the reference)"},
            "e2e": {"value": ops, "unit": "objects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    if wl == "c5":
        fps, n, cores = cpu_c5_frames_per_s(1e9, n_views=max(2, min(args.steps + 1, 6)))
        print(json.dumps({
            "impl": "reference", "metric": "frames/sec (16 obj-crops/frame, 640x480, SLAM mode)", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": n - 1,
            "warmup": 1, "ms_per_step": 1e3 / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 net / f64 solvers", "data": "synthetic",
            "config": dict(frame_config("c5", 1, args.gpus), views_per_sequence=WORKLOADS["c5"]["frames"], symmetric_objects=8, thresholds="T-LESS (evaluate.py:68-76)"),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": f"{n} views of one sequence through oracle/slam_frame_oracle.py"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    F = args.frames_per_step or WORKLOADS[wl]["frames"]
    if wl == "c4":
        F = WORKLOADS[wl]["frames"] // max(1, args.gpus)
    fps, n, cores, work = cpu_frames_per_s(wl, 0, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": 1e3 / fps, "higher_is_better": True,
        "scaling": "strong" if wl == "c4" else "weak", "vs_baseline": None, "dtype": "f32 net / f64 solvers", "data": "synthetic",
        "config": frame_config(wl, F, args.gpus),
        "timed_work": work,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{n} steps of ONE frame ({WORKLOADS[wl]['crops']} crops) of the workload each, oracle/ "
                                   "(the reference's g2o / Ceres / lambdatwist do not build here); torch intra-op threads = cores"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------------------
def conv_classes(dump_path, L, pair=True):
    """Per-op table of suo_profile_network -> kernel classes (which conv_tc_persistent_kernel instance runs the op).
    bytes = ALGORITHMIC bytes: every input, skip and output tensor of the layer once, 4 B per element, plus weights."""
    import csv
    names = {  # (mode, pre, res) -> (ncu class of the 128-wide instance, description, bound)
        (1, 0, 0): ("conv_tc_persistent_kernel<128, 1, 0, 1, 1, 1>", "3x3 convs (TMA-fed A, tcgen05 fp16x3)", "tensor"),
        (0, 0, 1): ("conv_tc_persistent_kernel<128, 0, 0, 1, 1, 2>", "1x1 convs + skip add (TMA-fed A, skip prefetched by TMA)", "hbm"),
        (0, 1, 0): ("conv_tc_persistent_kernel<128, 0, 1, 1, 0, 3>", "1x1 convs with BN+ReLU prologue (raw FP32 by TMA)", "hbm"),
        (0, 0, 0): ("conv_tc_persistent_kernel<128, 0, 0, 1, 0, 3>", "plain 1x1 convs", "hbm"),
        (2, 0, 0): ("conv_tc_persistent_kernel<64, 0, 0, 1, 0, 3>" if os.environ.get("SUO_STEM_TMA", "1") != "0" else "conv_tc_persistent_kernel<64, 2, 0, 1, 0, 1>",
                    "7x7/2 stem (TMA-fed from the zero-bordered input copy)" if os.environ.get("SUO_STEM_TMA", "1") != "0" else "7x7/2 stem", "hbm"),
    }
    out = {}
    for r in csv.DictReader(open(dump_path)):
        if int(r["type"]) != 0:
            continue
        key = (int(r["mode"]), int(r["pre"]), int(r["res"]))
        ncu, desc, bound = names.get(key, ("conv_tc_persistent_kernel", "other convs", "hbm"))
        side, cin, cout, K = int(r["side_out"]), int(r["Cin"]), int(r["Cout"]), int(r["K"])
        if key == (1, 0, 0) and cout == 128 and pair and side in (16, 32, 64) and os.environ.get("SUO_HALO", "1") != "0":
            ncu, desc = "conv3x3_halo_kernel", "3x3 convs (A-halo CTA pair: activations once per column shift, tcgen05.mma.cta_group::2, fp16x3)"
        elif key == (1, 0, 0) and cout == 128 and pair:      # the bottleneck's conv2 runs as a CTA pair (csrc/conv_pair.cu)
            ncu, desc = "conv3x3_pair_kernel", "3x3 convs (CTA pair: tcgen05.mma.cta_group::2 of M = 256, TMA-fed, fp16x3)"
        px = L * side * side
        px_in = px * 4 if key[0] == 2 else px
        b = 4.0 * (px_in * cin + px * cout * (2 if key[2] else 1)) + 4.0 * K * cout
        c = out.setdefault(desc, dict(name=f"{ncu}: {desc}", ncu_class=ncu, bound=bound, n=0, ms=0.0, gflop=0.0, bytes=0.0))
        c["n"] += 1; c["ms"] += float(r["ms"]); c["gflop"] += float(r["gflop"]); c["bytes"] += b
    try:
        os.remove(dump_path)
    except OSError:
        pass
    return out


def dtype_name(args):
    if args.conv_math == "fp16x3":
        return "fp16x3 (each FP32 operand split into two FP16 numbers, 22-bit mantissa, FP32 accumulate: fp32-equivalent) convs + f32 reductions + f64 solvers"
    if args.tf32_passes == 3:
        return "tf32x3 (3xTF32 split, fp32-equivalent) convs + f32 reductions + f64 solvers"
    return "tf32 convs + f32 reductions + f64 solvers"


def dist_env():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def pose_parity(ctx, b, x, outs, n_frames, crops, res, dev, sp):
    """BASELINE.json's "pose L2 err" on the keypoints the timed path itself produced: the first n_frames frames of the batch that was just
    run.  (1) the network's uv / cov / kp_mask of those crops go through the CPU oracle's gating -> PnP -> single-view BA and must give the
    poses the timed path returned (bar: 1e-4 relative); (2) the poses against the frame's ground truth."""
    import torch
    from oracle import frame_oracle
    from suo_slam_b200 import _lib
    lib, p = _lib.lib(), _lib.ptr
    F = b["images_u8"].shape[0]
    L, n = F * crops, n_frames * crops
    f32 = dict(dtype=torch.float32, device=dev)
    img = torch.from_numpy(np.ascontiguousarray(b["images_u8"].transpose(0, 3, 1, 2).astype(np.float32) / 255.0)).to(dev)     # object_slam.py:1092 (host division, as there)
    boxes, bi = torch.from_numpy(b["boxes"]).to(dev), torch.from_numpy(b["box_img"]).to(dev)
    uv, cov, km = torch.empty((L, NUM_KP, 2), **f32), torch.empty((L, NUM_KP, 4), **f32), torch.empty((L, NUM_KP), **f32)
    ctx.check(lib.suo_forward(ctx.handle, p(img.contiguous()), F, H, W, p(boxes), p(bi), L, None, p(uv), p(cov), None, None, None, p(km), None, 1, sp))
    torch.cuda.synchronize(dev)
    uv, cov, km = uv.cpu().numpy()[:n], cov.cpu().numpy()[:n].reshape(n, NUM_KP, 2, 2), km.cpu().numpy()[:n]
    got = {k: v.cpu().numpy()[:n] for k, v in outs.items()}
    same_kp = bool(np.array_equal(uv, got["uv"]) and np.array_equal(cov.reshape(n, NUM_KP, 4), got["cov"]))
    ref = frame_oracle.solve_from_keypoints(uv, cov, km, b["model_kps"][:n], b["model_mask"][:n], b["K_bbox"][:n], b["diameter"][:n], b["box_img"][:n],
                                            kp_var_thresh=KP_VAR_THRESH, bbox_thresh=BBOX_THRESH, seed=0)
    acc = np.nonzero(ref["accepted"])[0]
    rel = lambda a_, b_: float(np.linalg.norm(a_ - b_) / max(np.linalg.norm(b_), 1e-300))
    T_pnp, T_ba = got["T_pnp"].reshape(n, 4, 4), got["T_ba"].reshape(n, 3, 4)
    d_pnp = [rel(T_pnp[c][:3], ref["T_pnp"][c][:3]) for c in acc]
    d_ba = [rel(T_ba[c], ref["T_ba"][c]) for c in acc]
    t_gt = [rel(T_ba[c][:, 3], x["T_gt"][c][:3, 3]) for c in acc]
    uv_err = np.abs(uv - x["uv_gt"][:n])[got["used"].astype(bool)]
    return {"objects": int(n), "accepted_by_oracle": int(len(acc)), "keypoints_identical_to_timed_path": same_kp,
            "same_gating": bool(np.array_equal(got["used"].astype(bool), ref["kp_used"])),
            "same_ba_inliers": bool(np.array_equal(got["bain"].astype(bool), ref["ba_inliers"])),
            "pnp_rel_l2_vs_oracle_max": max(d_pnp) if d_pnp else None, "ba_rel_l2_vs_oracle_max": max(d_ba) if d_ba else None,
            "translation_rel_l2_vs_ground_truth_median": float(np.median(t_gt)) if t_gt else None,
            "gated_keypoint_abs_err_ndc_median": float(np.median(uv_err)) if uv_err.size else None,
            "inputs": "the timed path's own network keypoints (marker frames); north_star bar: 1e-4 relative vs the CPU solve on identical inputs"}


def run_native_frames(args):
    import ctypes as C
    import tempfile
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from suo_slam_b200 import _lib, dist as sdist, synth
    from suo_slam_b200.pkpnet import PkpNet

    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = WORKLOADS[args.workload]
    crops, res = wl["crops"], wl["res"]
    if args.workload == "c4":      # strong scaling: the fixed 64-frame sequence is split over the ranks
        F = wl["frames"] // world
    else:
        F = args.frames_per_step or wl["frames"]
    L = F * crops
    dev = torch.device("cuda", local)

    model = PkpNet(input_res=(res, res), max_crops=L)
    model.load_state_dict(synth.make_marker_state_dict(0))
    model.cuda(local)
    ctx = model.context()
    ctx.set_option(_lib.SUO_OPT_TF32_PASSES, args.tf32_passes)
    ctx.set_option(_lib.SUO_OPT_CONV_MATH, 1 if args.conv_math == "fp16x3" else 0)
    lib, hdl, p = _lib.lib(), ctx.handle, _lib.ptr
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    # distinct input sets per rank (disjoint frames of the stream), rotated between steps
    n_sets = max(2, args.input_sets)
    if args.workload == "c4":      # frame f of the sequence -> rank f mod world; every step is the same fixed sequence
        made = [make_batch(50_000 + world * f + rank, 1, crops, res) for f in range(F)]
        one = {k: np.concatenate([m[0][k] for m in made]) for k in made[0][0]}
        one["box_img"] = np.repeat(np.arange(F, dtype=np.int32), crops)
        x0 = {k: np.concatenate([m[1][k] for m in made]) for k in made[0][1]}
        sets_h, sets_x = [one] * n_sets, [x0] * n_sets
    else:
        made = [make_batch(10_000 * rank + 100 * s, F, crops, res) for s in range(n_sets)]
        sets_h, sets_x = [m[0] for m in made], [m[1] for m in made]
    pin = lambda a: torch.from_numpy(a).pin_memory()
    sets_pin = [{k: pin(v) for k, v in b.items()} for b in sets_h]
    sets_dev = [{k: v.to(dev) for k, v in b.items()} for b in sets_pin]
    f64 = dict(dtype=torch.float64, device=dev)
    mk_outs = lambda **kw: dict(T_pnp=torch.zeros((L, 16), dtype=torch.float64, **kw), T_ba=torch.zeros((L, 12), dtype=torch.float64, **kw),
                                used=torch.zeros((L, NUM_KP), dtype=torch.uint8, **kw), bain=torch.zeros((L, NUM_KP), dtype=torch.uint8, **kw),
                                uv=torch.zeros((L, NUM_KP, 2), **kw), cov=torch.zeros((L, NUM_KP, 4), **kw))
    outs_dev = mk_outs(device=dev)
    outs_host = [{k: v.pin_memory() for k, v in mk_outs().items()} for _ in range(2)]
    xch = [sdist.RecordExchange(ctx, L, world, local, native=not args.no_native_allgather and s == 0) for s in range(2)]
    if world > 1 and xch[0].comm is not None:
        xch[1].comm, xch[1].how = xch[0].comm, xch[0].how
    gathered_host = [torch.zeros((world * L, xch[0].rb), dtype=torch.uint8).pin_memory() for _ in range(2)]
    ev = [torch.cuda.Event() for _ in range(2)]
    id_base = rank * L

    def frame_args(b):
        return (p(b["images_u8"]), F, H, W, p(b["boxes"]), p(b["box_img"]), L)

    def solver_args(b):
        return (p(b["model_kps"]), p(b["model_mask"]), p(b["K_bbox"]), p(b["diameter"]), KP_VAR_THRESH, BBOX_THRESH, 0, 1)

    def step_device(i):
        b, o = sets_dev[i % n_sets], outs_dev
        # the camera's u8 frames go in as they are (lib/object_slam.py:327-328); float32(img)/255 (:1092) happens per tap on the GPU
        ctx.check(lib.suo_frames_u8(hdl, *frame_args(b), None, *solver_args(b), p(o["T_pnp"]), p(o["T_ba"]), p(o["used"]), p(o["bain"]),
                                    p(o["uv"]), p(o["cov"]), 1, sp))
        # the single exchange of the step: fixed-size result records of every rank's crops (packed on the device)
        xch[0].run(id_base, o["T_pnp"], o["T_ba"], o["used"], o["bain"], o["uv"], o["cov"], sp)

    def submit(i):
        b, sl = sets_pin[i % n_sets], i % 2
        o = outs_host[sl]
        ctx.check(lib.suo_frames_u8_submit(hdl, sl, *frame_args(b), *solver_args(b), p(o["T_pnp"]), p(o["T_ba"]), p(o["used"]), p(o["bain"]),
                                           p(o["uv"]), p(o["cov"]), p(xch[sl].rec) if world > 1 else None, id_base, sp))
        if world > 1:
            xch[sl].gather(sp)
            gathered_host[sl].copy_(xch[sl].out, non_blocking=True)
            ev[sl].record(stream)

    def wait(i):
        ctx.check(lib.suo_frames_wait(hdl, i % 2))
        if world > 1:
            ev[i % 2].synchronize()

    def run_e2e(n):
        submit(0)
        for i in range(1, n):
            submit(i)
            wait(i - 1)
        wait(n - 1)

    def timed(fn_steps, steps, warm):
        fn_steps(warm)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.kernel_launches()
        e0.record(stream)
        fn_steps(steps)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ctx.kernel_launches() - l0

    def dev_steps(n):
        for i in range(n):
            step_device(i)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    W_ = max(3, args.warmup)
    ms_dev, launches = timed(dev_steps, args.steps, W_)
    ms_e2e, _ = timed(run_e2e, args.steps, 2)
    ms_dev2, _ = timed(dev_steps, max(3, args.steps // 2), 1)      # the device-resident loop again, AFTER the e2e loop (clock / power drift between the two legs)
    clocks = sampler.stop() if rank == 0 else None
    ctx.check(lib.suo_check_range(hdl))       # fp16x3: no activation left the FP16 range during the run

    # what the solver stages did in the timed region: the outputs of one step of the device loop (set 0), read back
    step_device(0)
    torch.cuda.synchronize(dev)
    work = work_counts(outs_dev["used"].cpu().numpy(), outs_dev["T_pnp"].cpu().numpy(), outs_dev["bain"].cpu().numpy(), sets_h[0]["diameter"])
    tw = torch.tensor([work["crops"], work["gated_kp_per_crop"] * work["crops"], work["pnp_objects_run"], work["pnp_objects_solved"],
                       work["objects_accepted"], work["ba_edges"], work["ba_inlier_edges"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw)
        rec_all = sdist.unpack_records(xch[0].out)
        assert len(rec_all) == world * L and np.array_equal(rec_all["crop_id"], np.arange(world * L)), "all-gather lost records"
    tw = tw.cpu().numpy()
    work = {"per_step_all_gpus": True, "crops": int(tw[0]), "gated_kp_per_crop": float(tw[1] / tw[0]), "pnp_objects_run": int(tw[2]),
            "pnp_objects_solved": int(tw[3]), "objects_accepted": int(tw[4]), "ba_edges": int(tw[5]), "ba_inlier_edges": int(tw[6])}
    if work["gated_kp_per_crop"] < 4 or work["objects_accepted"] < work["crops"] // 2 or work["ba_edges"] == 0:
        raise SystemExit(f"bench: the timed region did no real solver work: {work}")

    # roofline: per-op CUDA events over the network program (eager launches on `stream`, after the timed region)
    conv_ms, other_ms = C.c_float(0), C.c_float(0)
    dump_path = os.path.join(tempfile.gettempdir(), f"suo_per_op_{os.getpid()}.csv")
    os.environ["SUO_PROFILE_DUMP"] = dump_path
    ctx.check(lib.suo_profile_network(hdl, L, 0, 2, C.byref(conv_ms), C.byref(other_ms), sp))
    torch.cuda.synchronize(dev)
    os.environ.pop("SUO_PROFILE_DUMP", None)

    # stand-alone stage timings (CUDA events on `stream`): heat-map reduction on buffers larger than L2, solver stages on the step's keypoints
    def ev_ms(fn, reps):
        fn(); torch.cuda.synchronize(dev)
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b_.record(stream)
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b_) / reps
    HM = res // 4
    lg = [torch.randn((L, NUM_KP, HM, HM), device=dev) for _ in range(2)]
    tick = [0]

    def red():
        tick[0] ^= 1
        ctx.check(lib.suo_heatmap_reduce(hdl, p(lg[tick[0]]), L, NUM_KP, HM, HM, None, None, p(outs_dev["uv"]), p(outs_dev["cov"]), None, None, None, None, 1, sp))
    red_ms = ev_ms(red, 10)
    b0 = sets_dev[0]
    km = torch.full((L, NUM_KP), 0.9, device=dev)

    def solve(run_ba):
        ctx.check(lib.suo_solve_keypoints(hdl, p(outs_dev["uv"]), p(outs_dev["cov"]), p(km), p(b0["box_img"]), F, L, p(b0["model_kps"]), p(b0["model_mask"]),
                                          p(b0["K_bbox"]), p(b0["diameter"]), KP_VAR_THRESH, BBOX_THRESH, 0, run_ba, p(outs_dev["T_pnp"]), p(outs_dev["T_ba"]),
                                          p(outs_dev["used"]), p(outs_dev["bain"]), 1, sp))
    step_device(0)
    pnp_ms = ev_ms(lambda: solve(0), 5)
    step_device(0)
    all_ms = ev_ms(lambda: solve(1), 5)
    del lg

    if rank == 0:
        frames_total = F * world * args.steps
        fps = frames_total / (ms_dev * 1e-3)
        fps_e2e = frames_total / (ms_e2e * 1e-3)
        peaks = load_peaks()
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_bw = float(peaks.get("hbm_gbs", 6500.0))
        gflop_crop = GFLOP_PER_CROP_256 * (res / 256) ** 2
        # DRAM traffic of the conv kernels of one step: from the committed ncu launch list of this same command; only valid for the
        # crop count it was captured at
        traffic, tr = None, {}
        for name in ("r2_conv_traffic.json", "r1_conv_traffic.json"):
            tr_path = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tr_path):
                tr = json.load(open(tr_path))
                if int(tr.get("crops_per_step", -1)) == L and int(tr.get("crop_res", 256)) == res:
                    traffic = float(tr["conv_dram_bytes_per_step"])
                    break
                tr = {}
        achieved_all = gflop_crop * L / max(conv_ms.value, 1e-6)          # GFLOP / ms == TFLOP/s
        if os.environ.get("SUO_BENCH_PER_OP"):            # keep the per-op table (profiles/)
            import shutil
            shutil.copy(dump_path, os.environ["SUO_BENCH_PER_OP"])
        classes = conv_classes(dump_path, L, pair=os.environ.get("SUO_PAIR", "1") != "0")
        dom = max(classes.values(), key=lambda c: c["ms"])                     # the kernel with the largest share of the step
        hbm_cls = max((c for c in classes.values() if c["bound"] == "hbm"), key=lambda c: c["ms"])
        ten_cls = max((c for c in classes.values() if c["bound"] == "tensor"), key=lambda c: c["ms"])
        tr_cls = tr.get("classes", {}) if traffic is not None else {}
        step_ms = ms_dev / args.steps

        def cls_traffic(c):     # per-launch DRAM bytes of that kernel from the committed ncu launch list
            t = tr_cls.get(c["ncu_class"])
            return None if not t else t["dram_bytes"] / t["launches"]

        def roof(c):
            """Roofline of one kernel class against the resource that bounds it.  achieved = algorithmic work of these launches / their summed
            duration (CUDA events per launch); traffic / algorithmic_bytes are per launch (average)."""
            base = {"bound": c["bound"], "kernel": c["name"], "launches_per_step": c["n"], "share_of_step": c["ms"] / step_ms,
                    "traffic": cls_traffic(c), "algorithmic_bytes": c["bytes"] / c["n"]}
            if c["bound"] == "tensor":
                tf = c["gflop"] / c["ms"]
                base.update({"achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "mma_achieved": 3 * tf, "mma_frac": 3 * tf / peak_tf,
                             "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400 (of fallback)",
                             "note": "algorithmic FLOPs = 2*M*N*K per conv; the reference computes in FP32 and tcgen05 has no FP32 MMA, so each product is "
                                     "3 f16-kind MMAs (fp16x3 split): mma_achieved = 3 x achieved is what the tensor pipe executes, the ceiling of frac is 1/3"})
            else:
                gb = c["bytes"] / c["ms"] * 1e-6
                base.update({"achieved": gb, "peak": peak_bw, "unit": "GB/s", "frac": gb / peak_bw,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6500 (of fallback)"})
            return base

        h2d = sum(v.numel() * v.element_size() for v in sets_pin[0].values())
        d2h = sum(v.numel() * v.element_size() for v in outs_host[0].values()) + (gathered_host[0].numel() if world > 1 else 0)
        red_bytes = L * NUM_KP * HM * HM * 4.0
        out = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": W_, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong" if args.workload == "c4" else "weak", "vs_baseline": None,
            "dtype": dtype_name(args),
            "data": "synthetic (fiducial-network weights, marker frames; see config)",
            "config": frame_config(args.workload, F, world),
            "engine": {"crops_per_step_per_gpu": L, "parallelism": f"frames sharded over {world} GPU(s), 1 all-gather of {xch[0].rb}-byte result records per crop per step",
                       "allgather": xch[0].how,
                       "l2": f"{n_sets} input sets rotated; per-step activation working set ({L} crops) >> 126 MB L2" if args.workload != "c4" else
                             "fixed sequence (the same frames every step); per-step activation working set >> 126 MB L2",
                       "conv_backend": "tcgen05", "conv_math": args.conv_math, "tf32_passes": args.tf32_passes,
                       "activation_bytes": dict(zip(("allocated", "one_allocation_per_tensor"), ctx.activation_bytes()))},
            "timed_work": work,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps, "how": "suo_frames_u8_submit / suo_frames_wait, two slots: the copies of batch i+1 overlap the kernels of batch i",
                    "device_resident_ms_per_step_measured_after_this_leg": ms_dev2 / max(3, args.steps // 2)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof(dom),                    # the kernel class with the largest share of the step, against ITS bound
            "roofline_tensor": roof(ten_cls),         # largest tensor-bound class (the 3x3 convs)
            "roofline_hbm": roof(hbm_cls),            # largest HBM-bound class (the 1x1 convs that add the skip tensor)
            "roofline_reduce": {"bound": "hbm", "kernel": "heatmap_reduce_reg_kernel (softmax + soft-argmax + 2x2 covariance + hard argmax, one read of the map)",
                                "achieved": red_bytes / red_ms * 1e-6, "peak": peak_bw, "unit": "GB/s", "frac": red_bytes / red_ms * 1e-6 / peak_bw,
                                "algorithmic_bytes": red_bytes, "ms": red_ms, "timed": "stand-alone, 2 logit buffers alternated (each > L2), CUDA events"},
            "solver_stage": {"gate_pnp_ms_per_step": pnp_ms, "gate_pnp_ba_ms_per_step": all_ms, "objects_per_step": L,
                             "objects_per_s": L / (all_ms * 1e-3), "share_of_step": all_ms / step_ms,
                             "bound": "latency / FP64 issue (SURVEY.md §8d): no bandwidth claim; see --workload ba512 for the BA kernel alone"},
            "conv_engine": {"achieved_all_convs": achieved_all, "unit": "TFLOP/s", "frac": achieved_all / peak_tf, "conv_ms_per_step": conv_ms.value,
                            "other_net_ms_per_step": other_ms.value, "dram_bytes_per_step": traffic,
                            "classes": {k: {"n": c["n"], "ms": round(c["ms"], 4), "TFLOP/s": round(c["gflop"] / c["ms"], 1),
                                            "GB/s": round(c["bytes"] / c["ms"] * 1e-6, 1), "bound": c["bound"]} for k, c in classes.items()}},
        }
        if not args.no_cpu_baseline and world == 1:
            cfps, n, cores, cwork = cpu_frames_per_s(args.workload, args.cpu_baseline_seconds)
            out["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port", "timed_work_per_frame": cwork,
                                   "sample": f"{n} timed frames ({crops} crops each) of the same workload through oracle/ on the host cores (torch threads = cores)"}
            try:
                step_device(0)
                torch.cuda.synchronize(dev)
                out["pose_err"] = pose_parity(ctx, sets_h[0], sets_x[0], outs_dev, min(4, F), crops, res, dev, sp)
            except Exception as e:      # the checker must never take the bench line down
                out["pose_err"] = {"error": repr(e)}
        else:
            out["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": host_cores(), "kind": "port", "sample": "skipped (N>1 or --no-cpu-baseline)"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
def ba512_config():
    return {"workload": WORKLOADS["ba512"]["desc"], "objects_per_step": 512, "keypoints": 12, "its": [20],
            "problem": "camera fixed at identity, object poses perturbed by exp(N(0,(5deg,5deg,5deg,10,10,20 mm))), 1 px noise, cam_k = [320,320,320,240] "
                       "(thirdparty/g2opy/python/examples/object_slam_demo.py:54-150); one LM problem (one lambda) per object"}


def run_native_ba512(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from suo_slam_b200 import _lib, synth
    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = _lib.Context(device=local, max_crops=1, crop_res=256, num_kp=NUM_KP)
    ctx.set_option(_lib.SUO_OPT_BA_BLOCK_DIAGONAL, 2)      # 512 single-object problems: one warp each, no host look at the graph structure
    lib, hdl, p = _lib.lib(), ctx.handle, _lib.ptr
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    n_obj, n_kp = 512, 12
    pr = synth.make_ba_problem(rank, n_obj=n_obj, n_kp=n_kp)       # objects mod N: every rank its own 512 (weak scaling)
    n_e, n_v = n_obj * n_kp, 2 * n_obj
    poses0 = np.zeros((n_v, 3, 4)); poses0[0::2] = pr["T_init"]; poses0[1::2, :, :3] = np.eye(3)
    fixed = np.zeros(n_v, np.uint8); fixed[1::2] = 1
    host = dict(pv=np.arange(0, n_v + 1, 2, dtype=np.int32), pe=np.arange(0, n_e + 1, n_kp, dtype=np.int32), poses0=poses0.reshape(n_v, 12), fixed=fixed,
                eo=np.repeat(np.arange(0, n_v, 2, dtype=np.int32), n_kp), ec=np.repeat(np.arange(1, n_v, 2, dtype=np.int32), n_kp),
                ck=np.tile(pr["cam_k"], (n_e, 1)), pp=pr["p_O"].reshape(n_e, 3), uv=pr["uv"].reshape(n_e, 2), info=pr["info"].reshape(n_e, 4),
                its=np.array([20], np.int32))
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in host.items()}
    d = {k: v.to(dev) for k, v in pinned.items()}
    poses = torch.zeros((n_v, 12), dtype=torch.float64, device=dev)
    inl = torch.ones(n_e, dtype=torch.uint8, device=dev)
    stats = torch.zeros((n_obj, 3), dtype=torch.int32, device=dev)
    h_poses, h_inl, h_stats = torch.zeros((n_v, 12), dtype=torch.float64).pin_memory(), torch.ones(n_e, dtype=torch.uint8).pin_memory(), torch.zeros((n_obj, 3), dtype=torch.int32).pin_memory()

    def call(a, po, il, st, on_dev):
        ctx.check(lib.suo_ba_batch(hdl, n_obj, p(a["pv"]), p(a["pe"]), p(po), p(a["fixed"]), n_v, p(a["eo"]), p(a["ec"]), p(a["ck"]), p(a["pp"]), p(a["uv"]),
                                   p(a["info"]), p(il), n_e, p(a["its"]), 1, float(np.sqrt(5.991)), 5.991, 0, p(st), on_dev, sp))

    def dev_steps(n):
        for _ in range(n):
            poses.copy_(d["poses0"]); inl.fill_(1)
            call(d, poses, inl, stats, 1)

    def e2e_steps(n):
        for _ in range(n):
            h_poses.copy_(pinned["poses0"]); h_inl.fill_(1)
            call(pinned, h_poses, h_inl, h_stats, 0)

    def timed(fn, steps, warm):
        fn(warm)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.kernel_launches()
        e0.record(stream); fn(steps); e1.record(stream)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ctx.kernel_launches() - l0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    W_ = max(3, args.warmup)
    ms_dev, launches = timed(dev_steps, args.steps, W_)
    ms_e2e, _ = timed(e2e_steps, args.steps, 2)
    clocks = sampler.stop() if rank == 0 else None
    st = stats.cpu().numpy()
    if rank == 0:
        trials, outer = int(st[:, 2].sum()), int(st[:, 1].sum())
        step_ms = ms_dev / args.steps
        # per LM trial and edge ~ 60 (residual) flop; per outer iteration and edge ~ 150 (Jacobian) + 170 (J^T W J); 120 per 6x6 solve
        flop = n_kp * (trials * 60.0 + outer * 320.0) + trials * 120.0
        T = poses.cpu().numpy().reshape(n_v, 3, 4)[0::2]
        terr = np.linalg.norm(T[:, :, 3] - pr["T_gt"][:, :, 3], axis=1) / np.linalg.norm(pr["T_gt"][:, :, 3], axis=1)
        out = {"metric": "objects/sec (LM-BA, 512 objects x 12 keypoints, 20 iterations)", "value": n_obj * world * args.steps / (ms_dev * 1e-3), "unit": "objects/s",
               "n_gpus": world, "steps": args.steps, "warmup": W_, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": ba512_config(),
               "timed_work": {"objects": n_obj, "edges": n_e, "lm_outer_iterations": outer, "lm_trials": trials, "lm_trials_per_s": trials / (step_ms * 1e-3),
                              "translation_rel_err_vs_ground_truth_median": float(np.median(terr))},
               "e2e": {"value": n_obj * world * args.steps / (ms_e2e * 1e-3), "unit": "objects/s",
                       "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for k, v in pinned.items()) + h_inl.numel()),
                       "d2h_bytes_per_step": int(h_poses.numel() * 8 + h_inl.numel() + h_stats.numel() * 4), "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": int(launches), "clocks": clocks,
               "roofline": {"bound": "latency / FP64 issue", "kernel": "ba_warp_kernel (one warp per object graph, FP64, LM state in registers, shuffle reductions, no barrier)",
                            "achieved": flop / (step_ms * 1e-3) * 1e-12, "peak": 40.0, "unit": "TFLOP/s (FP64)", "frac": flop / (step_ms * 1e-3) * 1e-12 / 40.0,
                            "peak_source": "nominal B200 FP64 (SURVEY.md §8d); the kernel is bound by the serial LM dependency chain of 512 tiny problems, not by FP64 issue",
                            "algorithmic_flop_per_step": flop, "traffic": None}}
        if not args.no_cpu_baseline and world == 1:
            ops, n = cpu_ba512_objects_per_s(args.cpu_baseline_seconds)
            out["cpu_baseline"] = {"value": ops, "unit": "objects/s", "cores": 1, "kind": "port",
                                   "sample": f"{n} passes over 64 of the 512 objects through oracle/ (g2o LM restatement, single-threaded like the reference, object_slam.py:440-442)"}
        else:
            out["cpu_baseline"] = {"value": None, "unit": "objects/s", "cores": 1, "kind": "port", "sample": "skipped"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
def run_native_latency(args):
    """One frame at a time through the drop-in surface exactly as ObjectSLAM calls it: model(img_th, [bboxes_th], [priors_th]) on CUDA tensors
    (lib/object_slam.py:1092-1099), .cpu().numpy() of uv / kp_mask / cov (:1100-1109), gating (:1110-1115), pnp() per object (:1144) and the
    single-view optimize() (:443-451) as one packed suo_ba_batch call.  Reports the median latency of a frame."""
    import torch
    import __graft_entry__ as ge
    ge.build()
    from suo_slam_b200 import ba, geometry, synth
    from suo_slam_b200.pkpnet import PkpNet
    world, rank, local = dist_env()
    if rank != 0:
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    crops, res = 8, 256
    model = PkpNet(input_res=(res, res), max_crops=32)
    model.load_state_dict(synth.make_marker_state_dict(0))
    model.cuda(local).eval()
    model.return_prob = False
    n_frames = max(8, args.steps)
    frames_ = [make_batch(20_000 + f, 1, crops, res) for f in range(n_frames)]

    def one(b):
        t0 = time.perf_counter()
        img = b["images_u8"][0].astype(np.float32) / 255.0                                     # object_slam.py:1092
        img_th = torch.from_numpy(img.transpose(2, 0, 1)[None]).to(dev)                       # :1093-1096
        pred = model(img_th, [torch.from_numpy(b["boxes"]).to(dev)], None)                     # :1099
        uv, km, cov = pred["uv"].cpu().numpy(), pred["kp_mask"].cpu().numpy(), pred["cov"].cpu().numpy()   # :1100-1109
        t1 = time.perf_counter()
        mask = (km > 0.3) & b["model_mask"].astype(bool)
        mask &= (uv.min(-1) > -BBOX_THRESH) & (uv.max(-1) < BBOX_THRESH)
        mask &= np.all(np.sqrt(cov[..., [0, 1], [0, 1]]) < 2 * KP_VAR_THRESH, axis=-1)        # :1110-1115
        poses, keep = [], []
        for c in range(crops):                                                                  # :1123-1165
            m = mask[c]
            r = geometry.pnp(b["model_kps"][c][m], uv[c][m].astype(np.float64), b["K_bbox"][c]) if m.sum() >= 4 else None
            if r is not None and r[0][2, 3] > 0.5 * b["diameter"][c]:
                poses.append(r[0]); keep.append(c)
        t2 = time.perf_counter()
        n_edges = 0
        if keep:                                                                                # optimize(), :703-930, single-view
            g = ba.single_view_graph(poses, [b["model_kps"][c][mask[c]] for c in keep], [uv[c][mask[c]].astype(np.float64) for c in keep],
                                     [cov[c][mask[c]].astype(np.float64) for c in keep], [b["K_bbox"][c] for c in keep])
            n_edges = len(g["e_obj"])
            ba.ba_batch([0, len(keep) + 1], [0, n_edges], g["poses"], g["fixed"], g["e_obj"], g["e_cam"], g["cam_k"], g["p"], g["uv"], g["info"],
                        np.ones(n_edges), [10, 10, 10, 10], ctx=model.context())
        t3 = time.perf_counter()
        return (t1 - t0, t2 - t1, t3 - t2, t3 - t0, int(mask.sum()), len(keep), n_edges)

    sampler = ClockSampler(local)
    sampler.start()
    for f in range(max(3, args.warmup)):
        one(frames_[f % n_frames][0])
    rows = np.array([one(frames_[f % n_frames][0]) for f in range(max(args.steps, 20))])
    clocks = sampler.stop()
    med = np.median(rows[:, :4], axis=0) * 1e3
    out = {"metric": METRIC, "value": 1e3 / med[3], "unit": "frames/s", "n_gpus": 1, "steps": len(rows), "warmup": max(3, args.warmup), "ms_per_step": float(med[3]),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_name(args), "data": "synthetic (fiducial-network weights, marker frames)",
           "config": frame_config("latency", 1, 1),
           "latency_ms": {"model_call_incl_h2d_d2h": float(med[0]), "gating_plus_pnp_per_object_calls": float(med[1]), "optimize_single_view": float(med[2]),
                          "frame_total_median": float(med[3]), "frame_total_p90": float(np.percentile(rows[:, 3], 90) * 1e3)},
           "timed_work": {"per_frame": True, "gated_kp_per_crop": float(rows[:, 4].mean() / crops), "objects_accepted": float(rows[:, 5].mean()), "ba_edges": float(rows[:, 6].mean())},
           "e2e": {"value": 1e3 / med[3], "unit": "frames/s", "h2d_bytes_per_step": int(3 * H * W * 4 + crops * 16), "d2h_bytes_per_step": int(crops * NUM_KP * 7 * 4)},
           "gpu_launches": int(model.context().kernel_launches()), "clocks": clocks, "roofline": None,
           "note": "wall clock per frame on the host (time.perf_counter around the same calls the reference times with utils.device_time, lib/utils/utils.py:20-23); "
                   "8 crops per launch: the conv kernels run far below their batch-256 efficiency (launch floors, 1.7-wave tails)"}
    if not args.no_cpu_baseline:
        cfps, n, cores, _ = cpu_frames_per_s("latency", args.cpu_baseline_seconds)
        out["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": f"{n} timed frames through oracle/ on the host cores"}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------------------------
def c5_sequence(seed, n_views, crops, res):
    from suo_slam_b200 import synth
    return synth.make_slam_sequence(seed, n_views=n_views, n_obj=crops, res=res, n_sym=crops // 2, radius=synth.MARKER_RADIUS * res / 256)


def c5_view_args(seq, v):
    objs = seq["objs"]
    return (v["view_id"], v["img"], seq["K"], [d["obj_id"] for d in v["dets"]], np.stack([d["bbox"] for d in v["dets"]]),
            np.stack([o["model_kps"] for o in objs]), np.stack([o["model_kps_mask"] for o in objs]),
            np.array([o["is_symmetric"] for o in objs]), np.array([o["diameter"] for o in objs]))


def cpu_c5_frames_per_s(seconds, n_views=3):
    """The CPU restatement of a SLAM-mode view (oracle/slam_frame_oracle.py: two torch-CPU forwards at 512^2, PnP, vote, priors, curr_only LM)."""
    import torch
    from oracle import slam_frame_oracle as sfo
    from suo_slam_b200 import synth
    wl = WORKLOADS["c5"]
    sd = synth.make_marker_state_dict(0)
    seq = c5_sequence(1000, n_views, wl["crops"], wl["res"])
    cores = host_cores()
    torch.set_num_threads(cores)
    st, times = sfo.State(), []
    t_end = time.perf_counter() + seconds
    for v in seq["views"]:
        t0 = time.perf_counter()
        sfo.process_view(st, sd, *c5_view_args(seq, v), res=wl["res"], kp_var_thresh=0.5, bbox_thresh=1.0, manual_kp_std=0.1, init_with_outliers=True)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() > t_end and len(times) >= 2:
            break
    return 1.0 / float(np.median(times[1:] if len(times) > 1 else times)), len(times), cores


def run_native_c5(args):
    """configs[4]: T-LESS-shape SLAM-mode views — 16 crops of 512x512 per frame (128x128 heat-maps), half of the objects symmetric: every view is
    ONE suo_slam_frame call (forward on the 8 non-symmetric crops -> PnP -> camera-pose vote -> priors of the 8 symmetric crops rendered on the
    device -> second forward -> PnP -> object (re-)initialisation -> curr_only LM).  Views of a sequence depend on each other, so
    every GPU tracks its own sequence (replicas); a step is one view.  `value`: the views' packed inputs (image, map state, history) resident in
    HBM, recorded from a first pass of the tracker; `e2e`: the tracker itself (host numpy in / out, one call per view)."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from suo_slam_b200 import _lib, slam, synth
    from suo_slam_b200.pkpnet import PkpNet
    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    wl = WORKLOADS["c5"]
    crops, res, n_views = wl["crops"], wl["res"], wl["frames"]
    model = PkpNet(input_res=(res, res), max_crops=crops)
    model.load_state_dict(synth.make_marker_state_dict(0))
    model.cuda(local)
    ctx = model.context()
    lib, hdl, p = _lib.lib(), ctx.handle, _lib.ptr
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    seq = c5_sequence(7000 + rank, n_views, crops, res)
    tless = dict(kp_var_thresh=0.5, bbox_thresh=1.0, manual_kp_std=0.1, init_with_outliers=True)       # evaluate.py:68-76 (T-LESS)

    def track(record=None):
        trk = slam.SlamTracker(model, **tless)
        trk.record = record
        outs = [trk.process_view(*c5_view_args(seq, v)) for v in seq["views"]]
        return trk, outs
    rec = []
    trk, outs = track(rec)
    cam_err = [float(np.linalg.norm(trk.cam_poses[v["view_id"]][:, 3] - v["T_GtoC"][:3, 3])) for v in seq["views"] if v["view_id"] in trk.cam_poses]
    work = {"views": len(outs), "views_with_camera_pose": int(sum(o["cam_ok"] for o in outs)), "gated_kp_per_crop": float(np.mean([o["kp_used"].sum() / crops for o in outs])),
            "objects_with_pnp_pose_per_view": float(np.mean([(~np.all(np.isclose(o["T_pnp"], np.eye(4)), axis=(1, 2))).sum() for o in outs])),
            "symmetric_crops_given_priors_per_view": float(np.mean([o["prior_mask"].any(1).sum() for o in outs])),
            "curr_only_edges_per_view": float(np.mean([o["status"][3] for o in outs])), "curr_only_inlier_edges_per_view": float(np.mean([o["status"][5] for o in outs])),
            "objects_in_map": len(trk.obj_poses), "camera_translation_err_mm_median": float(np.median(cam_err)) if cam_err else None}
    if work["views_with_camera_pose"] < len(outs) or work["curr_only_edges_per_view"] < 20:
        raise SystemExit(f"bench c5: the SLAM views did no real work: {work}")
    # device-resident replay of the recorded calls
    K = NUM_KP
    dv = []
    for r in rec:
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        h = r["hist"]
        dv.append(dict(img=t(r["img"]), K=t(r["K"]), boxes=t(r["boxes"]), mk=t(r["mk"]), mm=t(r["mm"]), diam=t(r["diam"]), mv=t(r["map_valid"]), Tm=t(r["T_map"]),
                       L=r["L"], n1=r["n1"], n_views=r["n_views"], nh=0 if h is None else len(h["crop"]),
                       h=None if h is None else {k: t(v) for k, v in h.items()}))
    L = crops
    o = dict(cam=torch.zeros(12, dtype=torch.float64, device=dev), st=torch.zeros(8, dtype=torch.int32, device=dev), Tp=torch.zeros((L, 16), dtype=torch.float64, device=dev),
             used=torch.zeros((L, K), dtype=torch.uint8, device=dev), bain=torch.zeros((L, K), dtype=torch.uint8, device=dev), uv=torch.zeros((L, K, 2), device=dev),
             cov=torch.zeros((L, K, 4), device=dev), To=torch.zeros((L, 12), dtype=torch.float64, device=dev), mv=torch.zeros(L, dtype=torch.uint8, device=dev))

    def dev_view(i):
        d = dv[i % len(dv)]
        h = d["h"]
        ctx.check(lib.suo_slam_frame(hdl, p(d["img"]), H, W, p(d["K"]), p(d["boxes"]), d["L"], d["n1"], p(d["mk"]), p(d["mm"]), p(d["diam"]), p(d["mv"]), p(d["Tm"]), d["n_views"], d["nh"],
                                     *((p(h["crop"]), p(h["T"]), p(h["K"]), p(h["off"]), p(h["mk"]), p(h["uv"]), p(h["cov"])) if h is not None else (None,) * 7),
                                     tless["kp_var_thresh"], tless["bbox_thresh"], tless["manual_kp_std"], 1, 0, p(o["cam"]), p(o["st"]), p(o["Tp"]), p(o["used"]), p(o["bain"]),
                                     p(o["uv"]), p(o["cov"]), None, None, None, p(o["To"]), p(o["mv"]), None, None, None, 0, 1, sp))

    def timed(fn, steps, warm):
        fn(warm)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.kernel_launches()
        e0.record(stream); fn(steps); e1.record(stream)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ctx.kernel_launches() - l0
    steps = max(args.steps, n_views)

    def dev_steps(n):
        for i in range(n):
            dev_view(i)

    def e2e_steps(n):
        done = 0
        while done < n:
            t = slam.SlamTracker(model, **tless)
            for v in seq["views"][: n - done]:
                t.process_view(*c5_view_args(seq, v))
                done += 1
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    W_ = max(3, args.warmup)
    ms_dev, launches = timed(dev_steps, steps, W_)
    ms_e2e, _ = timed(e2e_steps, steps, n_views)
    clocks = sampler.stop() if rank == 0 else None
    ctx.check(lib.suo_check_range(hdl))
    if rank == 0:
        h2d = H * W * 3 + crops * (16 + 24 * K + K + 8 + 1 + 96) + 72
        out = {"metric": "frames/sec (16 obj-crops/frame, 640x480, SLAM mode)", "value": steps * world / (ms_dev * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": W_,
               "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_name(args),
               "data": "synthetic (fiducial-network weights, marker SLAM sequences)", "config": dict(frame_config("c5", 1, world), views_per_sequence=n_views, symmetric_objects=crops // 2,
                                                                                                     thresholds="T-LESS (evaluate.py:68-76)"),
               "engine": {"parallelism": f"{world} independent sequence(s), one per GPU (SLAM-mode views are sequentially dependent: replicas only, no collective)",
                          "call": "one suo_slam_frame per view: 2 dependent forwards (8 + 8 crops of 512x512), PnP x2, vote, device-rendered priors, init / re-init, curr_only LM"},
               "timed_work": work,
               "e2e": {"value": steps * world / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(crops * (128 + K * 26 + 72 + 96 + 10) + 128),
                       "ms_per_step": ms_e2e / steps, "how": "SlamTracker.process_view: host numpy in, suo_slam_frame with host pointers, host bookkeeping of the map and history"},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": None,
               "note": "latency workload: 8 crops of 512^2 per forward (= 32 crops of 256^2 of conv work), two forwards and ~20 small solver / glue kernels per view in sequence"}
        if not args.no_cpu_baseline and world == 1:
            cfps, n, cores = cpu_c5_frames_per_s(args.cpu_baseline_seconds)
            out["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port",
                                   "sample": f"{n} views of one sequence through oracle/slam_frame_oracle.py (median of the views after the first)"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "ba512":
        run_native_ba512(a)
    elif a.workload == "latency":
        run_native_latency(a)
    elif a.workload == "c5":
        run_native_c5(a)
    else:
        run_native_frames(a)
