"""Seeded synthetic inputs for every BASELINE.json config.

No dataset or checkpoint ships with the reference (README.md:63-84 are Google
Drive links), so weights, frames and solver problems are generated here.  All
generators are pure numpy / torch-CPU and deterministic in ``seed``; the same
functions feed the CUDA path, the oracle and the golden-fixture scripts.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

from . import arch

# YCB-Video intrinsics used by the reference's data
K_YCBV = np.array([[1066.778, 0.0, 312.9869], [0.0, 1067.487, 241.3109], [0.0, 0.0, 1.0]])


def make_synthetic_state_dict(seed: int = 0, num_kp: int = arch.NUM_KP, peaky: float = 1.0):
    """A state dict with the reference's key names/shapes (arch.state_dict_spec).

    Each tensor is drawn from its own generator keyed by (seed, crc32(key)), so
    the values do not depend on construction order or on torch's module-init
    code.  Convs: U(-b, b) with b = sqrt(2 / fan_in) (measured: final logits of O(1-10),
    so 59 residual blocks neither explode nor vanish); BN running stats / affine
    are randomised so that BN folding is exercised.
    ``peaky`` scales the last tmpOut conv so heat-maps get an unambiguous peak.
    """
    sd = {}
    for key, shape in arch.state_dict_spec(num_kp):
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
        leaf = key.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            sd[key] = torch.tensor(100, dtype=torch.long)
            continue
        is_bn = len(shape) == 1 and (".bn" in key or key.startswith("backbone.bn1.")
                                     or (".lin_." in key and key.split(".")[3] == "1"))
        if is_bn:
            if leaf == "running_mean":
                t = torch.randn(shape, generator=g) * 0.1
            elif leaf == "running_var":
                t = torch.rand(shape, generator=g) + 0.5
            elif leaf == "weight":
                t = torch.rand(shape, generator=g) + 0.5
            else:  # bias
                t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "weight":
            fan_in = int(np.prod(shape[1:]))
            b = (2.0 / fan_in) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        else:  # conv / linear bias
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        sd[key] = t.to(torch.float32)
    last = f"backbone.tmpOut.{arch.N_STACK - 1}"
    sd[last + ".weight"] = sd[last + ".weight"] * peaky
    sd[last + ".bias"] = sd[last + ".bias"] * peaky
    return sd


def random_rotation(rng: np.random.Generator) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def so3_exp(w: np.ndarray) -> np.ndarray:
    th = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + W
    return np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th ** 2 * (W @ W)


def fix_K_for_bbox_ndc(K: np.ndarray, bbox) -> np.ndarray:
    """K_bbox = S·T·K: camera matrix projecting into the bbox's NDC square
    [-1,1]² with *negative* fy (reference lib/utils/utils.py:416-429).
    A float32 bbox — what the reference's data loader hands over (lib/datasets/bop.py:540-552) — keeps the reference's float32
    scalar arithmetic: ``w = x2 - x1`` is a float32 subtraction and ``2.0 / w`` a float32 division (NumPy >= 2: the Python
    float is weakly typed; NumPy 1.x divided in float64 — the 6e-8 this moves the scale by is the reference's own
    environment dependence, and this repo follows the NumPy it is tested with); everything after that is float64."""
    b = np.asarray(bbox)
    if b.dtype == np.float32:
        w32, h32 = np.float32(b[2] - b[0]), np.float32(b[3] - b[1])
        sx, sy = float(np.float32(2.0) / w32), float(np.float32(-2.0) / h32)
    else:
        sx, sy = 2.0 / (float(b[2]) - float(b[0])), -2.0 / (float(b[3]) - float(b[1]))
    x1, y1 = float(b[0]), float(b[1])
    T = np.eye(3)
    T[0, 2], T[1, 2] = -x1, -y1
    S = np.eye(3)
    S[0, :] *= sx
    S[1, :] *= sy
    S[0, 2] -= 1.0
    S[1, 2] += 1.0
    return S @ T @ np.asarray(K, dtype=np.float64)


def make_frame(seed: int, n_obj: int = 8, H: int = 480, W: int = 640, num_kp: int = arch.NUM_KP,
               K: np.ndarray = K_YCBV, uv_noise: float = 0.01, outlier_frac: float = 0.1):
    """One synthetic YCBV-shape frame (SURVEY.md §8d "Frame")."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    objs = []
    for o in range(n_obj):
        model_kps = rng.uniform(-60, 60, size=(num_kp, 3))
        n_valid = int(rng.integers(8, 23))
        mask = np.zeros(num_kp, dtype=bool)
        mask[rng.choice(num_kp, size=n_valid, replace=False)] = True
        for _ in range(50):
            R = random_rotation(rng)
            t = np.array([rng.uniform(-150, 150), rng.uniform(-100, 100), rng.uniform(600, 1200)])
            pc = model_kps @ R.T + t
            px = pc @ K.T
            px = px[:, :2] / px[:, 2:3]
            lo, hi = px[mask].min(0), px[mask].max(0)
            ctr, half = (lo + hi) / 2, (hi - lo) / 2 * 1.1
            box = np.array([ctr[0] - half[0], ctr[1] - half[1], ctr[0] + half[0], ctr[1] + half[1]])
            box[[0, 2]] = np.clip(box[[0, 2]], 0, W - 1)
            box[[1, 3]] = np.clip(box[[1, 3]], 0, H - 1)
            if box[2] - box[0] >= 10 and box[3] - box[1] >= 10:
                break
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, t
        Kb = fix_K_for_bbox_ndc(K, box)
        uv = pc @ Kb.T
        uv = uv[:, :2] / uv[:, 2:3]
        uv_meas = uv + rng.normal(scale=uv_noise, size=uv.shape)
        n_out = int(round(outlier_frac * n_valid))
        if n_out:
            idx = rng.choice(np.nonzero(mask)[0], size=n_out, replace=False)
            uv_meas[idx] = rng.uniform(-0.9, 0.9, size=(n_out, 2))
        sig = rng.uniform(0.005, 0.05, size=(num_kp, 2))
        rho = rng.uniform(-0.5, 0.5, size=num_kp)
        cov = np.zeros((num_kp, 2, 2))
        cov[:, 0, 0], cov[:, 1, 1] = sig[:, 0] ** 2, sig[:, 1] ** 2
        cov[:, 0, 1] = cov[:, 1, 0] = rho * sig[:, 0] * sig[:, 1]
        objs.append(dict(T_OtoC=T, model_kps=model_kps, model_kps_mask=mask, bbox=box.astype(np.float32),
                         K_bbox=Kb, uv_gt=uv, uv_meas=uv_meas, cov=cov, diameter=150.0))
    return dict(img=img, K=K.copy(), objs=objs)


def make_ba_problem(seed: int, n_obj: int = 512, n_kp: int = 12, noise_px: float = 1.0,
                    outlier_frac: float = 0.0):
    """BASELINE config 3: objects x keypoints seen by one fixed camera, object
    poses perturbed by exp(N(0,(5°,5°,5°,10,10,20mm))) (SURVEY.md §8d "C3";
    generator modelled on thirdparty/g2opy/python/examples/object_slam_demo.py:54-150,
    cam_k=[320,320,320,240])."""
    rng = np.random.default_rng(seed)
    cam_k = np.array([320.0, 320.0, 320.0, 240.0])
    p_O = rng.uniform(-60, 60, size=(n_obj, n_kp, 3))
    T_gt = np.zeros((n_obj, 3, 4))
    T_init = np.zeros((n_obj, 3, 4))
    uv = np.zeros((n_obj, n_kp, 2))
    info = np.zeros((n_obj, n_kp, 2, 2))
    for o in range(n_obj):
        R = random_rotation(rng)
        t = np.array([rng.uniform(-200, 200), rng.uniform(-150, 150), rng.uniform(600, 1200)])
        T_gt[o, :, :3], T_gt[o, :, 3] = R, t
        pc = p_O[o] @ R.T + t
        uv[o, :, 0] = cam_k[0] * pc[:, 0] / pc[:, 2] + cam_k[2]
        uv[o, :, 1] = cam_k[1] * pc[:, 1] / pc[:, 2] + cam_k[3]
        uv[o] += rng.normal(scale=noise_px, size=(n_kp, 2))
        n_out = int(round(outlier_frac * n_kp))
        if n_out:
            idx = rng.choice(n_kp, size=n_out, replace=False)
            uv[o, idx] += rng.uniform(-60, 60, size=(n_out, 2))
        sig = rng.uniform(0.5, 2.0, size=(n_kp, 2)) * noise_px
        rho = rng.uniform(-0.5, 0.5, size=n_kp)
        cov = np.zeros((n_kp, 2, 2))
        cov[:, 0, 0], cov[:, 1, 1] = sig[:, 0] ** 2, sig[:, 1] ** 2
        cov[:, 0, 1] = cov[:, 1, 0] = rho * sig[:, 0] * sig[:, 1]
        info[o] = np.linalg.inv(cov)
        dw = np.deg2rad(rng.normal(scale=5.0, size=3))
        dt = rng.normal(size=3) * np.array([10.0, 10.0, 20.0])
        dR = so3_exp(dw)
        T_init[o, :, :3] = R @ dR          # perturb in the object frame: T_gt * exp(delta)
        T_init[o, :, 3] = R @ dt + t
    return dict(cam_k=cam_k, p_O=p_O, uv=uv, info=info, T_gt=T_gt, T_init=T_init)


def make_global_graph(seed: int, n_views: int = 12, n_obj: int = 6, kp_range=(8, 16), see_prob: float = 0.7,
                      noise_px: float = 1.0, outlier_frac: float = 0.05, perturb: float = 1.0):
    """A multi-view object-SLAM graph as ObjectSLAM.optimize(curr_only=False) builds it in SLAM / SfM mode
    (lib/object_slam.py:736-837): object vertices 0..n_obj-1 (T_wo), then one camera vertex per view (T_cw,
    the first one fixed, :771), one EdgeSE3ProjectFromObject per detected keypoint, information = inv(cov).
    Scene modelled on thirdparty/g2opy/python/examples/object_slam_demo.py:54-150 (objects on a table, the
    camera circling it); pixel-unit intrinsics of YCB-V.  Returns the packed arrays of suo_ba_batch."""
    rng = np.random.default_rng(seed)
    cam_k = np.array([1066.778, 1067.487, 312.9869, 241.3109])
    T_wo = np.zeros((n_obj, 3, 4))
    kps = []
    for o in range(n_obj):
        T_wo[o, :, :3] = random_rotation(rng)
        T_wo[o, :, 3] = [rng.uniform(-250, 250), rng.uniform(-250, 250), rng.uniform(-40, 40)]
        kps.append(rng.uniform(-60, 60, size=(int(rng.integers(kp_range[0], kp_range[1] + 1)), 3)))
    T_cw = np.zeros((n_views, 3, 4))
    for v in range(n_views):
        ang = 2 * np.pi * v / n_views + rng.normal(scale=0.05)
        c = np.array([900 * np.cos(ang), 900 * np.sin(ang), rng.uniform(350, 600)])      # camera centre in the world
        z = -c / np.linalg.norm(c)                                                        # looks at the origin
        x = np.cross([0, 0, 1.0], z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z])                                                           # rows = camera axes in the world
        T_cw[v, :, :3], T_cw[v, :, 3] = R, -R @ c
    e_obj, e_cam, p, uv, info = [], [], [], [], []
    for v in range(n_views):
        seen = [o for o in range(n_obj) if rng.random() < see_prob] or [int(rng.integers(n_obj))]
        for o in seen:
            pw = kps[o] @ T_wo[o, :, :3].T + T_wo[o, :, 3]
            pc = pw @ T_cw[v, :, :3].T + T_cw[v, :, 3]
            m = np.c_[cam_k[0] * pc[:, 0] / pc[:, 2] + cam_k[2], cam_k[1] * pc[:, 1] / pc[:, 2] + cam_k[3]]
            m += rng.normal(scale=noise_px, size=m.shape)
            out = rng.random(len(m)) < outlier_frac
            m[out] += rng.uniform(-80, 80, size=(int(out.sum()), 2))
            sig = rng.uniform(0.5, 2.0, size=(len(m), 2)) * max(noise_px, 1e-3)
            rho = rng.uniform(-0.5, 0.5, size=len(m))
            for k in range(len(m)):
                cov = np.array([[sig[k, 0] ** 2, rho[k] * sig[k, 0] * sig[k, 1]], [rho[k] * sig[k, 0] * sig[k, 1], sig[k, 1] ** 2]])
                e_obj.append(o); e_cam.append(n_obj + v); p.append(kps[o][k]); uv.append(m[k]); info.append(np.linalg.inv(cov).ravel())
    def jitter(T, first_exact):
        out = T.copy()
        for i in range(len(T)):
            if first_exact and i == 0:
                continue
            dR = so3_exp(np.deg2rad(rng.normal(scale=2.0 * perturb, size=3)))
            out[i, :, :3] = dR @ T[i, :, :3]
            out[i, :, 3] = dR @ T[i, :, 3] + rng.normal(scale=8.0 * perturb, size=3)
        return out
    poses = np.concatenate([jitter(T_wo, False), jitter(T_cw, True)], 0)
    fixed = np.zeros(n_obj + n_views, np.uint8)
    fixed[n_obj] = 1
    n_e = len(e_obj)
    return dict(poses=poses, poses_gt=np.concatenate([T_wo, T_cw], 0), fixed=fixed, e_obj=np.asarray(e_obj, np.int32),
                e_cam=np.asarray(e_cam, np.int32), cam_k=np.tile(cam_k, (n_e, 1)), p=np.asarray(p), uv=np.asarray(uv),
                info=np.asarray(info), n_obj=n_obj, n_views=n_views)


def make_slam_scene(seed: int, n_views: int = 6, n_obj: int = 6, kp_range=(8, 16), noise_ndc: float = 0.01,
                    bad_pnp=(1,), bad_estimate=(2,), with_cov: bool = True):
    """SLAM-mode state as ObjectSLAM keeps it (lib/object_slam.py:100-123): ``obj_poses`` {obj: T_OtoG [4,4]},
    ``cam_poses`` {view: T_GtoC [4,4]}, ``detections`` {view: {obj: {pose, model_kp, K, uv_pred, cov_pred, inliers,
    bbox}}} with bbox-NDC keypoints (utils.fix_K_for_bbox_ndc) and the reference's debug noise (:1131).  Objects in
    ``bad_pnp`` get a wrong PnP pose in the LAST view (a bad camera-pose vote); objects in ``bad_estimate`` get a wrong
    map pose (candidates for re-initialisation).  The last view is the "current" one."""
    rng = np.random.default_rng(seed)
    g = make_global_graph(seed, n_views, n_obj, kp_range=kp_range, see_prob=1.0, noise_px=0.0, outlier_frac=0.0, perturb=0.0)
    T_wo, T_cw = g["poses_gt"][:n_obj], g["poses_gt"][n_obj:]
    kps = [g["p"][(g["e_obj"] == o) & (g["e_cam"] == n_obj)] for o in range(n_obj)]
    to44 = lambda T: np.vstack([T, [0, 0, 0, 1.0]])
    obj_poses, cam_poses, detections = {}, {}, {}
    for o in range(n_obj):
        T = to44(T_wo[o])
        if o in bad_estimate:
            T[:3, 3] += [60.0, -40.0, 30.0]
            T[:3, :3] = so3_exp(np.array([0.3, -0.2, 0.25])) @ T[:3, :3]
        obj_poses[10 + o] = T
    for v in range(n_views):
        cam_poses[100 + v] = to44(T_cw[v])
        detections[100 + v] = {}
        for o in range(n_obj):
            T_oc = to44(T_cw[v]) @ to44(T_wo[o])
            pc = kps[o] @ T_oc[:3, :3].T + T_oc[:3, 3]
            px = np.c_[K_YCBV[0, 0] * pc[:, 0] / pc[:, 2] + K_YCBV[0, 2], K_YCBV[1, 1] * pc[:, 1] / pc[:, 2] + K_YCBV[1, 2]]
            lo, hi = px.min(0), px.max(0)
            c, half = 0.5 * (lo + hi), 0.55 * np.maximum(hi - lo, 10.0)
            bbox = np.array([c[0] - half[0], c[1] - half[1], c[0] + half[0], c[1] + half[1]])
            Kb = fix_K_for_bbox_ndc(K_YCBV, bbox)
            h = pc @ Kb.T
            uv = (h[:, :2] / h[:, 2:3] + rng.normal(scale=noise_ndc, size=(len(pc), 2))).astype(np.float32)
            out = rng.random(len(uv)) < 0.1
            uv[out] += rng.uniform(-0.5, 0.5, size=(int(out.sum()), 2)).astype(np.float32)
            sig = rng.uniform(0.7, 1.5, size=(len(uv), 2)) * noise_ndc
            rho = rng.uniform(-0.4, 0.4, size=len(uv))
            cov = np.zeros((len(uv), 2, 2), np.float32)
            cov[:, 0, 0], cov[:, 1, 1] = sig[:, 0] ** 2, sig[:, 1] ** 2
            cov[:, 0, 1] = cov[:, 1, 0] = rho * sig[:, 0] * sig[:, 1]
            pose = T_oc.copy()
            pose[:3, 3] += rng.normal(scale=1.0, size=3)
            if v == n_views - 1 and o in bad_pnp:
                pose[:3, 3] += [80.0, 50.0, -120.0]
                pose[:3, :3] = so3_exp(np.array([-0.4, 0.3, 0.2])) @ pose[:3, :3]
            detections[100 + v][10 + o] = dict(pose=pose, model_kp=kps[o].copy(), K=Kb, uv_pred=uv, cov_pred=cov if with_cov else None,
                                               inliers=~out | (rng.random(len(uv)) < 0.3), bbox=bbox)
    return dict(obj_poses=obj_poses, cam_poses=cam_poses, detections=detections, view_ids=[100 + v for v in range(n_views)])


# ---- fiducial ("marker") network + frames: synthetic weights that really place keypoints ----------------------------
# No checkpoint ships with the reference (README.md:63-84), and a random-init network gates almost every keypoint out
# (lib/object_slam.py:1100-1115), so PnP / BA would run on empty inputs.  The generator below builds a state dict with
# the reference's exact architecture and key names in which 41 trunk channels carry a colour-marker detector through the
# skip connections of all 59 bottlenecks (every other weight stays seeded-random, every conv runs on dense data), and
# frames whose objects carry one coloured disc per model keypoint.  The frame path then produces gated keypoints, PnP
# consensus sets and non-empty BA graphs on both the CUDA path and the CPU oracle.
MARKER_T = 0.481            # detector threshold on the colour projection (own colour 0.5, closest other colour 0.462)
MARKER_GAIN = 1600.0        # tmpOut gain: peak logit ~ 30 over a background of O(1)
MARKER_RADIUS = 10.0        # disc radius in crop pixels (256x256 crop)


def marker_codes(num_kp: int = arch.NUM_KP) -> np.ndarray:
    """[num_kp, 3] unit colour directions: a golden-angle lattice on the part of the sphere at least 107 degrees away
    from black (zero padding at the crop border must not look like a marker); colour k = 0.5 + 0.5 * d_k."""
    i = np.arange(num_kp) + 0.5
    z = -0.3 + 1.3 * i / num_kp
    phi = i * np.pi * (3.0 - np.sqrt(5.0))
    r = np.sqrt(1.0 - z * z)
    a = np.ones(3) / np.sqrt(3.0)
    b = np.cross(a, [0.0, 0.0, 1.0])
    b /= np.linalg.norm(b)
    c = np.cross(a, b)
    return (r * np.cos(phi))[:, None] * b + (r * np.sin(phi))[:, None] * c + z[:, None] * a


def marker_colors_u8(num_kp: int = arch.NUM_KP) -> np.ndarray:
    return np.clip(np.rint(255.0 * (0.5 + 0.5 * marker_codes(num_kp))), 0, 255).astype(np.uint8)


def make_marker_state_dict(seed: int = 0, num_kp: int = arch.NUM_KP, noise: float = 0.15):
    """Reference-format state dict (arch.state_dict_spec) of the fiducial network.

    Trunk channels 0..num_kp-1 ("signal") are wired by hand, everything else is make_synthetic_state_dict(seed):
      * stem conv1_ (hg.py:67): channel k = 6x6 box filter (taps -2..+3: centred half a pixel to the right/below, which
        makes the stride-2 stem + 2x2 max-pool sample positions agree with the NDC pixel centres of pkpnet.py:19-26) of
        the colour projection (rgb - 0.5) . d_k, minus the threshold; bn1 = identity; ReLU keeps what exceeds it;
      * every bottleneck (Residual.py:20-35): conv3 rows of the signal channels are zero, so the skip path carries the
        signal unchanged (conv4 = identity on them where Cin != Cout); the branches still READ the signal channels;
      * the first bottleneck of each outer hourglass' low path (hg.py:42-44) computes -x on the signal channels, so the
        low-resolution pyramid carries no signal (its nearest-upsampled max-pool plateaus would swamp the sub-pixel peak);
      * lin_ = identity (+ identity BN) on the signal channels, tmpOut = MARKER_GAIN on the diagonal plus random weights
        on the other 215 features (an O(1) background texture), ll_ / tmpOut_ feed nothing back into the signal channels;
      * classifier (pkpnet.py:74-78): bias +2, small random weights -> kp_mask ~ 0.88 (gating is left to the covariance
        and bbox tests of lib/object_slam.py:1100-1115).
    """
    sd = make_synthetic_state_dict(seed, num_kp)
    S = num_kp
    d = torch.tensor(marker_codes(num_kp), dtype=torch.float32)

    def bn_identity(p, n):
        sd[p + ".weight"][:n] = 1.0
        sd[p + ".bias"][:n] = 0.0
        sd[p + ".running_mean"][:n] = 0.0
        sd[p + ".running_var"][:n] = 1.0

    w = sd["backbone.conv1_.weight"]
    w[:S] = 0.0
    w[:S, :3, 1:7, 1:7] = (d / 36.0)[:, :, None, None]
    sd["backbone.conv1_.bias"][:S] = -0.5 * d.sum(1) - MARKER_T
    bn_identity("backbone.bn1", S)
    eye = torch.arange(S)
    prefixes = [k[:-len(".conv3.weight")] for k in sd if k.endswith(".conv3.weight")]
    for p in prefixes:
        sd[p + ".conv3.weight"][:S] = 0.0
        sd[p + ".conv3.bias"][:S] = 0.0
        if p + ".conv4.weight" in sd:
            sd[p + ".conv4.weight"][:S] = 0.0
            sd[p + ".conv4.weight"][eye, eye, 0, 0] = 1.0
            sd[p + ".conv4.bias"][:S] = 0.0
    for i in range(arch.N_STACK):
        p = f"backbone.hourglass.{i}.low1_.0"
        bn_identity(p + ".bn", S)
        sd[p + ".conv1.weight"][:S] = 0.0
        sd[p + ".conv1.weight"][eye, eye, 0, 0] = 1.0
        sd[p + ".conv1.bias"][:S] = 0.0
        bn_identity(p + ".bn1", S)
        sd[p + ".conv2.weight"][:S] = 0.0
        sd[p + ".conv2.weight"][eye, eye, 1, 1] = 1.0
        sd[p + ".conv2.bias"][:S] = 0.0
        bn_identity(p + ".bn2", S)
        sd[p + ".conv3.weight"][eye, eye, 0, 0] = -1.0
        p = f"backbone.lin_.{i}"
        sd[p + ".0.weight"][:S] = 0.0
        sd[p + ".0.weight"][eye, eye, 0, 0] = 1.0
        sd[p + ".0.bias"][:S] = 0.0
        bn_identity(p + ".1", S)
        p = f"backbone.tmpOut.{i}"
        sd[p + ".weight"] *= noise
        sd[p + ".weight"][:, :S] = 0.0
        sd[p + ".weight"][eye, eye, 0, 0] = MARKER_GAIN
        sd[p + ".bias"][:] = 0.0
    for i in range(arch.N_STACK - 1):
        sd[f"backbone.ll_.{i}.weight"][:S] = 0.0
        sd[f"backbone.ll_.{i}.bias"][:S] = 0.0
        sd[f"backbone.tmpOut_.{i}.weight"][:S] = 0.0
        sd[f"backbone.tmpOut_.{i}.bias"][:S] = 0.0
    sd["classifier.2.weight"] *= 0.1
    sd["classifier.2.bias"][:] = 2.0
    return sd


def make_marker_frame(seed: int, n_obj: int = 8, H: int = 480, W: int = 640, num_kp: int = arch.NUM_KP,
                      K: np.ndarray = K_YCBV, res: int = 256, bg=(96, 160), radius: float = MARKER_RADIUS):
    """make_frame's geometry with an image the fiducial network can read: low-contrast noise background and, per object,
    one solid disc of colour k per valid model keypoint k.  The disc of a keypoint whose bbox-NDC position is (u, v) is
    painted where the network's soft-argmax reads (u, v): the reference's mesh grid is transposed (uv_x runs along heat-map
    ROWS, uv_y = -r[col]; lib/models/pkpnet.py:19-26,44-49), so inside the bbox the disc sits at the transposed position
    (column fraction (1 - v) / 2, row fraction (u + 1) / 2).  Discs are circles of `radius` crop pixels (ellipses in the frame),
    painted in object order — later objects occlude earlier ones and same-colour discs of other objects that fall into a
    crop act as the gross outliers PnP has to reject."""
    fr = make_frame(seed, n_obj, H, W, num_kp, K)
    fr["img"] = paint_markers(seed, fr["objs"], H, W, num_kp, res, bg, radius)
    return fr


def paint_markers(seed, objs, H, W, num_kp=arch.NUM_KP, res=256, bg=(96, 160), radius=MARKER_RADIUS):
    """The u8 image of make_marker_frame for a list of objects (dicts with bbox, model_kps_mask, uv_gt)."""
    rng = np.random.default_rng(seed ^ 0x5EED)
    img = rng.integers(bg[0], bg[1], size=(H, W, 3)).astype(np.float32)
    col = marker_colors_u8(num_kp).astype(np.float32)
    ss = 4
    sub = (np.arange(ss) + 0.5) / ss
    for o in objs:
        x1, y1, x2, y2 = [float(v) for v in o["bbox"]]
        bw, bh = x2 - x1, y2 - y1
        rx, ry = radius * bw / res, radius * bh / res
        for k in np.nonzero(o["model_kps_mask"])[0]:
            u, v = o["uv_gt"][k]
            cx, cy = x1 + 0.5 * (1.0 - v) * bw, y1 + 0.5 * (u + 1.0) * bh
            ix0, ix1 = int(np.floor(cx - rx)), int(np.ceil(cx + rx)) + 2
            iy0, iy1 = int(np.floor(cy - ry)), int(np.ceil(cy + ry)) + 2
            ix0, iy0, ix1, iy1 = max(ix0, 0), max(iy0, 0), min(ix1, W), min(iy1, H)
            if ix0 >= ix1 or iy0 >= iy1:
                continue
            # sample positions: pixel i is centred on the integer coordinate i (the convention of K and of roi_align's
            # aligned=False bilinear taps, lib/models/pkpnet.py:93), i.e. covers [i - 0.5, i + 0.5)
            xs = (np.arange(ix0, ix1)[:, None] - 0.5 + sub[None, :]).ravel()
            ys = (np.arange(iy0, iy1)[:, None] - 0.5 + sub[None, :]).ravel()
            inside = (((xs[None, :] - cx) / rx) ** 2 + ((ys[:, None] - cy) / ry) ** 2) <= 1.0
            alpha = inside.reshape(iy1 - iy0, ss, ix1 - ix0, ss).mean((1, 3)).astype(np.float32)
            patch = img[iy0:iy1, ix0:ix1]
            patch += alpha[:, :, None] * (col[k] - patch)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_slam_sequence(seed: int, n_views: int = 4, n_obj: int = 6, H: int = 480, W: int = 640, num_kp: int = arch.NUM_KP,
                       K: np.ndarray = K_YCBV, res: int = 256, n_sym=None, radius: float = MARKER_RADIUS):
    """A SLAM-mode sequence with IMAGES (BASELINE configs[4] / SURVEY.md §8d "C5"): n_obj objects on a table (world poses T_OtoG,
    per-object keypoint subsets as in make_frame), the camera circling them (scene modelled on
    thirdparty/g2opy/python/examples/object_slam_demo.py:54-150); the first n_sym objects (default half) are flagged symmetric
    (mesh_db[obj]["is_symmetric"], lib/object_slam.py:343).  Every view is a marker frame (paint_markers) with tight bboxes inflated
    10 % (bop.py:551).  The world frame is the FIRST camera frame, as ObjectSLAM defines it (first cam pose = identity, :410-411)."""
    rng = np.random.default_rng(seed)
    n_sym = n_obj // 2 if n_sym is None else n_sym
    objs = []
    for o in range(n_obj):
        mask = np.zeros(num_kp, dtype=bool)
        mask[rng.choice(num_kp, size=int(rng.integers(8, 23)), replace=False)] = True
        T = np.eye(4)
        T[:3, :3] = random_rotation(rng)
        T[:3, 3] = [rng.uniform(-170, 170), rng.uniform(-170, 170), rng.uniform(-30, 30)]
        objs.append(dict(obj_id=10 + o, T_wo=T, model_kps=rng.uniform(-60, 60, size=(num_kp, 3)), model_kps_mask=mask, diameter=150.0,
                         is_symmetric=o < n_sym))
    cams = []
    for v in range(n_views):
        ang = 0.25 * v + rng.normal(scale=0.02)
        c = np.array([1000 * np.cos(ang), 1000 * np.sin(ang), rng.uniform(450, 550)])      # camera centre in the table frame
        z = -c / np.linalg.norm(c)
        x = np.cross([0, 0, 1.0], z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        T = np.eye(4)
        T[:3, :3] = np.stack([x, y, z]); T[:3, 3] = -T[:3, :3] @ c
        cams.append(T)
    W0_inv = np.linalg.inv(cams[0])                       # table -> first camera = world
    for o in objs:
        o["T_OtoG"] = cams[0] @ o["T_wo"]
    views = []
    for v in range(n_views):
        T_GtoC = cams[v] @ W0_inv
        dets = []
        for o in objs:
            T_oc = T_GtoC @ o["T_OtoG"]
            pc = o["model_kps"] @ T_oc[:3, :3].T + T_oc[:3, 3]
            px = pc @ K.T
            px = px[:, :2] / px[:, 2:3]
            m = o["model_kps_mask"]
            lo, hi = px[m].min(0), px[m].max(0)
            ctr, half = (lo + hi) / 2, (hi - lo) / 2 * 1.1
            box = np.array([ctr[0] - half[0], ctr[1] - half[1], ctr[0] + half[0], ctr[1] + half[1]])
            box[[0, 2]] = np.clip(box[[0, 2]], 0, W - 1)
            box[[1, 3]] = np.clip(box[[1, 3]], 0, H - 1)
            Kb = fix_K_for_bbox_ndc(K, box)
            uv = pc @ Kb.T
            dets.append(dict(obj_id=o["obj_id"], bbox=box.astype(np.float32), model_kps_mask=m, uv_gt=uv[:, :2] / uv[:, 2:3], T_OtoC=T_oc))
        views.append(dict(view_id=100 + v, T_GtoC=T_GtoC, dets=dets, img=paint_markers(seed * 131 + v, dets, H, W, num_kp, res, radius=radius)))
    return dict(K=K.copy(), objs=objs, views=views)


def make_pnp_benchmark(seed: int, n: int = 250, pixel_sigma: float = 0.5, outlier_ratio: float = 0.5):
    """One experiment of the reference's PnP Monte-Carlo benchmark (thirdparty/lambdatwist/simulator.h:46-94,
    PointCloudWithNoisyMeasurements; points from getRandomPointsInfrontOfCamera :29-43): pose = (uniform rotation, unit
    translation), points at uniform normalised image positions in [-1,1]^2 and depths U(0.1, 100), measurement noise =
    a random unit 2-vector times sigma (pixel_sigma * 0.001, i.e. f = 1000), and outlier_ratio * n draws (with
    replacement, "exact ratio is not needed") moved at least 0.002 + 3 sigma away from the true projection.
    Returns xs [n,3], yns [n,2], Pcw [4,4]."""
    rng = np.random.default_rng(seed)
    unit = lambda d: (lambda v: v / np.linalg.norm(v))(rng.normal(size=d))
    q = unit(4)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    t = unit(3)
    yn = rng.uniform(-1, 1, size=(n, 2))
    dist = rng.uniform(0.1, 100, size=n)
    xc = np.c_[yn, np.ones(n)] * dist[:, None]
    xs = (xc - t) @ R                                  # Pwc * xc = R^T (xc - t)
    gt = yn.copy()
    sigma = pixel_sigma * 0.001
    yns = gt + np.stack([unit(2) for _ in range(n)]) * sigma
    for _ in range(int(n * outlier_ratio)):
        i = int(rng.integers(0, n))
        v = yns[i].copy()
        if rng.uniform() > 0.5:
            v = v + unit(2)
        while np.linalg.norm(gt[i] - v) < 0.002 + pixel_sigma * 0.001 * 3:
            v = v + unit(2) * 0.1 * rng.uniform(3, 10)
        yns[i] = v
    Pcw = np.eye(4)
    Pcw[:3, :3], Pcw[:3, 3] = R, t
    return xs, yns, Pcw


def pnp_benchmark_error(T_est: np.ndarray, Pcw: np.ndarray) -> float:
    """|angle| + |translation| of T_est * Pcw^-1 (thirdparty/lambdatwist/test_pnp.cpp:99-101); > 0.05 counts as a failure (:106)."""
    I = T_est @ np.linalg.inv(Pcw)
    c = np.clip((np.trace(I[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)
    return float(abs(np.arccos(c)) + np.linalg.norm(I[:3, 3]))
