"""TEST INFRASTRUCTURE ONLY — ctypes bindings of oracle/libgeom_oracle.so (the
CPU FP64 restatement in geom_oracle.cpp) and, when built, of the reference's
own P3P/P4P in oracle/_ref/libref_p4p.so.  Never imported by suo_slam_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build():
    subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libgeom_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_p3p.argtypes = [_dp, _dp, _dp, _dp]
        L.orc_p3p.restype = C.c_int
        L.orc_p4p.argtypes = [_dp, _dp, _ip, _dp]
        L.orc_sample4.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, _ip]
        L.orc_get_iterations.argtypes = [C.c_double]
        L.orc_get_iterations.restype = C.c_int
        L.orc_pnp.argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_uint64, C.c_uint64, C.c_int, _dp, _ip]
        L.orc_pnp_refine.argtypes = [_dp, _dp, C.c_int, C.c_double, _dp, _ip]
        L.orc_ba_optimize.argtypes = [C.c_int, _dp, _bp, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _bp, _ip, C.c_int,
                                      C.c_double, C.c_double, C.c_int, _ip]
        L.orc_ba_optimize_err.argtypes = L.orc_ba_optimize.argtypes + [_dp]
        L.orc_ba_optimize_err.restype = None
        L.orc_edge_eval.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_se3_oplus.argtypes = [_dp, _dp, _dp]
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own p4p/p3p (None if oracle/_ref was not built)."""
    global _REF
    if _REF is None:
        path = os.path.join(HERE, "_ref", "libref_p4p.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_p4p.argtypes = [_dp, _dp, C.c_int, _ip, _dp]
        L.ref_p3p.argtypes = [_dp, _dp, _dp, _dp]
        L.ref_p3p.restype = C.c_int
        _REF = L
    return _REF


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


def p3p(y2d, x3d, use_ref=False):
    Rs, Ts = np.zeros((4, 3, 3)), np.zeros((4, 3))
    L = ref_lib() if use_ref else lib()
    n = (L.ref_p3p if use_ref else L.orc_p3p)(_c(y2d), _c(x3d), Rs, Ts)
    return Rs[:n], Ts[:n]


def p4p(xs, ys, idx4, use_ref=False):
    T = np.zeros((4, 4))
    if use_ref:
        ref_lib().ref_p4p(_c(xs), _c(ys), len(xs), _c(idx4, np.int32), T)
    else:
        lib().orc_p4p(_c(xs), _c(ys), _c(idx4, np.int32), T)
    return T


def sample4(seed, obj_key, it, n):
    idx = np.zeros(4, np.int32)
    lib().orc_sample4(seed, obj_key, it, n, idx)
    return idx


def lambdatwist_pnp(xs, ys, threshold=0.001, seed=0, obj_key=0, refine=True):
    """Restated ``lambdatwist.pnp`` (pnp_python_binding.cpp:57-62): 4x4, identity = failure."""
    T = np.zeros((4, 4))
    stats = np.zeros(5, np.int32)
    lib().orc_pnp(_c(xs), _c(ys), len(xs), threshold, seed, obj_key, int(refine), T, stats)
    return T, dict(best_inliers=int(stats[0]), best_iter=int(stats[1]), total_iters=int(stats[2]),
                   refine_iters=(int(stats[3]), int(stats[4])))


def pnp(points_3d, points_2d, camera_matrix, seed=0, obj_key=0):
    """Restated ``pnp()`` of lib/object_slam.py:25-41."""
    assert points_3d.shape[0] == points_2d.shape[0]
    assert camera_matrix.shape == (3, 3)
    n = points_3d.shape[0]
    if n < 4:
        return None
    KinvT = np.linalg.inv(camera_matrix).T
    p2n = points_2d @ KinvT[:2, :2] + KinvT[2:3, :2]
    res, _ = lambdatwist_pnp(points_3d, p2n, seed=seed, obj_key=obj_key)
    if np.allclose(res, np.eye(4)):
        return None
    return res[:3, :], np.ones(n, dtype=bool)


def pnp_refine(xs, ys, T, threshold=0.001):
    T = _c(T).copy()
    it = np.zeros(2, np.int32)
    lib().orc_pnp_refine(_c(xs), _c(ys), len(xs), threshold, T, it)
    return T, it


def ba_optimize(poses, fixed, e_obj, e_cam, cam_k, p, uv, info, inliers, its, huber_delta=np.sqrt(5.991),
                chi2_gate=5.991, init_with_outliers=False):
    """One ObjectSLAM.optimize() solve over a packed graph (see orc_ba_optimize)."""
    poses = _c(poses).reshape(-1, 12).copy()
    inl = _c(inliers, np.uint8).copy()
    stats = np.zeros(3, np.int32)
    its = _c(its, np.int32)
    lib().orc_ba_optimize(len(poses), poses, _c(fixed, np.uint8), len(e_cam), _c(e_obj, np.int32), _c(e_cam, np.int32),
                          _c(cam_k).reshape(-1, 4), _c(p).reshape(-1, 3), _c(uv).reshape(-1, 2),
                          _c(info).reshape(-1, 4), inl, its, len(its), float(huber_delta), float(chi2_gate),
                          int(init_with_outliers), stats)
    return poses.reshape(-1, 3, 4), inl.astype(bool), dict(rounds=int(stats[0]), outer=int(stats[1]), trials=int(stats[2]))


def ba_optimize_err(poses, fixed, e_obj, e_cam, cam_k, p, uv, info, inliers, its, huber_delta=np.sqrt(5.991),
                    chi2_gate=5.991, init_with_outliers=False):
    """ba_optimize that also returns the error vector every edge is left holding (orc_ba_optimize_err): what e.chi2() reads next."""
    poses = _c(poses).reshape(-1, 12).copy()
    inl = _c(inliers, np.uint8).copy()
    stats = np.zeros(3, np.int32)
    its = _c(its, np.int32)
    err = np.zeros((len(e_cam), 2))
    lib().orc_ba_optimize_err(len(poses), poses, _c(fixed, np.uint8), len(e_cam), _c(e_obj, np.int32), _c(e_cam, np.int32),
                              _c(cam_k).reshape(-1, 4), _c(p).reshape(-1, 3), _c(uv).reshape(-1, 2), _c(info).reshape(-1, 4), inl, its,
                              len(its), float(huber_delta), float(chi2_gate), int(init_with_outliers), stats, err)
    return poses.reshape(-1, 3, 4), inl.astype(bool), dict(rounds=int(stats[0]), outer=int(stats[1]), trials=int(stats[2])), err


def edge_eval(T_obj, T_cam, cam_k, p, uv):
    err, Ji, Jj = np.zeros(2), np.zeros((2, 6)), np.zeros((2, 6))
    lib().orc_edge_eval(_c(T_obj).reshape(12), _c(T_cam).reshape(12), _c(cam_k), _c(p), _c(uv), err, Ji, Jj)
    return err, Ji, Jj


def se3_oplus(T, upd):
    out = np.zeros(12)
    lib().orc_se3_oplus(_c(T).reshape(12), _c(upd), out)
    return out.reshape(3, 4)
