"""PnP entry points with the reference's call surface.

``pnp(points_3d, points_2d, camera_matrix)`` is lib/object_slam.py:25-41 verbatim in
behaviour (None on < 4 points or on the identity "failure" pose; returns
(T[3,4], ones(n, bool))).  Underneath, instead of one pybind call into Lambda-Twist per
object, a whole batch of objects goes to the GPU in one ``suo_pnp_batch`` launch.
"""
from __future__ import annotations

import numpy as np

from . import _lib, runtime

DEFAULT_THRESHOLD = 0.001   # thirdparty/lambdatwist/pnp_python_binding.cpp:61
_call_counter = [0]


def pnp_batch(xs_list, ys_list, threshold: float = DEFAULT_THRESHOLD, seed: int = 0, obj_keys=None, ctx=None,
              return_stats: bool = False):
    """xs_list[i] [n_i,3], ys_list[i] [n_i,2] (pinhole-normalised). Returns T [n_obj,4,4]
    (identity = failure, as lambdatwist.pnp does)."""
    ctx = ctx or runtime.get_context()
    n_obj = len(xs_list)
    offs = np.zeros(n_obj + 1, np.int32)
    offs[1:] = np.cumsum([len(x) for x in xs_list])
    xs = np.ascontiguousarray(np.concatenate([np.asarray(x, np.float64).reshape(-1, 3) for x in xs_list]))
    ys = np.ascontiguousarray(np.concatenate([np.asarray(y, np.float64).reshape(-1, 2) for y in ys_list]))
    T = np.zeros((n_obj, 4, 4))
    stats = np.zeros((n_obj, 5), np.int32)
    keys = None if obj_keys is None else np.ascontiguousarray(obj_keys, np.uint64)
    ctx.check(_lib.lib().suo_pnp_batch(ctx.handle, _lib.ptr(xs), _lib.ptr(ys), _lib.ptr(offs), n_obj, float(threshold),
                                        int(seed), _lib.ptr(keys), _lib.ptr(T), _lib.ptr(stats), 0, None))
    return (T, stats) if return_stats else T


def lambdatwist_pnp(xs_in, ys_in, threshold: float = DEFAULT_THRESHOLD):
    """``lambdatwist.pnp`` (pnp_python_binding.cpp:57-62): one object, returns 4x4; never raises."""
    xs_in, ys_in = np.asarray(xs_in, np.float64), np.asarray(ys_in, np.float64)
    if xs_in.shape[0] < 4:
        return np.eye(4)
    # the reference's RANSAC stream advances from call to call (process-global engine); mirror that
    _call_counter[0] += 1
    return pnp_batch([xs_in], [ys_in], threshold, seed=0, obj_keys=[_call_counter[0]])[0]


def pnp(points_3d, points_2d, camera_matrix):
    assert points_3d.shape[0] == points_2d.shape[0], 'points 3D and points 2D must have same number of rows'
    assert camera_matrix.shape == (3, 3), "Camera matrix must be of shape (3,3)"
    num_pts = points_3d.shape[0]
    if num_pts < 4:
        return None
    KinvT = np.linalg.inv(camera_matrix).T
    points_2d_norm = points_2d @ KinvT[:2, :2] + KinvT[2:3, :2]
    res = lambdatwist_pnp(points_3d, points_2d_norm)
    if np.allclose(res, np.eye(4)):
        return None
    return res[:3, :], np.ones((num_pts), dtype=bool)
