"""Module-level drop-in for the subset of the reference's pybind module ``g2o`` that
``ObjectSLAM.optimize()`` touches (lib/object_slam.py:706-896; bindings in
thirdparty/g2opy/python/core/sparse_optimizer.h:28-149, python/types/object_slam/
types_object_slam.h:17-52, python/types/slam3d/se3quat.h:18-72, python/core/robust_kernel.h:47-58).

The graph is only *recorded* in Python; every ``optimizer.optimize(n)`` packs the level-0 edges
into arrays and runs ONE ``suo_ba_batch`` launch (the g2o Levenberg loop of
optimization_algorithm_levenberg.cpp:58-150 on the GPU).  With this module on the path as
``g2o`` the reference's optimize() body runs unchanged for single-view, curr_only AND global graphs
(cameras and objects free, the CHOLMOD path of lib/object_slam.py:710: the library eliminates the
cameras by a Schur complement, csrc/ba_global.cu).

g2o quirk kept: after ``optimize()`` the active edges hold the error of the last EVALUATED LM trial —
after a rejected trial that is the rejected state's error, not the error at the returned estimate
(optimization_algorithm_levenberg.cpp:120-141) — and the reference reads it through ``e.chi2()``
without recomputing (lib/object_slam.py:881-883).  The kernel leaves exactly those errors behind
(``suo_ba_last_errors``) and ``optimize()`` copies them into the edges, so the reference's 4-round flow
through this module classifies inliers like the packed path (``suo_ba_batch`` with rounds) and like g2o.
"""
from __future__ import annotations

import numpy as np

from . import ba as _ba


class SE3Quat:
    def __init__(self, R=None, t=None):
        self._T = np.eye(4)
        if R is not None:
            self._T[:3, :3] = np.asarray(R, dtype=np.float64)
            self._T[:3, 3] = np.asarray(t, dtype=np.float64).ravel()

    def matrix(self):
        return self._T.copy()

    def rotation_matrix(self):
        return self._T[:3, :3].copy()

    def translation(self):
        return self._T[:3, 3].copy()


class VertexSE3Expmap:
    def __init__(self):
        self._id, self._est, self._fixed = -1, SE3Quat(), False

    def set_id(self, i):
        self._id = int(i)

    def id(self):
        return self._id

    def set_estimate(self, pose: SE3Quat):
        self._est = pose

    def estimate(self) -> SE3Quat:
        return self._est

    def set_fixed(self, f: bool):
        self._fixed = bool(f)

    def fixed(self):
        return self._fixed

    def set_marginalized(self, m):
        pass


class RobustKernelHuber:
    def __init__(self, delta=1.0):
        self.delta = float(delta)


class _Edge:
    def __init__(self, cam_k):
        self.cam_k = np.asarray(cam_k, dtype=np.float64).reshape(4)
        self._verts = {}
        self._uv = np.zeros(2)
        self._info = np.eye(2)
        self._kernel = None
        self._level = 0
        self._err = np.zeros(2)

    def set_vertex(self, i, v):
        self._verts[int(i)] = v

    def set_measurement(self, uv):
        self._uv = np.asarray(uv, dtype=np.float64).reshape(2)

    def set_information(self, info):
        self._info = np.asarray(info, dtype=np.float64).reshape(2, 2)

    def set_robust_kernel(self, k):
        self._kernel = k

    def set_level(self, level):
        self._level = int(level)

    def level(self):
        return self._level

    def chi2(self):
        return float(self._err @ self._info @ self._err)          # base_edge.h:56-59

    def _project(self, p_c):
        k = self.cam_k
        return np.array([k[0] * p_c[0] / p_c[2] + k[2], k[1] * p_c[1] / p_c[2] + k[3]])


class EdgeSE3ProjectFromObject(_Edge):
    """vertex 0 = object (T_wo), vertex 1 = camera (T_cw); types_object_slam.cpp:45-60."""

    def __init__(self, cam_k, p_inO):
        super().__init__(cam_k)
        self.p = np.asarray(p_inO, dtype=np.float64).reshape(3)

    def compute_error(self):
        Two, Tcw = self._verts[0].estimate().matrix(), self._verts[1].estimate().matrix()
        pw = Two[:3, :3] @ self.p + Two[:3, 3]
        self._err = self._uv - self._project(Tcw[:3, :3] @ pw + Tcw[:3, 3])


class EdgeSE3ProjectFromFixedObject(_Edge):
    """vertex 0 = camera; the object pose is folded into p_inG (types_object_slam.h:66-79)."""

    def __init__(self, cam_k, p_inO, T_OtoG):
        super().__init__(cam_k)
        T = np.asarray(T_OtoG, dtype=np.float64)
        self.p = T[:3, :3] @ np.asarray(p_inO, dtype=np.float64).reshape(3) + T[:3, 3]

    def compute_error(self):
        Tcw = self._verts[0].estimate().matrix()
        self._err = self._uv - self._project(Tcw[:3, :3] @ self.p + Tcw[:3, 3])


class LinearSolverDenseSE3:
    pass


class LinearSolverCholmodSE3:
    pass


class BlockSolverSE3:
    def __init__(self, linear_solver):
        self.linear_solver = linear_solver


class OptimizationAlgorithmLevenberg:
    def __init__(self, solver):
        self.solver = solver


class SparseOptimizer:
    def __init__(self):
        self._verts, self._edges, self._level = [], [], 0
        self.last_stats = None

    def set_algorithm(self, alg):
        self._alg = alg

    def add_vertex(self, v):
        self._verts.append(v)
        return True

    def add_edge(self, e):
        self._edges.append(e)
        return True

    def edges(self):
        return list(self._edges)

    def vertices(self):
        return {v.id(): v for v in self._verts}

    def set_verbose(self, v):
        pass

    def initialize_optimization(self, level=0):
        self._level = int(level)
        return True

    def optimize(self, iterations):
        active = [e for e in self._edges if e.level() == self._level]
        if not active or not self._verts:
            return -1
        verts = sorted(self._verts, key=lambda v: v.id())          # sparse_optimizer.cpp:493-498
        index = {id(v): i for i, v in enumerate(verts)}
        poses = np.stack([v.estimate().matrix()[:3, :4] for v in verts])
        fixed = np.array([v.fixed() for v in verts], np.uint8)
        e_obj, e_cam, cam_k, p, uv, info = [], [], [], [], [], []
        for e in active:
            if isinstance(e, EdgeSE3ProjectFromObject):
                o, c = index[id(e._verts[0])], index[id(e._verts[1])]
                e_obj.append(o)
                e_cam.append(c)
            else:
                e_obj.append(-1)
                e_cam.append(index[id(e._verts[0])])
            cam_k.append(e.cam_k); p.append(e.p); uv.append(e._uv); info.append(e._info.ravel())
        kernels = [e._kernel for e in active if e._kernel is not None]
        delta = kernels[0].delta if kernels else 1e150            # no kernel == Huber with an unreachable threshold
        P, _, stats, err = _ba.ba_batch([0, len(verts)], [0, len(active)], poses, fixed, e_obj, e_cam, cam_k, p, uv, info,
                                        np.ones(len(active)), [int(iterations)], huber_delta=delta, chi2_gate=1e300,
                                        init_with_outliers=True, return_errors=True)
        for v, T in zip(verts, P):
            if not v.fixed():
                v.set_estimate(SE3Quat(T[:, :3], T[:, 3]))
        ran = int(stats[0, 1]) > 0
        for e, r, o, c in zip(active, err, e_obj, e_cam):
            if ran and ((o >= 0 and not fixed[o]) or not fixed[c]):
                e._err = r.copy()                                  # the error g2o leaves in the edge (see module docstring)
            else:
                e.compute_error()                                  # edge not in the active set / no LM iteration ran
        self.last_stats = stats[0]
        return int(stats[0, 1])
