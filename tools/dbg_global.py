import numpy as np
from oracle import geom
from suo_slam_b200 import ba, synth

def args(g): return (g["poses"], g["fixed"], g["e_obj"], g["e_cam"], g["cam_k"], g["p"], g["uv"], g["info"], np.ones(len(g["e_obj"])))
rng = np.random.default_rng(0)
for shuffle in (0, 1):
    for unary in (0, 1):
        for seed, V, N in ((50, 8, 4), (51, 15, 7)):
            g = synth.make_global_graph(seed, V, N)
            if shuffle:
                perm = rng.permutation(len(g["e_obj"]))
                for k in ("e_obj", "e_cam", "cam_k", "p", "uv", "info"): g[k] = g[k][perm]
            if unary:
                un = g["e_obj"] == 0
                T0 = g["poses"][0]
                g["p"][un] = g["p"][un] @ T0[:, :3].T + T0[:, 3]
                g["e_obj"][un] = -1
            for its in ([10, 10, 20, 20], [30]):
                Po, io, so = geom.ba_optimize(*args(g), its, init_with_outliers=True)
                P, inl, st = ba.ba_batch([0, len(g["poses"])], [0, len(g["e_obj"])], *args(g), its, init_with_outliers=True)
                print(f"shuffle={shuffle} unary={unary} seed={seed} its={its}: stats gpu {tuple(int(x) for x in st[0])} oracle {(so['rounds'], so['outer'], so['trials'])} "
                      f"inl mismatch {int((inl != io).sum())} max|dP| {np.abs(P - Po).max():.2e}")
