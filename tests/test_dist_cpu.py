"""CPU: the N>1 host-side path (frame sharding + the single all-gather of pose records) with
gloo, world_size 2 — the same code bench.py runs over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from suo_slam_b200 import dist as sdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, crops, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = sdist.shard_frames(n_frames, rank, world)
    ids = np.concatenate([np.arange(f * crops, (f + 1) * crops) for f in frames]) if len(frames) else np.zeros(0, int)
    L = len(ids)
    T_pnp = np.tile(np.eye(4), (L, 1, 1))
    T_pnp[:, 0, 3] = ids                          # recognisable payload
    T_ba = np.tile(np.eye(4)[:3], (L, 1, 1))
    T_ba[:, 2, 3] = 2.0 * ids
    used = np.ones((L, 41), bool)
    rec = sdist.pack_records(ids, T_pnp, T_ba, used, used)
    max_per_rank = -(-n_frames // world) * crops
    allr = sdist.allgather_records(rec, max_per_rank)
    q.put((rank, allr.numpy()))
    dist.destroy_process_group()


def test_round_robin_sharding_is_a_partition():
    for n, w in [(64, 8), (7, 2), (3, 4), (0, 2)]:
        parts = [sdist.shard_round_robin(n, r, w) for r in range(w)]
        allidx = np.sort(np.concatenate(parts)) if n else np.zeros(0, int)
        assert np.array_equal(allidx, np.arange(n))
        assert all(np.all(p % w == r) for r, p in enumerate(parts))


def test_allgather_records_world2():
    world, n_frames, crops = 2, 5, 8          # uneven: rank 0 gets 3 frames, rank 1 gets 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, crops, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        a = res[r]
        assert a.shape == (n_frames * crops, sdist.RECORD_WORDS)
        assert np.array_equal(a[:, 27], np.arange(n_frames * crops))      # every crop exactly once, ordered
        assert np.array_equal(a[:, 3], np.arange(n_frames * crops))       # T_pnp[0,3] payload
        assert np.array_equal(a[:, 12 + 11], 2.0 * np.arange(n_frames * crops))
        assert np.all(a[:, 24] == 41)
    assert np.array_equal(res[0], res[1])
