"""CPU: the N>1 host-side path (frame sharding + the single all-gather of the per-crop result records: poses, flags,
keypoints, covariances — SURVEY.md §5) with gloo, world_size 2.  On GPUs the records are packed by suo_pack_records and
exchanged by suo_allgather_results (tests/test_gpu_geom.py compares the device packing with pack_records_host)."""
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


def _payload(ids, K=41):
    """Recognisable per-crop results: every field of the record is a function of the crop id."""
    L = len(ids)
    rng = np.random.default_rng(7)
    T_pnp = np.tile(np.eye(4), (L, 1, 1))
    T_pnp[:, 0, 3] = ids
    T_pnp[ids % 5 == 0] = np.eye(4)               # PnP failure = identity (lib/object_slam.py:38-39)
    T_ba = np.tile(np.eye(4)[:3], (L, 1, 1))
    T_ba[:, 2, 3] = 2.0 * ids
    used = (np.arange(K)[None, :] + ids[:, None]) % 3 == 0
    bain = used & ((np.arange(K)[None, :] + ids[:, None]) % 2 == 0)
    uv = (ids[:, None, None] + rng.random((1, K, 2))).astype(np.float32)
    cov = (0.001 * ids[:, None, None, None] + rng.random((1, K, 2, 2))).astype(np.float32)
    return T_pnp, T_ba, used, bain, uv, cov


def _worker(rank, world, port, n_frames, crops, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = sdist.shard_frames(n_frames, rank, world)
    ids = np.concatenate([np.arange(f * crops, (f + 1) * crops) for f in frames]) if len(frames) else np.zeros(0, int)
    rec = sdist.pack_records_host(ids, *_payload(ids))
    max_per_rank = -(-n_frames // world) * crops                  # ranks with fewer frames pad with crop_id = -1
    padded = np.zeros(max_per_rank, rec.dtype)
    padded["crop_id"] = -1
    padded[: len(rec)] = rec
    t = torch.from_numpy(padded.view(np.uint8).reshape(max_per_rank, -1).copy())
    assert t.shape[1] == sdist.record_bytes() == 1240
    allr = sdist.allgather_records_torch(t)
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
    ids = np.arange(n_frames * crops)
    T_pnp, T_ba, used, bain, uv, cov = _payload(ids)
    for r in range(world):
        rec = sdist.unpack_records(res[r])
        assert np.array_equal(rec["crop_id"], ids)                          # every crop exactly once, ordered, padding dropped
        np.testing.assert_array_equal(rec["T_pnp"], T_pnp[:, :3])
        np.testing.assert_array_equal(rec["T_ba"], T_ba)
        assert np.array_equal(rec["accepted"], (ids % 5 != 0).astype(np.int32))
        assert np.array_equal(rec["n_used"], used.sum(1)) and np.array_equal(rec["n_ba_inliers"], bain.sum(1))
        np.testing.assert_array_equal(rec["uv"], uv)
        np.testing.assert_array_equal(rec["cov"], cov)
        assert np.array_equal(rec["flags"] & 1, used.astype(np.uint8)) and np.array_equal(rec["flags"] >> 1, bain.astype(np.uint8))
    assert np.array_equal(res[0], res[1])


def test_record_layout_matches_the_header():
    """record_dtype is the layout include/suo_b200.h documents for suo_pack_records (offsets in bytes)."""
    dt = sdist.record_dtype(41)
    assert dt.itemsize == sdist.record_bytes(41) == 1240
    off = {k: dt.fields[k][1] for k in dt.names}
    assert (off["T_pnp"], off["T_ba"], off["crop_id"], off["accepted"], off["n_used"], off["n_ba_inliers"]) == (0, 96, 192, 196, 200, 204)
    assert (off["uv"], off["cov"], off["flags"]) == (208, 208 + 8 * 41, 208 + 24 * 41)
    assert sdist.record_bytes(8) == 208 + 24 * 8 + 8 and sdist.record_bytes(30) == 208 + 24 * 30 + 32
