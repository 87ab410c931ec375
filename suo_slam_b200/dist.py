"""Multi-GPU plumbing: one process per GPU, crops (or whole frames) sharded
round-robin, weights replicated, and ONE all-gather of fixed-size per-crop result records
before any step that needs all objects (camera voting, joint graph).  The reference has no
inference-side distribution at all (only torch.nn.DataParallel for training,
lib/utils/training_utils.py:5-40); this is the B200 design for BASELINE configs 4/5.

torch.distributed is plumbing only: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

# float64 words per crop record: T_pnp 3x4, T_ba 3x4, n_used, n_ba_inliers, accepted, crop id
RECORD_WORDS = 12 + 12 + 4


def shard_round_robin(n_items: int, rank: int, world: int) -> np.ndarray:
    """Indices owned by `rank`: item i -> rank i mod world (BASELINE.json configs[3])."""
    return np.arange(rank, n_items, world)


def shard_frames(n_frames: int, rank: int, world: int) -> np.ndarray:
    """Frame-granular sharding (crops of one frame share a BA graph, so frames stay whole)."""
    return shard_round_robin(n_frames, rank, world)


def pack_records(ids, T_pnp, T_ba, kp_used, ba_inliers) -> torch.Tensor:
    L = len(ids)
    rec = np.zeros((L, RECORD_WORDS))
    rec[:, :12] = np.asarray(T_pnp)[:, :3, :].reshape(L, 12)
    rec[:, 12:24] = np.asarray(T_ba).reshape(L, 12)
    rec[:, 24] = np.asarray(kp_used).sum(-1)
    rec[:, 25] = np.asarray(ba_inliers).sum(-1)
    rec[:, 26] = 1.0 - np.all(np.isclose(np.asarray(T_pnp)[:, :3, :], np.eye(4)[:3]), axis=(1, 2))
    rec[:, 27] = np.asarray(ids)
    return torch.from_numpy(rec)


def allgather_records(rec: torch.Tensor, max_per_rank: int, group=None) -> torch.Tensor:
    """Single all-gather of [max_per_rank, RECORD_WORDS] per rank (ranks with fewer items pad with
    id = -1); returns the records of all ranks ordered by crop id."""
    world = dist.get_world_size(group)
    pad = torch.full((max_per_rank, RECORD_WORDS), -1.0, dtype=torch.float64, device=rec.device)
    pad[: rec.shape[0]] = rec
    out = torch.empty((world * max_per_rank, RECORD_WORDS), dtype=torch.float64, device=rec.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out[out[:, 27] >= 0]
    return out[torch.argsort(out[:, 27])]
