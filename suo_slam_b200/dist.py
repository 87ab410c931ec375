"""Multi-GPU plumbing: one process per GPU, frames sharded round-robin, weights
replicated, and ONE all-gather of fixed-size per-crop result records before any step that needs all
objects (camera voting lib/object_slam.py:975-1072, joint graph :736-837).  The reference has no
inference-side distribution at all (only torch.nn.DataParallel for training,
lib/utils/training_utils.py:5-40); this is the B200 design for BASELINE configs 4/5.

The record (include/suo_b200.h, suo_record_bytes) is packed on the device by ``suo_pack_records`` and
exchanged by ``suo_allgather_results`` = ncclAllGather on the caller's stream through a communicator
this module creates with the NCCL library torch has already loaded.  torch.distributed is plumbing
only (rendezvous of the NCCL unique id; gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, arch


def shard_round_robin(n_items: int, rank: int, world: int) -> np.ndarray:
    """Indices owned by `rank`: item i -> rank i mod world (BASELINE.json configs[3])."""
    return np.arange(rank, n_items, world)


def shard_frames(n_frames: int, rank: int, world: int) -> np.ndarray:
    """Frame-granular sharding (crops of one frame share a BA graph, so frames stay whole)."""
    return shard_round_robin(n_frames, rank, world)


# ---- the record ---------------------------------------------------------------------------------
def record_bytes(num_kp: int = arch.NUM_KP) -> int:
    return 208 + 24 * num_kp + ((num_kp + 7) & ~7)


def record_dtype(num_kp: int = arch.NUM_KP) -> np.dtype:
    """numpy view of one record; layout documented at suo_pack_records (include/suo_b200.h)."""
    pad = ((num_kp + 7) & ~7) - num_kp
    return np.dtype([("T_pnp", "<f8", (3, 4)), ("T_ba", "<f8", (3, 4)), ("crop_id", "<i4"), ("accepted", "<i4"),
                     ("n_used", "<i4"), ("n_ba_inliers", "<i4"), ("uv", "<f4", (num_kp, 2)), ("cov", "<f4", (num_kp, 2, 2)),
                     ("flags", "u1", (num_kp,)), ("pad", "u1", (pad,))])


def pack_records_host(ids, T_pnp, T_ba, kp_used, ba_inliers, uv, cov, num_kp: int = arch.NUM_KP) -> np.ndarray:
    """Host (numpy) statement of the record layout — what ``suo_pack_records`` writes on the device; the GPU tests
    compare the two byte for byte, the CPU tests exchange these over gloo."""
    L = len(ids)
    rec = np.zeros(L, record_dtype(num_kp))
    T_pnp = np.asarray(T_pnp, np.float64).reshape(L, -1)[:, :12].reshape(L, 3, 4)
    rec["T_pnp"] = T_pnp
    rec["T_ba"] = np.asarray(T_ba, np.float64).reshape(L, 3, 4)
    rec["crop_id"] = np.asarray(ids)
    eye = np.eye(4)[:3]
    rec["accepted"] = ~np.all(np.abs(T_pnp - eye) <= 1e-8 + 1e-5 * np.abs(eye), axis=(1, 2))
    u, b = np.asarray(kp_used).astype(bool), np.asarray(ba_inliers).astype(bool)
    rec["n_used"], rec["n_ba_inliers"] = u.sum(-1), b.sum(-1)
    rec["uv"] = np.asarray(uv, np.float32).reshape(L, num_kp, 2)
    rec["cov"] = np.asarray(cov, np.float32).reshape(L, num_kp, 2, 2)
    rec["flags"] = u.astype(np.uint8) | (b.astype(np.uint8) << 1)
    return rec


def unpack_records(buf, num_kp: int = arch.NUM_KP) -> np.ndarray:
    """bytes / uint8 tensor of gathered records -> structured array without the padding records (crop_id < 0),
    ordered by crop id."""
    if isinstance(buf, torch.Tensor):
        buf = buf.detach().cpu().numpy()
    rec = np.frombuffer(np.ascontiguousarray(buf).tobytes(), dtype=record_dtype(num_kp))
    rec = rec[rec["crop_id"] >= 0]
    return rec[np.argsort(rec["crop_id"], kind="stable")]


def allgather_records_torch(rec: torch.Tensor, group=None) -> torch.Tensor:
    """Fallback / CPU-test exchange of a uint8 record tensor [n_local, record_bytes] through torch.distributed."""
    world = dist.get_world_size(group)
    out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=torch.uint8, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    return out


# ---- NCCL communicator for suo_allgather_results ------------------------------------------------
class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


class NcclComm:
    """An ncclComm_t over the ranks of the default torch.distributed group, created with the libnccl.so.2 that torch
    loaded (so ``suo_allgather_results`` and torch share one NCCL).  The unique id travels through torch.distributed."""

    def __init__(self, device: int):
        self.lib = C.CDLL("libnccl.so.2")
        self.lib.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        self.lib.ncclCommDestroy.argtypes = [C.c_void_p]
        self.lib.ncclGetErrorString.restype = C.c_char_p
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = _UniqueId()
        if rank == 0:
            self._check(self.lib.ncclGetUniqueId(C.byref(uid)))
        t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=torch.device("cuda", device))
        dist.broadcast(t, 0)
        C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        self.comm = C.c_void_p()
        torch.cuda.set_device(device)
        self._check(self.lib.ncclCommInitRank(C.byref(self.comm), world, uid, rank))
        self.world = world

    def _check(self, r):
        if r != 0:
            raise _lib.SuoError(f"NCCL error {r}: {self.lib.ncclGetErrorString(r).decode()}")

    @property
    def handle(self):
        return self.comm

    def close(self):
        if self.comm:
            self.lib.ncclCommDestroy(self.comm)
            self.comm = C.c_void_p()


class RecordExchange:
    """Per-step exchange for one rank: pack on the device, one all-gather (native NCCL when available, else
    torch.distributed), all on the caller's stream; no host work between the frame path and the collective."""

    def __init__(self, ctx: _lib.Context, n_local: int, world: int, device: int, native: bool = True):
        self.ctx, self.n_local, self.world = ctx, n_local, world
        self.rb = int(_lib.lib().suo_record_bytes(ctx.num_kp))
        dev = torch.device("cuda", device)
        self.rec = torch.zeros((n_local, self.rb), dtype=torch.uint8, device=dev)
        self.out = torch.zeros((world * n_local, self.rb), dtype=torch.uint8, device=dev)
        self.comm = None
        self.how = "none (1 rank)"
        if world > 1:
            self.how = "torch.distributed all_gather_into_tensor (NCCL)"
            if native:
                try:
                    self.comm = NcclComm(device)
                    self.how = "suo_allgather_results (ncclAllGather on the frame path's stream)"
                except Exception as e:          # a missing libnccl symbol must not take the job down: torch's NCCL still works
                    self.how += f" [native communicator unavailable: {e!r}]"

    def run(self, id_base, T_pnp, T_ba, kp_used, ba_inliers, uv, cov, stream_ptr):
        p = _lib.ptr
        self.ctx.check(_lib.lib().suo_pack_records(self.ctx.handle, None, int(id_base), p(T_pnp), p(T_ba), p(kp_used), p(ba_inliers),
                                                   p(uv), p(cov), self.n_local, p(self.rec), 1, stream_ptr))
        return self.gather(stream_ptr)

    def gather(self, stream_ptr):
        """All-gather of self.rec (already packed on the device, e.g. by suo_frames_u8_submit) into self.out."""
        p = _lib.ptr
        if self.world == 1:
            return self.rec
        if self.comm is not None:
            self.ctx.check(_lib.lib().suo_allgather_results(self.ctx.handle, self.comm.handle, p(self.rec), self.rb, self.n_local,
                                                            p(self.out), stream_ptr))
        else:
            dist.all_gather_into_tensor(self.out, self.rec)
        return self.out
