"""Checkpoint -> packed weights.

The reference saves ``{'model': state_dict, 'epoch': int, 'args': Namespace, ...}`` with torch.save
(train.py:173-181, state built in train.py's epoch loop) and ObjectSLAM.__init__ reloads it with
``torch.load(chkpt_path)``, ``PkpNet.load_state_dict(checkpoint['model'])``, ``checkpoint['args']``,
``checkpoint['epoch']`` (lib/object_slam.py:92-97).  On torch >= 2.6 that torch.load needs
``weights_only=False`` because of the pickled argparse.Namespace; training under
``DataParallelWrapper`` (lib/utils/training_utils.py:5-40) prefixes every key with ``module.``.

``load_checkpoint`` handles both and returns what ObjectSLAM reads; ``convert`` writes the BN-folded, packed blob
``suo_load_weights`` consumes (suo_slam_b200.weights.pack_state_dict) next to a small JSON header, so a deployment
needs neither torch pickles nor the folding step at start-up:

    python -m suo_slam_b200.checkpoint [--unsafe] results/.../model_best.pth.tar model_best.suo
"""
from __future__ import annotations

import json
import struct
import sys

import torch

from . import arch, weights

FILE_MAGIC = b"SUOW1\0\0\0"


def _strip_prefix(sd):
    if sd and all(k.startswith("module.") for k in sd):
        return {k[len("module."):]: v for k, v in sd.items()}
    return sd


def load_checkpoint(path: str, allow_pickle: bool = False):
    """-> (state_dict ready for PkpNet.load_state_dict, epoch, args) from a reference checkpoint file.

    The file is read with ``weights_only=True`` plus ``argparse.Namespace`` (the only non-tensor object ObjectSLAM needs,
    lib/object_slam.py:94-96) on the allow-list, so a third-party checkpoint cannot run code at load time.  Reference
    checkpoints written by train.py:349-355 also pickle the Adam optimizer OBJECT; those need ``allow_pickle=True``
    (full unpickling: only for files you trust)."""
    import argparse
    try:
        with torch.serialization.safe_globals([argparse.Namespace]):
            ck = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:
        if not allow_pickle:
            raise ValueError(f"{path}: cannot be read with weights_only=True ({type(e).__name__}: {str(e)[:200]}); if you trust the file, "
                             "pass allow_pickle=True / --unsafe (reference checkpoints pickle their optimizer object)") from e
        ck = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(ck, dict) or "model" not in ck:
        raise ValueError(f"{path}: not a SUO-SLAM checkpoint (expected a dict with a 'model' entry, train.py:173-181)")
    sd = _strip_prefix(dict(ck["model"]))
    missing = [k for k, _ in arch.state_dict_spec(arch.NUM_KP) if k not in sd]
    if missing:
        raise ValueError(f"{path}: state dict is missing {len(missing)} PkpNet keys, e.g. {missing[:3]}")
    return sd, int(ck.get("epoch", -1)), ck.get("args")


def convert(path_in: str, path_out: str, allow_pickle: bool = False) -> dict:
    """Reference checkpoint -> packed blob file.  Returns the JSON header that was written."""
    sd, epoch, args = load_checkpoint(path_in, allow_pickle)
    blob = weights.pack_state_dict(sd, arch.NUM_KP)
    meta = dict(epoch=epoch, num_kp=arch.NUM_KP, source=str(path_in), **weights.program_summary(blob),
                train_args={k: repr(v) for k, v in vars(args).items()} if hasattr(args, "__dict__") else None)
    head = json.dumps(meta).encode()
    with open(path_out, "wb") as f:
        f.write(FILE_MAGIC + struct.pack("<QQ", len(head), len(blob)) + head + blob)
    return meta


def load_packed(path: str):
    """-> (blob bytes for suo_load_weights / PkpNet.load_packed, JSON header)."""
    with open(path, "rb") as f:
        if f.read(8) != FILE_MAGIC:
            raise ValueError(f"{path}: not a packed SUO weights file")
        nh, nb = struct.unpack("<QQ", f.read(16))
        meta = json.loads(f.read(nh))
        blob = f.read(nb)
    if len(blob) != nb:
        raise ValueError(f"{path}: truncated")
    return blob, meta


if __name__ == "__main__":
    argv = [a for a in sys.argv[1:] if a != "--unsafe"]
    if len(argv) != 2:
        sys.exit(__doc__)
    print(json.dumps(convert(argv[0], argv[1], allow_pickle="--unsafe" in sys.argv), indent=1))
