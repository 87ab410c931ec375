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

    python -m suo_slam_b200.checkpoint results/.../model_best.pth.tar model_best.suo
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


def load_checkpoint(path: str):
    """-> (state_dict ready for PkpNet.load_state_dict, epoch, args) from a reference checkpoint file."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(ck, dict) or "model" not in ck:
        raise ValueError(f"{path}: not a SUO-SLAM checkpoint (expected a dict with a 'model' entry, train.py:173-181)")
    sd = _strip_prefix(dict(ck["model"]))
    missing = [k for k, _ in arch.state_dict_spec(arch.NUM_KP) if k not in sd]
    if missing:
        raise ValueError(f"{path}: state dict is missing {len(missing)} PkpNet keys, e.g. {missing[:3]}")
    return sd, int(ck.get("epoch", -1)), ck.get("args")


def convert(path_in: str, path_out: str) -> dict:
    """Reference checkpoint -> packed blob file.  Returns the JSON header that was written."""
    sd, epoch, args = load_checkpoint(path_in)
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
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    print(json.dumps(convert(sys.argv[1], sys.argv[2]), indent=1))
