"""Run the UNMODIFIED reference on top of libsuo_b200: ``python -m suo_slam_b200.dropin evaluate.py <evaluate.py's own arguments>``.

The reference reaches its hot path through three imports: ``import g2o`` and ``import lambdatwist`` (lib/object_slam.py:9-10,
the two pybind modules) and ``from .models.pkpnet import PkpNet`` (lib/object_slam.py:14).  ``install()`` puts this package's drop-ins under
exactly those module names *before* the reference is imported — ``sys.modules['g2o']``, ``sys.modules['lambdatwist']`` and
``sys.modules['lib.models.pkpnet']`` — so neither ``evaluate.py`` nor anything under the reference's ``lib/`` is edited; everything else of the
reference (datasets, ObjectSLAM's bookkeeping, metrics) runs as it is.

Two environment shims the reference needs on current NumPy / PyTorch, independent of this library (SURVEY.md §8b "runs unchanged" checklist), are
installed only with ``compat=True`` (the command line does): ``numpy.int`` / ``numpy.math`` aliases (evaluate.py:390, lib/utils) and
``torch.load`` defaulting to ``weights_only=False`` (lib/object_slam.py:92 loads a checkpoint that pickles an argparse.Namespace)."""
from __future__ import annotations

import sys
import types


def install(compat: bool = False):
    from . import g2o, lambdatwist, pkpnet
    sys.modules["g2o"] = g2o
    sys.modules["lambdatwist"] = lambdatwist
    m = types.ModuleType("lib.models.pkpnet")
    m.__doc__ = "suo_slam_b200 drop-in for the reference's lib/models/pkpnet.py (PkpNet only: the forward runs in libsuo_b200)"
    m.PkpNet = pkpnet.PkpNet
    sys.modules["lib.models.pkpnet"] = m
    if compat:
        import math
        import numpy as np
        if not hasattr(np, "int"):
            np.int = int
        if not hasattr(np, "math"):
            np.math = math
        import torch
        if not getattr(torch.load, "_suo_compat", False):
            _load = torch.load

            def load(*a, **kw):
                kw.setdefault("weights_only", False)
                return _load(*a, **kw)
            load._suo_compat = True
            torch.load = load
    return m


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    import os
    import runpy
    install(compat=True)
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    runpy.run_path(script, run_name="__main__")
