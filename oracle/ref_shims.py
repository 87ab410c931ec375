"""TEST INFRASTRUCTURE ONLY — import shims so the read-only reference tree at
/root/reference can be imported in the BUILD container (it does not exist on
the GPU box).  Used only by oracle/gen_golden_*.py to produce the committed
fixtures under tests/golden/ and by tests that are skipped when the tree is
absent.  Nothing in the product path imports this.

Shims: numpy.int / numpy.math aliases (removed in NumPy>=1.24)
and a stub matplotlib whose pyplot.cm.get_cmap returns a callable, needed by
lib/labeling/kp_config.py:2,97-101 and lib/utils/utils.py:12.
"""
import math
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SUO_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "models"))


def install():
    if not hasattr(np, "int"):
        np.int = int  # noqa
    if not hasattr(np, "math"):
        np.math = math  # noqa
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")

        class _CM:
            @staticmethod
            def get_cmap(name):
                def f(x):
                    x = np.asarray(x, dtype=np.float64)
                    return np.stack([x, 1 - x, 0.5 * np.ones_like(x), np.ones_like(x)], -1)
                return f
        plt.cm = _CM()
        mpl.pyplot = plt
        mpl.cm = _CM()
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.cm"] = types.ModuleType("matplotlib.cm")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference_pkpnet():
    """Returns the reference's lib.models.pkpnet module (unmodified)."""
    install()
    import importlib
    return importlib.import_module("lib.models.pkpnet")


def import_reference_utils():
    install()
    import importlib
    return importlib.import_module("lib.utils.utils")
