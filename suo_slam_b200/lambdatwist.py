"""Module-level drop-in for the reference's pybind module ``lambdatwist``
(thirdparty/lambdatwist/pnp_python_binding.cpp:57-62): ``lambdatwist.pnp(xs_in, ys_in,
threshold=0.001) -> ndarray[4,4]`` (identity on failure, never raises).  Put this package's
directory on sys.path ahead of the reference build and
``import lambdatwist`` in lib/object_slam.py:10 resolves here."""
from .geometry import lambdatwist_pnp as pnp  # noqa: F401
