"""molar_b200 — B200-native drop-in for MolAR's data-parallel hot path.

`molar_b200.api` mirrors the names of the reference's Python front end (pymolar:
molar_python/src/lib.rs:137-376, molar_python/python/pymolar/molar.pyi:130-213) on top of the
C ABI in include/molar_b200.h.  Importing the package does not need a GPU; calling into it does.
"""
from . import _capi  # noqa: F401
from . import comm  # noqa: F401
from .api import (System, Sel, PeriodicBox, IsometryTransform, Trajectory, distance_search, fit_transform,  # noqa: F401
                  rmsd, rmsd_py, rmsd_mw, MolarB200Error, probe_trajectory, load_trajectory)

__version__ = "0.1.0"
