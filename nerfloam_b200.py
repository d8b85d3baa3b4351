"""Importable alias of the package directory `nerf-loam_b200/` (a hyphen is not a valid identifier).

    import nerfloam_b200 as nl
    nl.svo.Octree(), nl.grid.svo_intersect(...), nl.render_helpers.bundle_adjust_frames(...)
"""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("nerf-loam_b200")
sys.modules[__name__] = _pkg
