"""Misorientation statistics (drop-in for the M-index part of pydrex.stats)."""

import ctypes

import numpy as np

from . import _lib
from . import geometry as _geo


def _max_misorientation(system):
    """Maximum misorientation angle for the lattice system (stats.py:203-215)."""
    s = _geo.LatticeSystem
    if system in (s.orthorhombic, s.rhombohedral):
        return 120
    if system in (s.tetragonal, s.hexagonal):
        return 90
    if system in (s.triclinic, s.monoclinic):
        return 180
    raise ValueError(f"unsupported lattice system: {system}")


def misorientations_random(low, high, system):
    """Expected density of misorientation angles in (low, high) degrees for a random
    texture (stats.py:130-200); closed form evaluated by the library's host code."""
    lib = _lib.load()
    max_t = _max_misorientation(system)
    M, N = system.value
    if not 0 <= low <= high <= max_t:
        raise ValueError(
            f"bounds must obey `low` ∈ [0, `high`) and `high` < {max_t}.\n"
            + f"You've supplied (`low`, `high`) = ({low}, {high})."
        )
    out = ctypes.c_double()
    rc = lib.drex_misorientations_random_host(float(low), float(high), int(M), int(N),
                                              ctypes.byref(out))
    if rc == _lib.DREX_E_UNSUPPORTED:
        raise AssertionError(_lib.last_error())  # reference: `assert False` (stats.py:198)
    _lib.check(rc)
    return out.value


def misorientation_hist(orientations, system, bins=None):
    """Misorientation-angle histogram (stats.py:83-127): density over 1-degree bins.

    As in the reference, `bins` is accepted and ignored (stats.py:127 passes θmax).
    Returns `(hist, bin_edges)` like numpy.histogram(..., density=True).
    """
    from .diagnostics import _mindex_device

    _, counts = _mindex_device(np.asarray(orientations)[None], system)
    counts = counts[0].astype(np.float64)
    tmax = _max_misorientation(system)
    return counts / counts.sum(), np.linspace(0.0, tmax, tmax + 1)
