"""Lattice systems and symmetry operations (drop-in for the M-index part of pydrex.geometry)."""

from enum import Enum, unique

import numpy as np


@unique
class LatticeSystem(Enum):
    """Crystallographic lattice systems; value = (a, b) of Grimmer (1979) (geometry.py:11-34)."""

    triclinic = (1, 1)
    monoclinic = (2, 2)
    orthorhombic = (2, 4)
    rhombohedral = (3, 6)
    tetragonal = (4, 8)
    hexagonal = (6, 12)


def _quat(axis, angle):
    half = 0.5 * angle
    return np.array([*(np.sin(half) * np.asarray(axis, dtype=float)), np.cos(half)])


def symmetry_operations(system):
    """Symmetry operations for `system` (geometry.py:121-182): scalar-last quaternions for
    rotations, 4x4 matrices for reflections."""
    axes = ([0, 0, 1], [0, 1, 0], [1, 0, 0])
    ident = np.array([0.0, 0.0, 0.0, 1.0])
    if system == LatticeSystem.triclinic:
        return [ident]
    if system in (LatticeSystem.monoclinic, LatticeSystem.orthorhombic):
        rotations = [_quat(v, np.pi) for v in axes]
        reflections = [np.diag(x) for x in ([1, -1, -1, 1], [1, -1, 1, -1], [1, 1, -1, -1])]
        return [ident, *rotations, *reflections]
    if system == LatticeSystem.rhombohedral:
        return [ident, *[_quat(v, i * np.pi / 3) for v in axes for i in (1, 2)]]
    if system == LatticeSystem.tetragonal:
        return [ident, *[_quat(v, i * np.pi / 2) for v in axes for i in (1, 2, 3)]]
    if system == LatticeSystem.hexagonal:
        r3 = [_quat(v, i * np.pi / 3) for v in axes for i in (1, 2)]
        r6 = [_quat(v, i * np.pi / 6) for v in axes for i in (1, 3, 5)]
        return [ident, *r3, *r6]
    raise ValueError(f"unsupported lattice system: {system}")
