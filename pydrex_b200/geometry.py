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


def misorientation_angles(q1_array, q2_array):
    """Minimum misorientation angles for collections of rotation quaternions
    (geometry.py:75-118): q1_array [N, A, 4], q2_array [N, B, 4] -> [N] angles in degrees, the
    smallest over all A x B pairs.  Evaluated in the precision of the inputs (float32 or
    float64) like the reference, returned as float64.  The pair table of the reference is
    never materialised."""
    from . import _lib

    torch = _lib.require_cuda()
    lib = _lib.load()
    is_torch = type(q1_array).__module__.startswith("torch")
    if q1_array.shape[0] != q2_array.shape[0]:
        raise ValueError(
            "the first dimensions of q1_array and q2_array must be of equal length"
            + f" but {q1_array.shape[0]} != {q2_array.shape[0]}"
        )
    dev = torch.device("cuda", torch.cuda.current_device())
    a = q1_array if is_torch else torch.as_tensor(np.ascontiguousarray(q1_array))
    b = q2_array if type(q2_array).__module__.startswith("torch") else torch.as_tensor(
        np.ascontiguousarray(q2_array))
    f32 = a.dtype == torch.float32 and b.dtype == torch.float32
    dtype = torch.float32 if f32 else torch.float64
    a = a.to(device=dev, dtype=dtype).contiguous()
    b = b.to(device=dev, dtype=dtype).contiguous()
    M, A, B = int(a.shape[0]), int(a.shape[1]), int(b.shape[1])
    out = torch.empty((M,), dtype=torch.float64, device=dev)
    _lib.check(lib.drex_misorientation_angles(_lib.ptr(a), _lib.ptr(b), int(f32), M, A, B,
                                              _lib.ptr(out), _lib.stream_ptr(torch)))
    return out if is_torch else out.cpu().numpy()
