"""Texture diagnostics on device (drop-in for the M-index part of pydrex.diagnostics)."""

import numpy as np

from . import _lib
from . import geometry as _geo
from . import stats as _stats


def _mindex_device(orientation_stack, system):
    """M-index and pair-angle histogram counts for a stack [T, N, 3, 3] in one launch set."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    if not isinstance(system, _geo.LatticeSystem):
        raise ValueError(f"unsupported lattice system: {system}")
    tmax = _stats._max_misorientation(system)
    a, b = system.value
    dev = torch.device("cuda", torch.cuda.current_device())
    is_torch = type(orientation_stack).__module__.startswith("torch")
    o = orientation_stack if is_torch else torch.as_tensor(
        np.ascontiguousarray(orientation_stack, dtype=np.float64))
    o = o.to(device=dev, dtype=torch.float64).contiguous()
    T, N = int(o.shape[0]), int(o.shape[1])
    o_soa = torch.empty((T, 9, N), dtype=torch.float64, device=dev)
    st = _lib.stream_ptr(torch)
    P = _lib.ptr
    _lib.check(lib.drex_pack_soa_f64(P(o.reshape(T, N, 9)), P(o_soa), T, N, st))
    ws_bytes = lib.drex_mindex_workspace_bytes(T, N, a, b)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    M = torch.empty(T, dtype=torch.float64, device=dev)
    counts = torch.empty((T, tmax), dtype=torch.int64, device=dev)
    rc = lib.drex_mindex_f32(P(o_soa), T, N, a, b, P(M), P(counts), P(ws), ws_bytes, st)
    if rc == _lib.DREX_E_UNSUPPORTED and "density undefined" in _lib.last_error():
        raise AssertionError(_lib.last_error())  # reference: `assert False` (stats.py:198)
    _lib.check(rc)
    return M.cpu().numpy(), counts.cpu().numpy()


def misorientation_index(orientations, system, bins=None):
    """M-index of a polycrystal texture (diagnostics.py:332-367; Skemer et al. 2005)."""
    M, _ = _mindex_device(np.asarray(orientations)[None], system)
    return float(M[0])


def misorientation_indices(orientation_stack, system, bins=None, ncpus=None, pool=None):
    """M-indices of a series of textures [T, N, 3, 3] (diagnostics.py:273-329).

    `ncpus` and `pool` are accepted for signature compatibility; the snapshots are
    batched on the GPU instead of being farmed out to processes.
    """
    M, _ = _mindex_device(orientation_stack, system)
    return M


def elasticity_components(voigt_matrices):
    """Elasticity decomposition of a series of Voigt matrices (diagnostics.py:38-182,
    Browaeys & Chevrot 2004): bulk and shear modulus, % anisotropy, the percentages of the
    norm carried by each symmetry class and the hexagonal (transverse-isotropy) axis.

    Same dictionary keys as the reference.  The hexagonal axis is defined up to sign (in
    the reference it inherits the arbitrary sign of LAPACK's eigenvectors).
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    is_torch = type(voigt_matrices).__module__.startswith("torch")
    C = voigt_matrices if is_torch else torch.as_tensor(
        np.ascontiguousarray(voigt_matrices, dtype=np.float64))
    C = C.to(device=dev, dtype=torch.float64).reshape(-1, 36).contiguous()
    P = int(C.shape[0])
    out = torch.empty((P, 11), dtype=torch.float64, device=dev)
    _lib.check(lib.drex_elasticity_components_f64(_lib.ptr(C), P, _lib.ptr(out),
                                                  _lib.stream_ptr(torch)))
    o = out.cpu().numpy()
    keys = ("bulk_modulus", "shear_modulus", "percent_anisotropy", "percent_hexagonal",
            "percent_tetragonal", "percent_orthorhombic", "percent_monoclinic",
            "percent_triclinic")
    res = {k: o[:, i].copy() for i, k in enumerate(keys)}
    res["hexagonal_axis"] = o[:, 8:11].copy()
    return res


_AXIS_ROW = {"a": 0, "b": 1, "c": 2}


def _axis_scatter(orientation_stack, axis):
    """[T, 9] rows of (P, G, R, mean axis, eigenvalues descending) for textures [T, N, 3, 3]."""
    if axis not in _AXIS_ROW:
        raise ValueError(f"axis must be 'a', 'b', or 'c', not {axis}")
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    is_torch = type(orientation_stack).__module__.startswith("torch")
    o = orientation_stack if is_torch else torch.as_tensor(
        np.ascontiguousarray(orientation_stack, dtype=np.float64))
    o = o.to(device=dev, dtype=torch.float64).contiguous()
    T, N = int(o.shape[0]), int(o.shape[1])
    o_soa = torch.empty((T, 9, N), dtype=torch.float64, device=dev)
    out = torch.empty((T, 9), dtype=torch.float64, device=dev)
    st = _lib.stream_ptr(torch)
    _lib.check(lib.drex_pack_soa_f64(_lib.ptr(o.reshape(T, N, 9)), _lib.ptr(o_soa), T, N, st))
    _lib.check(lib.drex_axis_scatter_f64(_lib.ptr(o_soa), T, N, _AXIS_ROW[axis], _lib.ptr(out), st))
    return out.cpu().numpy()


def symmetry_pgr(orientations, axis="a"):
    """Point, Girdle and Random eigenvalue diagnostics of one crystallographic axis
    (diagnostics.py:233-270; Vollmer 1990).  Returns the tuple (P, G, R)."""
    r = _axis_scatter(np.asarray(orientations)[None], axis)[0]
    return float(r[0]), float(r[1]), float(r[2])


def bingham_average(orientations, axis="a"):
    """Antipodally symmetric mean direction of one crystallographic axis
    (diagnostics.py:185-213); unit vector, defined up to sign."""
    return _axis_scatter(np.asarray(orientations)[None], axis)[0][3:6].copy()
