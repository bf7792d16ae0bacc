"""Texture statistics (drop-in for pydrex.stats: resampling and the M-index parts)."""

import ctypes

import numpy as np

from . import _lib
from . import geometry as _geo


def _plain_seed(seed):
    return isinstance(seed, (int, np.integer)) and not isinstance(seed, bool) and 0 <= int(seed) < 2**64


def numpy_uniforms(seeds, n, offsets=None):
    """`numpy.random.default_rng(seed).random(n)` for every seed, as a [T, n] float64 tensor on
    the current CUDA device -- generated there, bit for bit numpy's stream
    (`drex_numpy_uniform_f64`: SeedSequence + PCG64).  `offsets[t]` variates of stream t are
    skipped first.  Seeds that are not plain integers in [0, 2**64) (None, sequences,
    SeedSequence objects) are drawn by numpy on the host instead."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    seeds = list(seeds)
    T, n = len(seeds), int(n)
    if not all(_plain_seed(s) for s in seeds):
        offsets = [0] * T if offsets is None else offsets
        rows = [np.random.default_rng(seed=s).random(int(k) + n)[int(k):] for s, k in zip(seeds, offsets)]
        return torch.as_tensor(np.stack(rows) if T else np.empty((0, n))).to(dev)
    out = torch.empty((T, n), dtype=torch.float64, device=dev)
    if T == 0 or n == 0:
        return out
    d_seeds = torch.as_tensor(np.array([int(s) for s in seeds], dtype=np.uint64).view(np.int64)).to(dev)
    d_off = None
    if offsets is not None:
        d_off = torch.as_tensor(np.array([int(k) for k in offsets], dtype=np.uint64).view(np.int64)).to(dev)
    _lib.check(_lib.load().drex_numpy_uniform_f64(_lib.ptr(d_seeds), _lib.ptr(d_off), T, n, _lib.ptr(out),
                                                  _lib.stream_ptr(torch)))
    return out


def resample_orientations(orientations, fractions, n_samples=None, seed=None):
    """Return new samples from `orientations` weighted by the volume distribution
    (stats.py:14-64).

    - `orientations` — NxMx3x3 array of orientations
    - `fractions` — NxM array of grain volume fractions
    - `n_samples` — optional number of samples to return, default is M
    - `seed` — optional seed for the random number generator

    Returns the Nx`n_samples`x3x3 orientations and associated grain volumes.  numpy in,
    numpy out; torch CUDA tensors in, torch CUDA tensors out.  The uniform variates are
    numpy's `default_rng(seed).random` stream, drawn texture by texture like the reference
    (generated on the device for an integer seed, `numpy_uniforms`); sorting, running sum,
    search and gather run on the device.  Grains with EQUAL fractions
    are ordered by index (numpy.argsort(kind="stable")); the reference's default argsort
    leaves that order to the numpy build, see csrc/drex_resample.cuh.
    """
    is_torch = type(orientations).__module__.startswith("torch")
    o_shape = tuple(orientations.shape) if is_torch else np.asarray(orientations).shape
    f_shape = tuple(fractions.shape) if type(fractions).__module__.startswith("torch") \
        else np.asarray(fractions).shape
    if (
        len(o_shape) != 4
        or len(f_shape) != 2
        or o_shape[0] != f_shape[0]
        or o_shape[1] != f_shape[1]
        or o_shape[2] != o_shape[3] != 3
    ):
        raise ValueError(
            "invalid shape of input arrays,"
            + f" got orientations of shape {o_shape}"
            + f" and fractions of shape {f_shape}"
        )
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    T, N = int(f_shape[0]), int(f_shape[1])
    if n_samples is None:
        n_samples = N
    n_samples = int(n_samples)
    if _plain_seed(seed):  # one seeded stream, cut into per-texture pieces on the device
        uniforms = numpy_uniforms([seed] * T, n_samples, offsets=[i * n_samples for i in range(T)])
    else:
        rng = np.random.default_rng(seed=seed)
        uniforms = torch.as_tensor(rng.random((T, n_samples))).to(dev)

    def to_dev(x):
        if not type(x).__module__.startswith("torch"):
            x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64))
        return x.to(device=dev, dtype=torch.float64).contiguous()

    o, f = to_dev(orientations), to_dev(fractions)
    out_o = torch.empty((T, n_samples, 3, 3), dtype=torch.float64, device=dev)
    out_f = torch.empty((T, n_samples), dtype=torch.float64, device=dev)
    nbytes = lib.drex_resample_workspace_bytes(T, N)
    ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=dev)
    _lib.check(lib.drex_resample_orientations_f64(
        _lib.ptr(o), _lib.ptr(f), _lib.ptr(uniforms), T, N, n_samples, _lib.ptr(out_o),
        _lib.ptr(out_f), _lib.ptr(ws), nbytes, _lib.stream_ptr(torch)))
    if is_torch:
        return out_o, out_f
    return out_o.cpu().numpy(), out_f.cpu().numpy()


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
