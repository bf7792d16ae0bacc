"""Pathline construction on the device (drop-in for pydrex.pathlines, reference
pathlines.py:10-138).

`get_pathline` keeps the reference's single-particle signature and return value
`(timestamps, get_position)`.  `get_pathlines` is the batched form the GPU wants: all
particles of a cloud traced in one launch, returning a `Pathlines` record that already holds,
for every outer interval, the Chebyshev table of L(t) that `ParticleEnsemble.update` /
`drex_integrate_f64` consume -- positions never travel to the host.

The flow must be one of the analytic fields of `pydrex_b200.velocity`: Python callables cannot
run on the GPU, and there is no CPU tracer in this package.
"""

import ctypes

import numpy as np

from . import _lib
from . import flow as _flow
from . import velocity as _velocity

#: the reference integrates "forever" = 100 Myr in seconds (pathlines.py:91)
TIME_LIMIT = -100e6 * 365.25 * 8.64e4


class Pathlines:
    """Result of `get_pathlines` (device tensors; S = regular_steps, K = nodes per interval).

    - `timestamps` [S+1, P] ascending in the first axis, last row 0 (t < 0 along the path)
    - `positions` [S+1, P, 3], `strains` [S+1, P] (tensorial strain accumulated since row 0)
    - `node_positions` [S, P, K, 3], `L_nodes` [S, P, K, 3, 3] at the Chebyshev-Lobatto nodes
      of each interval (`flow.chebyshev_lobatto(K)`), ready for `ParticleEnsemble.update(
      timestamps[j], timestamps[j + 1], L_nodes[j], kind=_lib.L_CHEBYSHEV)`
    - `t_start` [P], `status` [P] (0 ok, 4 = time limit reached, see `_lib.PATH_MESSAGES`)
    """

    def __init__(self, field, timestamps, positions, strains, node_positions, L_nodes, t_start,
                 status):
        self.field = field
        self.timestamps, self.positions, self.strains = timestamps, positions, strains
        self.node_positions, self.L_nodes = node_positions, L_nodes
        self.t_start, self.status = t_start, status

    @property
    def n_particles(self):
        return int(self.timestamps.shape[1])

    @property
    def n_intervals(self):
        return int(self.timestamps.shape[0]) - 1

    def interval(self, j):
        """`(t0 [P], t1 [P], L_nodes [P, K, 3, 3])` of outer interval j."""
        return self.timestamps[j], self.timestamps[j + 1], self.L_nodes[j]

    def position_callable(self, p):
        """`get_position(t)` of particle p, like the interpolant the reference returns: the
        Chebyshev interpolant through the traced node positions of the interval containing
        t (the first / last interval extrapolates)."""
        ts = self.timestamps[:, p].cpu().numpy()
        nodes = self.node_positions[:, p].cpu().numpy()
        return _PositionInterpolant(ts, nodes)


class _PositionInterpolant:
    def __init__(self, timestamps, node_positions):
        self.timestamps = timestamps
        self.node_positions = node_positions

    def __call__(self, t):
        ts = self.timestamps
        j = int(np.clip(np.searchsorted(ts, t, side="right") - 1, 0, len(ts) - 2))
        width = ts[j + 1] - ts[j]
        if width == 0.0:
            return self.node_positions[j, 0].copy()
        return _flow.barycentric_interpolate(self.node_positions[j], (t - ts[j]) / width)


def _status_check(status):
    bad = np.nonzero((status != 0) & (status != 4))[0]
    if bad.size:
        code = int(status[bad[0]])
        raise ValueError(f"pathline of particle {int(bad[0])}: {_lib.PATH_MESSAGES[code]}"
                         + (f" ({bad.size} particles failed)" if bad.size > 1 else ""))


def get_pathlines(final_locations, get_velocity, get_velocity_gradient, min_coords, max_coords,
                  max_strain, regular_steps, n_nodes=9, rtol=1e-11, atol=None, timestamps=None,
                  time_limit=TIME_LIMIT, max_steps=1000000, check=True):
    """Trace the pathlines of P particles backwards from `final_locations` [P, 3] in one
    launch (batched pathlines.get_pathline, pathlines.py:10-116).

    - `get_velocity`, `get_velocity_gradient` — callables from `pydrex_b200.velocity` (or the
      `DrexFlowField` descriptor itself as `get_velocity`)
    - `min_coords`, `max_coords`, `max_strain` — as in the reference
    - `regular_steps` — number of outer intervals S; time stamps are
      `linspace(t_start, 0, S + 1)` per particle
    - `n_nodes` — Chebyshev-Lobatto nodes per interval for the L(t) tables (>= 2)
    - `rtol`, `atol` — local error control of the tracer on position and strain alike (the
      reference's LSODA call uses 1e-5 / 1e-8; the device default is tighter because the
      work is negligible: atol = 1e-14 x the largest box edge)
    - `timestamps` — optional [S+1, P] ascending stamps (<= 0) to use instead of linspace
    Returns a `Pathlines` record of device tensors.
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    field = get_velocity if isinstance(get_velocity, _lib.DrexFlowField) else \
        _velocity.field_of(get_velocity, get_velocity_gradient)
    dev = torch.device("cuda", torch.cuda.current_device())
    f64 = torch.float64
    x = final_locations if torch.is_tensor(final_locations) else torch.as_tensor(
        np.ascontiguousarray(final_locations, dtype=np.float64))
    x = x.to(device=dev, dtype=f64).reshape(-1, 3).contiguous()
    P, S, K = int(x.shape[0]), int(regular_steps), int(n_nodes)
    if S < 1:
        raise ValueError("regular_steps must be a positive number of intervals")
    if K < 2:
        raise ValueError("n_nodes must be at least 2")
    lo = (ctypes.c_double * 3)(*[float(v) for v in np.asarray(min_coords).ravel()])
    hi = (ctypes.c_double * 3)(*[float(v) for v in np.asarray(max_coords).ravel()])
    if atol is None:
        atol = 1e-14 * max(hi[i] - lo[i] for i in range(3))
    given = timestamps is not None
    if given:
        ts = timestamps if torch.is_tensor(timestamps) else torch.as_tensor(np.asarray(timestamps))
        ts = ts.to(device=dev, dtype=f64).reshape(S + 1, P).contiguous().clone()
    else:
        ts = torch.empty((S + 1, P), dtype=f64, device=dev)
    pos = torch.empty((S + 1, P, 3), dtype=f64, device=dev)
    strains = torch.empty((S + 1, P), dtype=f64, device=dev)
    node_pos = torch.empty((S, P, K, 3), dtype=f64, device=dev)
    L_nodes = torch.empty((S, P, K, 3, 3), dtype=f64, device=dev)
    t_start = torch.empty((P,), dtype=f64, device=dev)
    status = torch.empty((P,), dtype=torch.int32, device=dev)
    Pp = _lib.ptr
    _lib.check(lib.drex_pathlines_f64(
        ctypes.byref(field), Pp(x), P, ctypes.byref(lo), ctypes.byref(hi), float(max_strain),
        float(time_limit), float(rtol), float(atol), S, K, int(given), int(max_steps),
        Pp(t_start), Pp(ts), Pp(pos), Pp(strains), Pp(node_pos), Pp(L_nodes), None, 0, None,
        Pp(status), _lib.stream_ptr(torch)))
    if check:
        _status_check(status.cpu().numpy())
    return Pathlines(field, ts, pos, strains, node_pos, L_nodes, t_start, status)


def solver_timestamps(final_location, get_velocity, get_velocity_gradient, min_coords, max_coords,
                      max_strain, rtol=1e-5, atol=1e-8, time_limit=TIME_LIMIT, max_record=100000):
    """Ascending accepted step times of the adaptive tracer for one particle, from t_start
    to 0: what `get_pathline(..., regular_steps=None)` hands out."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    field = _velocity.field_of(get_velocity, get_velocity_gradient)
    dev = torch.device("cuda", torch.cuda.current_device())
    x = torch.as_tensor(np.ascontiguousarray(final_location, dtype=np.float64)).to(dev).reshape(1, 3)
    lo = (ctypes.c_double * 3)(*[float(v) for v in np.asarray(min_coords).ravel()])
    hi = (ctypes.c_double * 3)(*[float(v) for v in np.asarray(max_coords).ravel()])
    times = torch.zeros((1, max_record), dtype=torch.float64, device=dev)
    n_taken = torch.zeros((1,), dtype=torch.int32, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    Pp = _lib.ptr
    _lib.check(lib.drex_pathlines_f64(
        ctypes.byref(field), Pp(x), 1, ctypes.byref(lo), ctypes.byref(hi), float(max_strain),
        float(time_limit), float(rtol), float(atol), 0, 0, 0, 0, None, None, None, None, None,
        None, Pp(times), int(max_record), Pp(n_taken), Pp(status), _lib.stream_ptr(torch)))
    _status_check(status.cpu().numpy())
    n = min(int(n_taken[0]), max_record)
    return np.concatenate([times[0, :n].cpu().numpy()[::-1], [0.0]])


def get_pathline(final_location, get_velocity, get_velocity_gradient, min_coords, max_coords,
                 max_strain, regular_steps=None, **kwargs):
    """Determine the pathline for a particle in a steady state flow (pathlines.py:10-116).

    The pathline terminates at `final_location` and is calculated backwards in time (t < 0)
    until `max_strain` has accumulated or a domain boundary is reached.  Returns
    `(timestamps, get_position)`: ascending time points ending at 0 -- `regular_steps + 1`
    regular ones, or the tracer's own accepted steps when `regular_steps` is None -- and a
    callable evaluating the position at a time inside that range.

    `rtol` / `atol` keyword arguments are honoured (defaults as in the reference: 1e-5 /
    1e-8 for the solver steps handed out when `regular_steps` is None); the keyword
    arguments the reference forwards to scipy (`method`, `first_step`, ...) are ignored with
    the exception of the illegal ones, which warn like the reference does.
    """
    import warnings

    for key in ("events", "jac", "dense_output", "args"):
        if kwargs.pop(key, None) is not None:
            warnings.warn(f"ignoring illegal keyword argument: {key}")
    final_location = np.asarray(final_location, dtype=np.float64)
    stamps = None
    if regular_steps is None:
        stamps = solver_timestamps(final_location, get_velocity, get_velocity_gradient,
                                   min_coords, max_coords, max_strain,
                                   rtol=kwargs.get("rtol", 1e-5), atol=kwargs.get("atol", 1e-8))
        regular_steps = len(stamps) - 1
    extra = {k: kwargs[k] for k in ("rtol", "atol") if k in kwargs and stamps is None}
    paths = get_pathlines(final_location[None], get_velocity, get_velocity_gradient, min_coords,
                          max_coords, max_strain, regular_steps,
                          timestamps=None if stamps is None else stamps[:, None], **extra)
    return paths.timestamps[:, 0].cpu().numpy(), paths.position_callable(0)
