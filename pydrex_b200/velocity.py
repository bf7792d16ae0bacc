"""Steady-state analytic flows (drop-in for pydrex.velocity, reference velocity.py:17-332).

Each factory returns `(get_velocity, get_velocity_gradient)` callables with the reference's
signature `f(t, x)` (x a 3-vector, result a 3-vector / 3x3 matrix), so they plug into
`Mineral.update_orientations` and `pathlines.get_pathline` unchanged.  The callables also
carry a `.field` descriptor (`_lib.DrexFlowField`): that is what the device consumes --
`pathlines.get_pathlines` traces whole particle clouds through the flow on the GPU and
`evaluate(field, points)` evaluates velocity and gradient for many points in one launch.
Calling a callable with a single point goes through the same kernel (one tiny launch).
"""

import numpy as np

from . import _lib


def _to_indices2d(horizontal, vertical):
    """geometry.to_indices2d (reference geometry.py:317-334)."""
    table = {("X", "Y"): (0, 1), ("X", "Z"): (0, 2), ("Y", "X"): (1, 0),
             ("Y", "Z"): (1, 2), ("Z", "X"): (2, 0), ("Z", "Y"): (2, 1)}
    key = (str(horizontal).upper(), str(vertical).upper())
    if key not in table:
        raise ValueError
    return table[key]


class _FlowCallable:
    """`f(t, x)` for one analytic flow; `.field` describes it to the device.

    x may be one point [3] (reference signature) or many [P, 3]; either way the closed form
    is evaluated by the library on the GPU (there is no host copy of the formulas).
    """

    def __init__(self, field, gradient):
        self.field = field
        self.gradient = gradient

    def __call__(self, t, x):
        fl = self.field
        pts = np.asarray(x, dtype=np.float64)
        single = pts.ndim == 1
        v, L = evaluate(fl, pts.reshape(-1, 3))
        if fl.kind == _lib.FLOW_CELL and np.isnan(v).any():  # velocity.py:35-41
            bad = pts.reshape(-1, 3)[np.isnan(v).any(axis=1)][0]
            lim = fl.b / 2
            raise ValueError(
                f"position ({bad[fl.i0]}, {bad[fl.i1]}) is outside domain with xᵢ ∈ [-{lim}, {lim}]")
        out = L if self.gradient else v
        return out[0] if single else out


def _pair(kind, i0, i1, a, b=0.0):
    field = _lib.DrexFlowField(int(kind), int(i0), int(i1), float(a), float(b))
    return _FlowCallable(field, False), _FlowCallable(field, True)


def simple_shear_2d(direction, deformation_plane, strain_rate):
    """Simple shear velocity and velocity gradient callables (velocity.py:100-157).

    - `direction` (one of {"X", "Y", "Z"}) — velocity vector direction
    - `deformation_plane` (one of {"X", "Y", "Z"}) — direction of velocity gradient
    - `strain_rate` — 1/2 × magnitude of the largest eigenvalue of the velocity gradient
    """
    try:
        i0, i1 = _to_indices2d(direction, deformation_plane)
    except ValueError:
        raise ValueError(
            "unsupported shear type with"
            + f" direction = {direction}, deformation_plane = {deformation_plane}"
        )
    return _pair(_lib.FLOW_SIMPLE_SHEAR, i0, i1, strain_rate)


def cell_2d(horizontal, vertical, velocity_edge, edge_length=2.0):
    """Steady-state 2D Stokes convection cell centred at the origin (velocity.py:160-259):
    u = U cos(πx/d) sin(πz/d) ĥ − U sin(πx/d) cos(πz/d) v̂."""
    if edge_length < 0:
        raise ValueError(f"edge length of 2D cell must be positive, not {edge_length}")
    try:
        i0, i1 = _to_indices2d(horizontal, vertical)
    except ValueError:
        raise ValueError(
            "unsupported convection cell geometry with"
            + f" horizontal = {horizontal}, vertical = {vertical}"
        )
    return _pair(_lib.FLOW_CELL, i0, i1, velocity_edge, edge_length)


def corner_2d(horizontal, vertical, plate_speed):
    """Steady-state 2D corner flow (velocity.py:262-332): the analytic solution for an
    isoviscous half-space under a spreading plate moving at `plate_speed`."""
    try:
        i0, i1 = _to_indices2d(horizontal, vertical)
    except ValueError:
        raise ValueError(
            "unsupported corner flow geometry with"
            + f" horizontal = {horizontal}, vertical = {vertical}"
        )
    return _pair(_lib.FLOW_CORNER, i0, i1, plate_speed)


def field_of(get_velocity, get_velocity_gradient=None):
    """The device descriptor behind a pair of callables made by this module."""
    field = getattr(get_velocity, "field", None)
    other = getattr(get_velocity_gradient, "field", field)
    if field is None or other is not field:
        raise ValueError(
            "device pathline tracing needs the analytic flows of pydrex_b200.velocity "
            "(both callables from the same simple_shear_2d / cell_2d / corner_2d call); "
            "arbitrary Python callables cannot run on the GPU and there is no CPU fallback")
    return field


def evaluate(field, points):
    """Velocity [P, 3] and velocity gradient [P, 3, 3] of an analytic flow at `points`
    [P, 3] in one launch.  numpy in, numpy out; torch CUDA tensor in, torch out."""
    if not isinstance(field, _lib.DrexFlowField):
        field = field_of(field)
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    is_torch = type(points).__module__.startswith("torch")
    x = points if is_torch else torch.as_tensor(np.ascontiguousarray(points, dtype=np.float64))
    x = x.to(device=dev, dtype=torch.float64).reshape(-1, 3).contiguous()
    P = int(x.shape[0])
    v = torch.empty((P, 3), dtype=torch.float64, device=dev)
    L = torch.empty((P, 3, 3), dtype=torch.float64, device=dev)
    import ctypes

    _lib.check(lib.drex_flow_field_f64(ctypes.byref(field), _lib.ptr(x), P, _lib.ptr(v),
                                       _lib.ptr(L), _lib.stream_ptr(torch)))
    if is_torch:
        return v, L
    return v.cpu().numpy(), L.cpu().numpy()
