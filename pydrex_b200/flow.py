"""Host-side tabulation of the velocity gradient L(t) along a pathline.

The reference evaluates two Python callables inside every right-hand side,
`get_velocity_gradient(t, get_position(t))` (minerals.py:391-392).  Python cannot run on
the device, so each outer interval [t0, t1] is tabulated once on the host and the kernel
interpolates (include/drex_b200.h DREX_L_*):

  constant   all samples equal                        -> 1 node
  linear     Fluidity-style linear-in-time L          -> 2 nodes
  Chebyshev  K = 2^m + 1 Chebyshev-Lobatto nodes, barycentric interpolation; K doubles
             (the node sets nest, so no sample is wasted) until the interpolant
             reproduces the new samples to `rtol`;
  piecewise  dense uniform table, linear between nodes, when L(t) is not smooth on the
             interval (the Chebyshev refinement did not converge).
"""

import numpy as np

from . import _lib


def chebyshev_lobatto(K):
    """Nodes on [0, 1], ascending: x_j = (1 - cos(pi j / (K - 1))) / 2."""
    if K == 1:
        return np.zeros(1)
    return 0.5 * (1.0 - np.cos(np.pi * np.arange(K) / (K - 1)))


def barycentric_interpolate(values, x):
    """Evaluate the Chebyshev-Lobatto interpolant of `values[K, ...]` at x in [0, 1].

    The interpolant the kernel evaluates (csrc/drex_integrate.cuh, there through its
    Chebyshev coefficients); used here to decide K.
    """
    values = np.asarray(values, dtype=np.float64)
    K = values.shape[0]
    nodes = chebyshev_lobatto(K)
    d = x - nodes
    hit = np.abs(d) < 1e-15
    if hit.any():
        return values[int(np.argmax(hit))]
    w = np.where(np.arange(K) % 2 == 1, -1.0, 1.0)
    w[0] *= 0.5
    w[-1] *= 0.5
    w = w / d
    return np.tensordot(w, values, axes=(0, 0)) / w.sum()


def tabulate_velocity_gradient(get_velocity_gradient, get_position, t0, t1, rtol=1e-10,
                               max_nodes=129):
    """Sample L(t) = get_velocity_gradient(t, get_position(t)) on [t0, t1].

    Returns `(kind, nodes)` with nodes of shape [K, 3, 3].
    """

    def L_at(x):
        # The right end is sampled just inside the interval: callers commonly switch flows
        # with `L1 if t < t_switch else L2` at an outer time stamp, and the right-hand
        # sides that matter all lie at t < t1 (only the error estimate of the last step
        # touches t1 itself).
        t = t0 + (t1 - t0) * min(x, 1.0 - 1e-12)
        return np.asarray(get_velocity_gradient(t, get_position(t)), dtype=np.float64).reshape(3, 3)

    K = 5
    xs = chebyshev_lobatto(K)
    vals = {float(x): L_at(x) for x in xs}
    samples = np.stack([vals[float(x)] for x in xs])
    scale = np.abs(samples).max()
    if not np.isfinite(scale):
        raise ValueError("velocity gradient is not finite on the requested interval")
    tol = rtol * scale
    # constant / linear are declared from the five nodes AND the four points halfway between
    # them, so that a feature narrower than the node spacing is not dropped silently
    xm = 0.5 * (xs[:-1] + xs[1:])
    if np.abs(samples - samples[0]).max() <= tol:
        if max(np.abs(L_at(x) - samples[0]).max() for x in xm) <= tol:
            return _lib.L_CONSTANT, samples[:1].copy()
    else:
        lin = samples[0][None] + xs[:, None, None] * (samples[-1] - samples[0])[None]
        if np.abs(samples - lin).max() <= tol and max(
                np.abs(L_at(x) - (samples[0] + x * (samples[-1] - samples[0]))).max()
                for x in xm) <= tol:
            return _lib.L_LINEAR, np.stack([samples[0], samples[-1]])
    while True:
        K2 = 2 * (K - 1) + 1
        xs2 = chebyshev_lobatto(K2)
        err = 0.0
        for j, x in enumerate(xs2):
            if j % 2 == 0:  # node of the coarser set
                vals.setdefault(float(x), samples[j // 2])
                continue
            v = L_at(x)
            vals[float(x)] = v
            err = max(err, np.abs(v - barycentric_interpolate(samples, x)).max())
        samples2 = np.stack([vals[float(x)] for x in xs2])
        if err <= tol:
            return _lib.L_CHEBYSHEV, samples2
        if K2 >= max_nodes:
            # not smooth on this interval (e.g. a flow switch inside it): polynomial
            # interpolation would ring, so hand the kernel a dense piecewise-linear table
            # and let the step-size control find the kink
            xs_u = np.linspace(0.0, 1.0, 2 * max_nodes - 1)
            return _lib.L_PIECEWISE, np.stack([L_at(x) for x in xs_u])
        K, samples = K2, samples2
