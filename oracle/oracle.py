"""Python face of the CPU parity oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, `__graft_entry__.smoke()` and bench.py's `cpu_baseline` / `--impl reference`
legs may import this module; nothing under `pydrex_b200/` does.

It wraps `libdrex_oracle.so` (drex_oracle.c, the C restatement of the reference's
arithmetic) with ctypes and restates the one piece of the hot path that is host control
flow in the reference: the LSODA driver of `Mineral.update_orientations`
(/root/reference/src/pydrex/minerals.py:312-541).

Third-party arithmetic: the time integrator is `scipy.integrate.LSODA` (ODEPACK lsoda),
which the reference calls at minerals.py:518-533.  scipy is unpinned by the reference
(pyproject.toml:36 `scipy >= 1.2`); the golden vectors were generated with the scipy of
this image (1.18.1) and the oracle calls the same installed scipy, exactly as the
reference would.

Parity status: PINNED against tests/golden/*.npz (see tests/test_oracle.py).
"""

import ctypes
import os
import subprocess

import numpy as np
from scipy.integrate import LSODA

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdrex_oracle.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile the C restatement with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "drex_oracle.c")
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    ):
        subprocess.run(["make", "-C", _HERE, "-B", "libdrex_oracle.so"], check=True,
                       capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_strain_rate_max.restype = ctypes.c_double
        L.oracle_strain_rate_max.argtypes = [_dp]
        L.oracle_left_stretch.argtypes = [_dp, _dp]
        L.oracle_get_crss.argtypes = [ctypes.c_int, ctypes.c_int, _dp]
        L.oracle_slip_invariants.argtypes = [_dp, _dp, _dp]
        L.oracle_rotation_and_strain.argtypes = (
            [ctypes.c_int, ctypes.c_int, _dp, _dp, _dp] + [ctypes.c_double] * 3 + [_dp] * 6
        )
        L.oracle_derivatives.argtypes = (
            [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64]
            + [_dp] * 5 + [ctypes.c_double] * 5 + [_dp, _dp]
        )
        L.oracle_grain_batch.argtypes = (
            [ctypes.c_int, ctypes.c_int, ctypes.c_int64, _dp, _dp] + [ctypes.c_double] * 3 + [_dp, _dp]
        )
        L.oracle_extract_vars.argtypes = [_dp, ctypes.c_int64, _dp, _dp, _dp]
        L.oracle_apply_gbs.argtypes = [_dp, _dp, ctypes.c_double, _dp, ctypes.c_int64]
        L.oracle_eval_rhs.argtypes = (
            [_dp, _dp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
            + [ctypes.c_double] * 5 + [_dp, _dp]
        )
        L.oracle_voigt_to_tensor.argtypes = [_dp, _dp]
        L.oracle_tensor_to_voigt.argtypes = [_dp, _dp]
        L.oracle_rotate_tensor.argtypes = [_dp, _dp, _dp]
        L.oracle_voigt_average_accumulate.argtypes = [
            _dp, _dp, ctypes.c_int64, _dp, ctypes.c_double, _dp]
        L.oracle_matrix_to_quat.argtypes = [_dp, _dp]
        L.oracle_symmetry_ops.argtypes = [ctypes.c_int, ctypes.c_int, _dp]
        L.oracle_max_misorientation.argtypes = [ctypes.c_int, ctypes.c_int]
        L.oracle_misorientations_random.restype = ctypes.c_double
        L.oracle_misorientations_random.argtypes = [
            ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.oracle_misorientation_index.restype = ctypes.c_double
        L.oracle_misorientation_index.argtypes = [
            _dp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _dp, _i64p]
        _lib = L
    return _lib


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


# ----------------------------------------------------------------------------- core.py


def get_crss(phase, fabric):
    out = np.empty(4)
    if lib().oracle_get_crss(int(phase), int(fabric), out.ctypes.data_as(_dp)) != 0:
        raise ValueError(f"unsupported phase/fabric: {phase}/{fabric}")
    return out


def slip_invariants(strain_rate, orientation):
    D, Dp = _c(strain_rate)
    a, ap = _c(orientation)
    out = np.empty(4)
    lib().oracle_slip_invariants(Dp, ap, out.ctypes.data_as(_dp))
    return out


def grain_probe(phase, fabric, orientation, strain_rate, velocity_gradient, p, n, lam):
    """All per-grain intermediates of core._get_rotation_and_strain (core.py:687-760)."""
    a, ap = _c(orientation)
    D, Dp = _c(strain_rate)
    Lm, Lp = _c(velocity_gradient)
    da, E = np.empty(9), np.empty(1)
    inv, rates, G, g0 = np.empty(4), np.empty(4), np.empty(9), np.empty(1)
    rc = lib().oracle_rotation_and_strain(
        int(phase), int(fabric), ap, Dp, Lp, p, n, lam,
        da.ctypes.data_as(_dp), E.ctypes.data_as(_dp), inv.ctypes.data_as(_dp),
        rates.ctypes.data_as(_dp), G.ctypes.data_as(_dp), g0.ctypes.data_as(_dp))
    if rc != 0:
        raise ValueError("unsupported phase/fabric")
    return {"orientation_change": da.reshape(3, 3), "strain_energy": E[0], "invariants": inv,
            "slip_rates": rates, "deformation_rate": G.reshape(3, 3), "slip_rate_softest": g0[0]}


def derivatives(regime, phase, fabric, n_grains, orientations, fractions, strain_rate,
                velocity_gradient, deformation_gradient_spin, stress_exponent,
                deformation_exponent, nucleation_efficiency, gbm_mobility, volume_fraction):
    """core.derivatives (core.py:347-474), same positional signature."""
    o, op = _c(orientations)
    f, fp = _c(fractions)
    D, Dp = _c(strain_rate)
    Lm, Lp = _c(velocity_gradient)
    S, Sp = _c(deformation_gradient_spin)
    od = np.empty((n_grains, 3, 3))
    fd = np.empty(n_grains)
    rc = lib().oracle_derivatives(
        int(regime), int(phase), int(fabric), int(n_grains), op, fp, Dp, Lp, Sp,
        float(stress_exponent), float(deformation_exponent), float(nucleation_efficiency),
        float(gbm_mobility), float(volume_fraction),
        od.ctypes.data_as(_dp), fd.ctypes.data_as(_dp))
    if rc == -1:
        raise ValueError("unsupported phase/fabric")
    if rc == -2:
        raise ValueError("this deformation mechanism is not yet supported.")
    if rc == -3:
        raise ValueError(f"regime must be a valid `DeformationRegime`, not {regime}")
    return od, fd


# ----------------------------------------------------------------------------- minerals.py


def extract_vars(y, n_grains):
    y, yp = _c(y)
    F, o, f = np.empty((3, 3)), np.empty((n_grains, 3, 3)), np.empty(n_grains)
    lib().oracle_extract_vars(yp, n_grains, F.ctypes.data_as(_dp), o.ctypes.data_as(_dp),
                              f.ctypes.data_as(_dp))
    return F, o, f


def apply_gbs(orientations, fractions, gbs_threshold, orientations_prev, n_grains):
    o = np.array(orientations, dtype=np.float64, order="C")
    f = np.array(fractions, dtype=np.float64, order="C")
    op, opp = _c(orientations_prev)
    lib().oracle_apply_gbs(o.ctypes.data_as(_dp), f.ctypes.data_as(_dp), float(gbs_threshold),
                           opp, n_grains)
    return o, f


class RhsCounter:
    def __init__(self):
        self.nfev = 0
        self.nsteps = 0


def update_orientations(orientations, fractions, regime, phase, fabric, params,
                        deformation_gradient, get_velocity_gradient, pathline,
                        counter=None, **kwargs):
    """One call of Mineral.update_orientations (minerals.py:312-541) on explicit state.

    Returns (F_new, orientations_new, fractions_new).  `get_velocity_gradient(t, x)` and
    `pathline = (t0, t1, get_position)` are the reference's callables.
    """
    n_grains = len(fractions)
    t0, t1, get_position = pathline
    if not callable(get_velocity_gradient) or not callable(get_position):
        raise ValueError("unable to evaluate velocity gradient / position callable")
    volume_fraction = params["phase_fractions"][list(params["phase_assemblage"]).index(phase)]
    p = float(params["stress_exponent"])
    n = float(params["deformation_exponent"])
    lam = float(params["nucleation_efficiency"])
    M = float(params["gbm_mobility"])
    work = np.empty(10 * n_grains)
    workp = work.ctypes.data_as(_dp)
    L_ = lib()

    def eval_rhs(t, y):  # minerals.py:388-453
        Lm = np.ascontiguousarray(get_velocity_gradient(t, get_position(t)), dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        dy = np.empty_like(y)
        rc = L_.oracle_eval_rhs(
            Lm.ctypes.data_as(_dp), y.ctypes.data_as(_dp), n_grains, int(regime), int(phase),
            int(fabric), p, n, lam, M, float(volume_fraction), dy.ctypes.data_as(_dp), workp)
        if rc != 0:
            raise ValueError("unsupported regime/phase/fabric")
        if counter is not None:
            counter.nfev += 1
        return dy

    o_prev = np.ascontiguousarray(orientations, dtype=np.float64)
    y0 = np.hstack((np.asarray(deformation_gradient, dtype=np.float64).ravel(),
                    o_prev.ravel(), np.asarray(fractions, dtype=np.float64)))
    solver = LSODA(
        eval_rhs, t0, y0, t1,
        atol=kwargs.pop("atol", np.abs(y0 * 1e-6) + 1e-4),       # minerals.py:523
        rtol=kwargs.pop("rtol", 1e-6),                            # minerals.py:524
        first_step=kwargs.pop("first_step", np.abs(t1 - t0) * 1e-1),  # minerals.py:525
        # banded-Jacobian switch of Mineral.__post_init__ (minerals.py:264-275): without it
        # ODEPACK's n^2 work array overflows for N > 4632; no Jacobian is ever formed
        lband=kwargs.pop("lband", 6000 if n_grains > 4632 else None),
        uband=kwargs.pop("uband", 6000 if n_grains > 4632 else None),
        **kwargs,
    )
    while True:
        message = solver.step()
        if message is not None and solver.status == "failed":
            raise RuntimeError(message)  # reference: pydrex.exceptions.IterationError
        if counter is not None:
            counter.nsteps += 1
        if solver.status != "running":
            break
    # minerals.py:464-474 act on a copy of the solver state on every internal step and are
    # discarded by ODEPACK except after the final one.
    _, o, f = extract_vars(solver.y, n_grains)
    o, f = apply_gbs(o, f, params["gbs_threshold"], o_prev, n_grains)
    y_end = np.hstack((solver.y[:9], o.ravel(), f))
    F, o, f = extract_vars(y_end, n_grains)  # minerals.py:536-538
    return F, o, f


# ----------------------------------------------------------------------------- diagnostics


def voigt_average(orientations, fractions, stiffness, phase_fraction=1.0, out=None):
    """One (snapshot, phase) term of minerals.voigt_averages (minerals.py:729-739)."""
    o, op = _c(orientations)
    f, fp = _c(fractions)
    S, Sp = _c(stiffness)
    if out is None:
        out = np.zeros((6, 6))
    lib().oracle_voigt_average_accumulate(op, fp, len(f), Sp, float(phase_fraction),
                                          out.ctypes.data_as(_dp))
    return out


def rotate_voigt(stiffness, rotation):
    """elastic_tensor_to_voigt(rotate(voigt_to_elastic_tensor(S), R)) (tensors.py:167-273)."""
    S, Sp = _c(stiffness)
    R, Rp = _c(rotation)
    T, Tr, V = np.empty(81), np.empty(81), np.empty((6, 6))
    lib().oracle_voigt_to_tensor(Sp, T.ctypes.data_as(_dp))
    lib().oracle_rotate_tensor(T.ctypes.data_as(_dp), Rp, Tr.ctypes.data_as(_dp))
    lib().oracle_tensor_to_voigt(Tr.ctypes.data_as(_dp), V.ctypes.data_as(_dp))
    return V


def matrix_to_quat(m):
    m, mp = _c(m)
    q = np.empty(4)
    lib().oracle_matrix_to_quat(mp, q.ctypes.data_as(_dp))
    return q


def symmetry_ops(system):
    """4x4 matrices acting on scalar-last quaternions; `system` = LatticeSystem.value."""
    buf = np.zeros((16, 4, 4))
    n = lib().oracle_symmetry_ops(int(system[0]), int(system[1]), buf.ctypes.data_as(_dp))
    if n < 0:
        raise ValueError(f"unsupported lattice system: {system}")
    return buf[:n].copy()


def misorientations_random(low, high, system):
    return lib().oracle_misorientations_random(float(low), float(high), int(system[0]),
                                               int(system[1]))


def misorientation_index(orientations, system, return_hist=False):
    """diagnostics.misorientation_index (diagnostics.py:332-367); system = (a, b) tuple."""
    o, op = _c(orientations)
    tmax = lib().oracle_max_misorientation(int(system[0]), int(system[1]))
    if tmax < 0:
        raise ValueError(f"unsupported lattice system: {system}")
    hist = np.empty(tmax)
    counts = np.empty(tmax, dtype=np.int64)
    M = lib().oracle_misorientation_index(op, len(o), int(system[0]), int(system[1]),
                                          hist.ctypes.data_as(_dp), counts.ctypes.data_as(_i64p))
    if return_hist:
        return M, hist, counts
    return M


def resample_orientations(orientations, fractions, n_samples=None, seed=None):
    """numpy restatement of stats.resample_orientations (reference stats.py:14-64) with the
    one freedom the reference leaves open pinned down: grains of EQUAL fraction are ordered
    by index (argsort kind="stable"; the reference's default argsort is unstable and its tie
    order depends on the numpy build)."""
    orientations = np.asarray(orientations, dtype=np.float64)
    fractions = np.asarray(fractions, dtype=np.float64)
    rng = np.random.default_rng(seed=seed)
    if n_samples is None:
        n_samples = fractions.shape[1]
    out_o = np.empty((len(fractions), n_samples, 3, 3))
    out_f = np.empty((len(fractions), n_samples))
    for i, (frac, orient) in enumerate(zip(fractions, orientations)):
        order = np.argsort(frac, kind="stable")
        frac_ascending = frac[order]
        cumfrac = frac_ascending.cumsum()
        cumfrac[-1] = 1.0
        picks = np.searchsorted(cumfrac, rng.random(n_samples))
        out_o[i] = orient[order][picks]
        out_f[i] = frac_ascending[picks]
    return out_o, out_f


# ---- steady flows and pathlines ------------------------------------------


def flow_eval(desc, x):
    """numpy restatement of velocity._simple_shear_2d/_cell_2d/_corner_2d and their gradients
    (reference velocity.py:17-97).  desc = (kind, i0, i1, a, b) with kind 0 simple shear
    (a = strain_rate), 1 convection cell (a = velocity_edge, b = edge_length), 2 corner flow
    (a = plate_speed).  Returns (v[3], L[3,3]); NaN where the reference returns NaN or raises."""
    kind, a, b = int(desc[0]), int(desc[1]), int(desc[2])
    pa, pb = float(desc[3]), float(desc[4])
    x = np.asarray(x, dtype=np.float64)
    v, L = np.zeros(3), np.zeros((3, 3))
    if kind == 0:
        v[a] = x[b] * pa
        L[a, b] = 2 * pa
    elif kind == 1:
        if abs(x[a]) > pb / 2 or abs(x[b]) > pb / 2:
            return np.full(3, np.nan), np.full((3, 3), np.nan)
        cx, cz = np.cos(np.pi * x[a] / pb), np.cos(np.pi * x[b] / pb)
        sx, sz = np.sin(np.pi * x[a] / pb), np.sin(np.pi * x[b] / pb)
        v[a], v[b] = pa * cx * sz, -pa * sx * cz
        L[a, a], L[a, b] = -np.pi / pb * sz * sx, np.pi / pb * cz * cx
        L[b, b], L[b, a] = -np.pi / pb * cz * cx, np.pi / pb * sz * sx
        L *= pa
    else:
        h, w = x[a], x[b]
        if abs(h) < 1e-15 and abs(w) < 1e-15:
            return np.full(3, np.nan), np.full((3, 3), np.nan)
        pf = 2 * pa / np.pi
        v[a] = pf * (np.arctan2(h, -w) + h * w / (h**2 + w**2))
        v[b] = pf * w**2 / (h**2 + w**2)
        L[a, a], L[a, b] = -(h**2) * w, h**3
        L[b, a], L[b, b] = -h * w**2, h**2 * w
        L *= 4 * pa / (np.pi * (h**2 + w**2) ** 2)
    return v, L


def trace_pathline(desc, final_location, min_coords, max_coords, max_strain, regular_steps,
                   rtol=1e-12, atol=None, t_limit=-100e6 * 365.25 * 8.64e4):
    """Restatement of pathlines.get_pathline (reference pathlines.py:10-116) as a clean ODE
    problem: the state is (x, remaining strain) with d(strain)/dt = max|eig(sym L)|
    (utils.strain_increment, utils.py:144-156), integrated backwards from t = 0 by scipy's
    DOP853 at tight tolerance; terminal events are "strain budget used up" and "left the box"
    (pathlines._is_inside, pathlines.py:133-138).  The reference accumulates the strain with
    a rectangle rule over LSODA's (rtol 1e-5) steps, so it agrees with this only to its own
    accuracy -- see tests/test_oracle.py for the measured gap.
    Returns (timestamps[S+1], positions[S+1,3], strains[S+1])."""
    from scipy.integrate import solve_ivp

    lo, hi = np.asarray(min_coords, float), np.asarray(max_coords, float)
    if atol is None:
        atol = 1e-13 * (hi - lo).max()

    def rhs(t, y):
        v, L = flow_eval(desc, y[:3])
        return np.append(v, np.abs(np.linalg.eigvalsh((L + L.T) / 2)).max())

    def budget(t, y):
        return y[3]

    def inside(t, y):  # >= 0 inside the box; degenerate (zero-width) axes do not count
        d = np.concatenate([(y[:3] - lo), (hi - y[:3])])
        wide = np.concatenate([hi > lo, hi > lo])
        return d[wide].min() if wide.any() else 1.0

    budget.terminal = inside.terminal = True
    budget.direction = inside.direction = -1
    y0 = np.append(np.asarray(final_location, float), float(max_strain))
    sol = solve_ivp(rhs, [0.0, t_limit], y0, method="DOP853", rtol=rtol, atol=atol,
                    events=[budget, inside], dense_output=True)
    t_start = sol.t[-1]
    ts = np.linspace(t_start, 0.0, regular_steps + 1)
    ys = sol.sol(ts).T
    return ts, ys[:, :3], ys[:, 3] - ys[0, 3]
