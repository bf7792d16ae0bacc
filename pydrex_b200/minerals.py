"""Mineral texture state and its integration: drop-in for the hot path of pydrex.minerals.

`Mineral.update_orientations`, `update_all` and `voigt_averages` keep the reference's
signatures (/root/reference/src/pydrex/minerals.py:312-320, 659-667, 688-693); the state
lives in the same growing Python lists of numpy arrays.  The integration itself is the
on-device adaptive integrator of csrc/drex_integrate.cuh (one launch per call); there is
no CPU fallback.
"""

import ctypes
import io
import logging
from dataclasses import dataclass, field
from zipfile import ZipFile

import numpy as np

from . import _lib
from . import core as _core
from . import exceptions as _err
from . import flow as _flow

_log = logging.getLogger("pydrex")

OLIVINE_PRIMARY_AXIS = {
    _core.MineralFabric.olivine_A: "a",
    _core.MineralFabric.olivine_B: "c",
    _core.MineralFabric.olivine_C: "c",
    _core.MineralFabric.olivine_D: "a",
    _core.MineralFabric.olivine_E: "a",
}

OLIVINE_SLIP_SYSTEMS = (
    ([0, 1, 0], [1, 0, 0]),
    ([0, 0, 1], [1, 0, 0]),
    ([0, 1, 0], [0, 0, 1]),
    ([1, 0, 0], [0, 0, 1]),
)


@dataclass
class StiffnessTensors:
    """Stiffness tensors (Voigt 6x6, GPa) of the original D-Rex code (minerals.py:52-99)."""

    olivine: np.ndarray = field(
        default_factory=lambda: np.array(
            [
                [320.71, 69.84, 71.22, 0.0, 0.0, 0.0],
                [69.84, 197.25, 74.8, 0.0, 0.0, 0.0],
                [71.22, 74.8, 234.32, 0.0, 0.0, 0.0],
                [0.0, 0.0, 0.0, 63.77, 0.0, 0.0],
                [0.0, 0.0, 0.0, 0.0, 77.67, 0.0],
                [0.0, 0.0, 0.0, 0.0, 0.0, 78.36],
            ]
        )
    )
    enstatite: np.ndarray = field(
        default_factory=lambda: np.array(
            [
                [236.9, 79.6, 63.2, 0.0, 0.0, 0.0],
                [79.6, 180.5, 56.8, 0.0, 0.0, 0.0],
                [63.2, 56.8, 230.4, 0.0, 0.0, 0.0],
                [0.0, 0.0, 0.0, 84.3, 0.0, 0.0],
                [0.0, 0.0, 0.0, 0.0, 79.4, 0.0],
                [0.0, 0.0, 0.0, 0.0, 0.0, 80.1],
            ]
        )
    )

    def __iter__(self):
        # indexable by MineralPhase: olivine (0) first, enstatite (1) second
        yield self.olivine
        yield self.enstatite


def _tolerances(kwargs, span, n_grains):
    """Map the LSODA keyword arguments of the reference (minerals.py:523-529) onto the
    device integrator's error weights.

    `atol` may be a scalar or, like LSODA's, an array over y = [F(9) | orientations(9N) |
    fractions(N)]; the device integrator keeps one absolute tolerance per BLOCK of y, so an
    array must be uniform within each block (anything else raises instead of being
    silently replaced by its minimum)."""
    kw = dict(kwargs)
    rtol = float(kw.pop("rtol", 1e-6))
    if "atol" in kw:
        atol = np.asarray(kw.pop("atol"), dtype=np.float64)
        if atol.ndim == 0 or atol.size == 1:
            atol_F = atol_o = atol_f = float(atol.reshape(-1)[0])
        elif atol.size == 9 + 10 * n_grains:
            atol = atol.reshape(-1)
            blocks = (atol[:9], atol[9:9 + 9 * n_grains], atol[9 + 9 * n_grains:])
            if any(b.min() != b.max() for b in blocks):
                raise ValueError(
                    "per-component `atol` must be uniform within the deformation-gradient (9), "
                    "orientation (9 n_grains) and fraction (n_grains) blocks of the state vector")
            atol_F, atol_o, atol_f = (float(b[0]) for b in blocks)
        else:
            raise ValueError(f"`atol` must be a scalar or have 9 + 10 n_grains = "
                             f"{9 + 10 * n_grains} entries, not {atol.size}")
        atol_rel = 0.0
    else:
        atol_F = atol_o = atol_f = 1e-4  # |y0| * 1e-6 + 1e-4
        atol_rel = 1e-6
    first_step = kw.pop("first_step", None)
    max_step = kw.pop("max_step", None)
    kw.pop("min_step", None)
    kw.pop("lband", None)
    kw.pop("uband", None)
    max_steps = int(kw.pop("max_steps", 100000))
    if kw:
        raise TypeError(f"unsupported integrator options: {sorted(kw)}")
    tol = _lib.DrexTolerances()
    tol.rtol, tol.atol_abs, tol.atol_rel = rtol, atol_o, atol_rel
    tol.atol_abs_F, tol.atol_abs_f = atol_F, atol_f
    tol.first_step_frac = 0.0
    if first_step is not None and span != 0:
        tol.first_step_frac = min(1.0, abs(float(first_step) / span))
    tol.max_step_frac = 0.0
    if max_step is not None and np.isfinite(max_step) and span != 0:
        tol.max_step_frac = abs(float(max_step) / span)
    tol.max_steps = max_steps
    tol.method = 0
    return tol, first_step is not None


def _regime_segments(get_regime, get_position, t0, t1, n_probe=16):
    """Cut [t0, t1] where `get_regime(t, x(t))` changes.

    The reference evaluates get_regime inside every right-hand side (minerals.py:409-410);
    the device integrator holds the regime fixed over one launch, so an interval that crosses
    a regime boundary becomes several launches (boundary located by bisection to 1e-12 of the
    interval).  Returns [(ta, tb, regime), ...] in time order."""
    span = t1 - t0

    def regime_at(t):
        return int(get_regime(t, get_position(t)))

    if span == 0:
        return [(t0, t1, regime_at(t0))]
    ts = t0 + span * np.linspace(0.0, 1.0, n_probe + 1)
    ts[-1] = t0 + span * (1.0 - 1e-12)  # like L(t): the right end is sampled just inside
    regs = [regime_at(t) for t in ts]
    cuts = []
    for i in range(n_probe):
        if regs[i] == regs[i + 1]:
            continue
        lo, hi = ts[i], ts[i + 1]
        while abs(hi - lo) > 1e-12 * abs(span):
            mid = 0.5 * (lo + hi)
            if mid == lo or mid == hi:
                break
            if regime_at(mid) == regs[i]:
                lo = mid
            else:
                hi = mid
        cuts.append(hi)
    edges = [t0, *cuts, t1]
    return [(a, b, regime_at(0.5 * (a + b))) for a, b in zip(edges[:-1], edges[1:])]


def _integrate_minerals(minerals, params, deformation_gradients, get_velocity_gradient, pathline,
                        get_regime, kwargs):
    """Advance several phases of ONE particle over one interval (one launch, or one per
    deformation-regime segment of the interval)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    time_start, time_end, get_position = pathline
    if not callable(get_velocity_gradient):
        raise ValueError(
            "unable to evaluate velocity gradient callable."
            + " You must provide a callable with signature f(t, x)"
            + " that returns a 3x3 matrix."
        )
    if not callable(get_position):
        raise ValueError(
            "unable to evaluate position callable."
            + " You must provide a callable with signature f(t)"
            + " that returns a 3-component array."
        )
    active, vol_fracs = [], []
    for m in minerals:
        if m.phase not in params["phase_assemblage"]:
            _log.warning("skipping %s mineral, phase omitted from configuration", m.phase)
            continue
        try:
            vol_fracs.append(
                params["phase_fractions"][list(params["phase_assemblage"]).index(m.phase)])
        except IndexError:
            raise ValueError(f"phase must be a valid `MineralPhase`, not {m.phase}") from None
        active.append(m)
    if not active:
        return [np.asarray(F, dtype=np.float64) for F in deformation_gradients]
    N = active[0].n_grains
    if any(m.n_grains != N for m in active):
        # different ensemble sizes cannot share a launch
        out = []
        for m, F in zip(minerals, deformation_gradients):
            out.extend(_integrate_minerals([m], params, [F], get_velocity_gradient, pathline,
                                           get_regime, kwargs))
        return out

    t_a, t_b = float(time_start), float(time_end)
    if get_regime is not None:
        segments = _regime_segments(get_regime, get_position, t_a, t_b)
    else:
        segments = [(t_a, t_b, None)]
    tol, explicit_first = _tolerances(kwargs, t_b - t_a, N)
    dev = torch.device("cuda", torch.cuda.current_device())
    S = len(active)
    n_chunks = -(-N // _lib.chunk_grains())
    f64 = torch.float64
    o_aos = torch.as_tensor(np.stack([np.asarray(m.orientations[-1], dtype=np.float64)
                                      for m in active]).reshape(S, N, 9)).to(dev)
    f = torch.as_tensor(np.stack([np.asarray(m.fractions[-1], dtype=np.float64)
                                  for m in active])).to(dev)
    F_by_mineral = {id(m): F for m, F in zip(minerals, deformation_gradients)}
    F = torch.as_tensor(np.stack([np.asarray(F_by_mineral[id(m)], dtype=np.float64).reshape(9)
                                  for m in active])).to(dev)
    vf = torch.tensor([float(v) for v in vol_fracs], dtype=f64, device=dev)
    # per-chunk step suggestions carried from the previous call (same interval length assumed;
    # a step that does not fit the new interval is clamped by the kernel)
    h_io = torch.zeros((S, n_chunks), dtype=f64, device=dev)
    if not explicit_first and len(segments) == 1:
        for k, m in enumerate(active):
            h_prev = getattr(m, "_h_next", None)
            if h_prev is not None and np.shape(h_prev) == (n_chunks,):
                h_io[k] = torch.as_tensor(h_prev)
    o_start = torch.empty((S, 9, N), dtype=f64, device=dev)  # interval-start orientations, SoA
    o_cur = o_start
    o_new = torch.empty_like(o_start)
    stats = torch.zeros((4, S), dtype=torch.int32, device=dev)
    evals = torch.zeros(S, dtype=torch.int64, device=dev)
    ws_bytes = lib.drex_integrate_workspace_bytes(S, N, dev.index or 0, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr(torch)
    P = _lib.ptr
    _lib.check(lib.drex_pack_soa_f64(P(o_aos), P(o_start), S, N, st))
    totals = np.zeros((3, S), dtype=np.int64)
    total_evals = np.zeros(S, dtype=np.int64)
    for i_seg, (ta, tb, regime) in enumerate(segments):
        last = i_seg == len(segments) - 1
        if regime is not None:
            for m in active:
                m.regime = regime
        kind, nodes = _flow.tabulate_velocity_gradient(get_velocity_gradient, get_position, ta, tb)
        meta_host = np.array([[0] * S, [int(m.phase) for m in active],
                              [int(m.fabric) for m in active], [int(m.regime) for m in active]],
                             dtype=np.int32)
        meta = torch.as_tensor(meta_host).to(dev)
        t0 = torch.tensor([ta], dtype=f64, device=dev)
        t1 = torch.tensor([tb], dtype=f64, device=dev)
        Ln = torch.as_tensor(np.ascontiguousarray(nodes, dtype=np.float64).reshape(1, -1, 9)).to(dev)
        # grain-boundary sliding acts once, at the end of the whole interval, against the
        # orientations at its start (minerals.py:464-474): inner segments run without it
        p = _lib.DrexParams(float(params["stress_exponent"]), float(params["deformation_exponent"]),
                            float(params["nucleation_efficiency"]), float(params["gbm_mobility"]),
                            float(params["gbs_threshold"]) if last else 0.0)
        if o_new is o_start:  # never write over the interval-start orientations
            o_new = torch.empty_like(o_start)
        o_prev = o_start if (last and o_cur is not o_start) else None
        _lib.check(lib.drex_integrate_f64(
            S, 1, N, P(meta[0]), P(meta[1]), P(meta[2]), P(meta[3]), P(vf),
            P(meta_host[1]), P(meta_host[2]), P(meta_host[3]),
            P(F), P(o_cur), P(f), P(o_prev), P(o_new), P(f), P(t0), P(t1), int(kind),
            int(nodes.shape[0]), P(Ln), ctypes.byref(p), ctypes.byref(tol), P(h_io), None,
            P(stats[0]), P(stats[1]), P(stats[2]), P(stats[3]), P(evals), P(ws), ws_bytes, st))
        stats_h = stats.cpu().numpy()
        for k, m in enumerate(active):
            if stats_h[0, k] != 0:
                raise _err.IterationError(
                    f"integration of {m.phase!r} failed: "
                    f"{_lib.STATUS_MESSAGES.get(int(stats_h[0, k]))}")
        totals += stats_h[1:4]
        total_evals += evals.cpu().numpy()
        o_cur, o_new = o_new, o_cur
        if not last:
            h_io.zero_()
    _lib.check(lib.drex_unpack_soa_f64(P(o_cur), P(o_aos), S, N, st))
    o_h = o_aos.cpu().numpy().reshape(S, N, 3, 3)
    f_h = f.cpu().numpy()
    F_h = F.cpu().numpy().reshape(S, 3, 3)
    h_h = h_io.cpu().numpy()
    results = {}
    for k, m in enumerate(active):
        m.orientations.append(o_h[k].copy())
        m.fractions.append(f_h[k].copy())
        m._h_next = h_h[k].copy() if len(segments) == 1 else None
        # counters per grain: right-hand sides and accepted / rejected steps averaged over
        # the chunks of 32 grains, which step independently
        m.n_rhs = getattr(m, "n_rhs", 0) + int(round(total_evals[k] / N))
        m.n_steps = getattr(m, "n_steps", 0) + int(round(totals[1, k] / n_chunks))
        m.n_rejected = getattr(m, "n_rejected", 0) + int(round(totals[2, k] / n_chunks))
        if get_regime is not None:
            m.regime = get_regime(time_end, get_position(time_end))
        results[id(m)] = F_h[k].copy()
    return [results.get(id(m), np.asarray(F0, dtype=np.float64))
            for m, F0 in zip(minerals, deformation_gradients)]


@dataclass
class Mineral:
    """Polycrystal texture of a single mineral phase (minerals.py:124-656).

    Same constructor arguments and attributes as the reference: `phase`, `fabric`,
    `regime`, `n_grains`, `fractions_init`, `orientations_init`, `fractions`,
    `orientations` (growing lists of numpy arrays), `seed`, `lband`, `uband` (the two
    band arguments are accepted and unused: no Jacobian is ever formed).
    """

    phase: int = _core.MineralPhase.olivine
    fabric: int = _core.MineralFabric.olivine_A
    regime: int = _core.DeformationRegime.matrix_dislocation
    n_grains: int = _core.DefaultParams().number_of_grains
    fractions_init: np.ndarray | None = None
    orientations_init: np.ndarray | None = None
    fractions: list = field(default_factory=list)
    orientations: list = field(default_factory=list)
    seed: int | None = None
    lband: int | None = None
    uband: int | None = None

    def __post_init__(self):
        """Initialise random orientations and uniform volume fractions (minerals.py:256-285)."""
        if self.fractions_init is None:
            self.fractions_init = np.full(self.n_grains, 1.0 / self.n_grains)
        if self.orientations_init is None:
            from scipy.spatial.transform import Rotation

            self.orientations_init = Rotation.random(
                self.n_grains, random_state=self.seed
            ).as_matrix()
        self.fractions.append(self.fractions_init)
        self.orientations.append(self.orientations_init)
        del self.fractions_init
        del self.orientations_init
        _log.info("created %s", self)

    def __str__(self):
        return (
            f"{self.__class__.__qualname__}(phase={self.phase!r}, fabric={self.fabric!r}, "
            f"regime={self.regime!r}, n_grains={self.n_grains}, "
            f"fractions=<list of {len(self.fractions)} x {np.shape(self.fractions[0])}>, "
            f"orientations=<list of {len(self.orientations)} x {np.shape(self.orientations[0])}>)"
        )

    def __eq__(self, other):
        if other.__class__ is not self.__class__:
            return False
        return (
            self.phase == other.phase
            and self.fabric == other.fabric
            and self.regime == other.regime
            and self.n_grains == other.n_grains
            and len(self.fractions) == len(other.fractions)
            and len(self.orientations) == len(other.orientations)
            and all(np.array_equal(a, b) for a, b in zip(self.fractions, other.fractions))
            and all(np.array_equal(a, b) for a, b in zip(self.orientations, other.orientations))
            and self.seed == other.seed
            and self.lband == other.lband
            and self.uband == other.uband
        )

    def update_orientations(self, params, deformation_gradient, get_velocity_gradient, pathline,
                            get_regime=None, **kwargs):
        """Update orientations and volume distribution (minerals.py:312-541).

        - `params` — dict with the keys of `DefaultParams.as_dict()`
        - `deformation_gradient` — 3x3 deformation gradient tensor
        - `get_velocity_gradient` — callable f(t, x) returning a 3x3 matrix
        - `pathline` — (t_start, t_end, get_position) with get_position a callable f(t)
        - `get_regime` — optional callable f(t, x) returning a `DeformationRegime`; the
          interval is probed for regime changes and cut there (one launch per segment, the
          device integrator holds the regime fixed over a launch), and it is evaluated once
          more at the end to update `self.regime`

        Keyword arguments understood: `rtol`, `atol`, `first_step`, `max_step` (the LSODA
        options of the reference, minerals.py:523-529) and `max_steps`.
        Returns the updated deformation gradient; appends the new texture to
        `self.orientations` / `self.fractions`.  Raises `IterationError` when the
        integrator fails (minerals.py:457-459).
        """
        return _integrate_minerals([self], params, [deformation_gradient], get_velocity_gradient,
                                   pathline, get_regime, kwargs)[0]

    def save(self, filename, postfix=None):
        """Save all stored snapshots to NPZ in the reference's format (minerals.py:543-593)."""
        if len(self.fractions) != len(self.orientations):
            raise ValueError("Length of stored results must match.")
        if not (self.fractions[0].shape[0] == self.orientations[0].shape[0] == self.n_grains):
            raise ValueError("Size of CPO data arrays must match number of grains.")
        data = {
            "meta": np.array([self.phase, self.fabric, self.regime], dtype=np.uint8),
            "fractions": np.stack(self.fractions),
            "orientations": np.stack(self.orientations),
        }
        if postfix is not None:
            with ZipFile(filename, mode="a", allowZip64=True) as archive:
                for key, value in data.items():
                    with archive.open(f"{key}_{postfix}", "w", force_zip64=True) as file:
                        buffer = io.BytesIO()
                        np.save(buffer, value)
                        file.write(buffer.getvalue())
        else:
            np.savez(filename, **data)

    def load(self, filename, postfix=None):
        """Load CPO data from an NPZ file written by `save` (minerals.py:595-621)."""
        if not filename.endswith(".npz"):
            raise ValueError(f"Must only load from numpy NPZ format. Cannot load from {filename}.")
        data = np.load(filename)
        sfx = "" if postfix is None else f"_{postfix}"
        self.phase, self.fabric, self.regime = (int(x) for x in data["meta" + sfx])
        self.fractions = list(data["fractions" + sfx])
        self.orientations = list(data["orientations" + sfx])
        self.n_grains = len(self.fractions[0])

    @classmethod
    def from_file(cls, filename, postfix=None):
        """Construct a `Mineral` from an NPZ file written by `save` (minerals.py:623-656)."""
        if not filename.endswith(".npz"):
            raise ValueError(f"Must only load from numpy NPZ format. Cannot load from {filename}.")
        data = np.load(filename)
        sfx = "" if postfix is None else f"_{postfix}"
        phase, fabric, regime = data["meta" + sfx]
        fractions = list(data["fractions" + sfx])
        orientations = list(data["orientations" + sfx])
        mineral = cls(int(phase), int(fabric), int(regime), n_grains=len(fractions[0]),
                      fractions_init=fractions[0], orientations_init=orientations[0])
        mineral.fractions = fractions
        mineral.orientations = orientations
        return mineral


def update_all(minerals, params, deformation_gradient, get_velocity_gradient, pathline,
               get_regime=None, **kwargs):
    """Update all mineral phases of one particle (minerals.py:659-684).

    The phases do not couple (they share only L(t) and F), so they go to the device
    together: one launch, one CTA per phase.  Returns the last phase's deformation gradient
    like the reference.
    """
    Fs = _integrate_minerals(list(minerals), params, [deformation_gradient] * len(minerals),
                             get_velocity_gradient, pathline, get_regime, kwargs)
    return Fs[-1]


def voigt_averages(minerals, phase_assemblage, phase_fractions,
                   elastic_tensors=StiffnessTensors()):
    """Voigt-averaged elastic tensors of a collection of minerals (minerals.py:688-740).

    Returns `[n_steps, 6, 6]` (GPa).  Raises ValueError for unequal grain counts or
    unequal numbers of stored snapshots.
    """
    n_grains = minerals[0].n_grains
    if not np.all([m.n_grains == n_grains for m in minerals[1:]]):
        raise ValueError("cannot average minerals with unequal grain counts")
    n_steps = len(minerals[0].orientations)
    if not np.all([len(m.orientations) == n_steps for m in minerals[1:]]):
        raise ValueError("cannot average minerals with variable-length orientation arrays")
    if not np.all([len(m.fractions) == n_steps for m in minerals]):
        raise ValueError("cannot average minerals with variable-length grain volume arrays")
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    tensors = [np.asarray(S, dtype=np.float64) for S in elastic_tensors]
    phase_assemblage = list(phase_assemblage)
    n_min = len(minerals)
    S_total = n_min * n_steps
    o = np.empty((n_min, n_steps, n_grains, 9))
    f = np.empty((n_min, n_steps, n_grains))
    stiff = np.empty((n_min, n_steps, 36))
    pfrac = np.empty((n_min, n_steps))
    for k, m in enumerate(minerals):
        idx = phase_assemblage.index(m.phase)
        o[k] = np.asarray(m.orientations, dtype=np.float64).reshape(n_steps, n_grains, 9)
        f[k] = np.asarray(m.fractions, dtype=np.float64)
        stiff[k] = tensors[idx].reshape(36)
        pfrac[k] = phase_fractions[idx]
    o_d = torch.as_tensor(o.reshape(S_total, n_grains, 9)).to(dev)
    f_d = torch.as_tensor(f.reshape(S_total, n_grains)).to(dev)
    stiff_d = torch.as_tensor(stiff.reshape(S_total, 36)).to(dev)
    pfrac_d = torch.as_tensor(pfrac.reshape(S_total)).to(dev)
    o_soa = torch.empty((S_total, 9, n_grains), dtype=torch.float64, device=dev)
    Cbar = torch.empty((S_total, 36), dtype=torch.float64, device=dev)
    st = _lib.stream_ptr(torch)
    P = _lib.ptr
    _lib.check(lib.drex_pack_soa_f64(P(o_d), P(o_soa), S_total, n_grains, st))
    _lib.check(lib.drex_voigt_average_f64(P(o_soa), P(f_d), P(stiff_d), P(pfrac_d), S_total,
                                          n_grains, P(Cbar), 0, st))
    return Cbar.reshape(n_min, n_steps, 6, 6).sum(dim=0).cpu().numpy()
