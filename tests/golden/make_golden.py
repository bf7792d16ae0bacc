"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (`python tests/golden/make_golden.py`): it imports the
reference from /root/reference/src through `_refharness` and writes small `.npz`
fixtures that travel with the repo.  The GPU box never runs this script.

What each fixture pins (reference file:line):
- derivatives.npz   `pydrex.core.derivatives` (src/pydrex/core.py:347-474) for every
                    (phase, fabric), regimes 0/1/4/6/7, default and non-default exponents,
                    plus the per-grain helpers (core.py:477-684).
- integrate.npz     `Mineral.update_orientations` / `update_all`
                    (src/pydrex/minerals.py:312-541, 659-684) at default and tight LSODA
                    tolerances, constant and time-dependent velocity gradients.
- diagnostics.npz   `minerals.voigt_averages` (minerals.py:688-740) and
                    `diagnostics.misorientation_index` (diagnostics.py:332-367) with its
                    histogram (stats.py:83-127) and random-texture curve (stats.py:130-200).
- config1_n5000.npz summary numbers of config 1 at full size.
"""

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import _refharness  # noqa: E402

pydrex = _refharness.import_reference()

import numba  # noqa: E402
import scipy  # noqa: E402
from pydrex import core, diagnostics, geometry, minerals, stats, tensors, utils, velocity  # noqa: E402
from pydrex import pathlines as ref_pathlines  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

VERSIONS = json.dumps(
    {"numpy": np.__version__, "scipy": scipy.__version__, "numba": numba.__version__}
)

DEFAULT_PARAMS = {
    "phase_assemblage": (core.MineralPhase.olivine,),
    "phase_fractions": (1.0,),
    "stress_exponent": 1.5,
    "deformation_exponent": 3.5,
    "gbm_mobility": 125,
    "gbs_threshold": 0.3,
    "nucleation_efficiency": 5.0,
    "number_of_grains": 5000,
}


def random_L(rng):
    """Random traceless velocity gradient, normalised like minerals.py:421-439."""
    L = rng.normal(size=(3, 3))
    L -= np.eye(3) * np.trace(L) / 3
    D = (L + L.T) / 2
    s = np.abs(np.linalg.eigvalsh(D)).max()
    return L / s, D / s


def gen_derivatives():
    rng = np.random.default_rng(20261019)
    out = {"versions": VERSIONS}
    combos = [(0, 0), (0, 1), (0, 2), (0, 3), (0, 4), (1, 5)]
    case = 0
    N = 96
    for phase, fabric in combos:
        for regime in (4, 6):
            for p, n, lam in ((1.5, 3.5, 5.0), (1.2, 3.0, 2.5)):
                o = Rotation.random(N, random_state=1000 + case).as_matrix()
                # make a few exactly axis-aligned grains to exercise ties and zeros
                o[0] = np.eye(3)
                o[1] = np.array([[0.0, 1, 0], [-1, 0, 0], [0, 0, 1]])
                f = rng.random(N) ** 3
                f /= f.sum()
                if case % 3 == 0:
                    L = np.array([[0.0, 0, 0], [2.0, 0, 0], [0, 0, 0]])
                    D = (L + L.T) / 2
                elif case % 3 == 1:
                    L, D = random_L(rng)
                else:
                    # uniaxial compression: identity grain has all-zero invariants
                    L = np.diag([0.5, 0.5, -1.0])
                    D = L.copy()
                spin = rng.normal(size=(3, 3))
                M, vf = 125.0, 0.7
                od, fd = core.derivatives(
                    regime, phase, fabric, N, o, f, D, L, spin, p, n, lam, M, vf
                )
                key = f"c{case:02d}"
                out[key + "_meta"] = np.array([regime, phase, fabric], dtype=np.int64)
                out[key + "_scal"] = np.array([p, n, lam, M, vf])
                out[key + "_o"] = o
                out[key + "_f"] = f
                out[key + "_L"] = L
                out[key + "_D"] = D
                out[key + "_spin"] = spin
                out[key + "_od"] = od
                out[key + "_fd"] = fd
                # per-grain helpers (reference private API used by tests/test_core.py)
                crss = core.get_crss(phase, fabric)
                inv = np.array([core._get_slip_invariants(D, o[g]) for g in range(N)])
                rates = np.zeros((N, 4))
                defr = np.zeros((N, 3, 3))
                soft = np.zeros(N)
                energy = np.zeros(N)
                for g in range(N):
                    if np.all(inv[g] == 0):
                        continue
                    if phase == 0:
                        idx = np.argsort(np.abs(inv[g] / crss))
                        rates[g] = core._get_slip_rates_olivine(inv[g], idx, crss, n)
                    else:
                        idx = np.argsort(1 / crss)
                        if abs(inv[g][-1]) > 1e-15:
                            rates[g, -1] = 1
                    defr[g] = core._get_deformation_rate(phase, o[g], rates[g])
                    soft[g] = core._get_slip_rate_softest(defr[g], L)
                    energy[g] = core._get_strain_energy(crss, rates[g], idx, soft[g], p, n, lam)
                out[key + "_crss"] = crss
                out[key + "_inv"] = inv
                out[key + "_rates"] = rates
                out[key + "_defr"] = defr
                out[key + "_soft"] = soft
                out[key + "_energy"] = energy
                case += 1
    # trivial regimes
    for regime in (0, 1, 7):
        o = Rotation.random(8, random_state=77).as_matrix()
        f = np.full(8, 1 / 8)
        L, D = random_L(rng)
        spin = rng.normal(size=(3, 3))
        od, fd = core.derivatives(regime, 0, 0, 8, o, f, D, L, spin, 1.5, 3.5, 5.0, 125.0, 1.0)
        key = f"c{case:02d}"
        out[key + "_meta"] = np.array([regime, 0, 0], dtype=np.int64)
        out[key + "_scal"] = np.array([1.5, 3.5, 5.0, 125.0, 1.0])
        out[key + "_o"] = o
        out[key + "_f"] = f
        out[key + "_L"] = L
        out[key + "_D"] = D
        out[key + "_spin"] = spin
        out[key + "_od"] = od
        out[key + "_fd"] = fd
        case += 1
    out["n_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "derivatives.npz"), **out)
    print("derivatives.npz:", case, "cases")


def run_reference(mins, params, get_L, timestamps, get_x, F0=None, **kw):
    F = np.eye(3) if F0 is None else F0
    nfev = 0
    for t0, t1 in zip(timestamps[:-1], timestamps[1:]):
        F = minerals.update_all(mins, params, F, get_L, (t0, t1, get_x), **kw)
    return F


def gen_integrate():
    out = {"versions": VERSIONS}
    zeros = lambda t: np.zeros(3)  # noqa: E731

    # --- I1: config 1 at small N: olivine A, simple shear, M*=125, chi=0.3
    for tag, N, kw in (
        ("i1_default", 400, {}),
        ("i1_tight", 400, {"rtol": 1e-10, "atol": 1e-12}),
    ):
        _, get_L = velocity.simple_shear_2d("Y", "X", 1.0)
        params = dict(DEFAULT_PARAMS, number_of_grains=N)
        m = minerals.Mineral(n_grains=N, seed=8816)
        ts = np.linspace(0, 1, 25)[:6]
        F = run_reference([m], params, get_L, ts, zeros, **kw)
        out[tag + "_ts"] = ts
        out[tag + "_L"] = get_L(0.0, np.zeros(3))
        out[tag + "_F"] = F
        out[tag + "_o"] = np.stack(m.orientations)
        out[tag + "_f"] = np.stack(m.fractions)

    # --- I2: olivine A + enstatite AB, direction change, phase fractions 0.7/0.3
    for tag, kw in (("i2_default", {}), ("i2_tight", {"rtol": 1e-10, "atol": 1e-12})):
        N = 300
        _, L1 = velocity.simple_shear_2d("X", "Z", 5e-7)
        _, L2 = velocity.simple_shear_2d("Y", "X", 5e-7)
        t_switch = 2e5

        def get_L(t, x, L1=L1, L2=L2, t_switch=t_switch):
            return L1(t, x) if t < t_switch else L2(t, x)

        params = dict(
            DEFAULT_PARAMS,
            number_of_grains=N,
            phase_assemblage=(core.MineralPhase.olivine, core.MineralPhase.enstatite),
            phase_fractions=(0.7, 0.3),
        )
        ens = minerals.Mineral(1, 5, 4, n_grains=N, seed=245623)
        olA = minerals.Mineral(0, 0, 4, n_grains=N, seed=245623)
        ts = np.linspace(0, 4e5, 9)
        F = run_reference([ens, olA], params, get_L, ts, lambda t: np.full(3, np.nan), **kw)
        out[tag + "_ts"] = ts
        out[tag + "_L1"] = L1(0.0, np.zeros(3))
        out[tag + "_L2"] = L2(0.0, np.zeros(3))
        out[tag + "_tswitch"] = np.array(t_switch)
        out[tag + "_F"] = F
        out[tag + "_ol_o"] = np.stack(olA.orientations)
        out[tag + "_ol_f"] = np.stack(olA.fractions)
        out[tag + "_en_o"] = np.stack(ens.orientations)
        out[tag + "_en_f"] = np.stack(ens.fractions)

    # --- I3: time-dependent L along a corner-flow pathline (config 3 at small size)
    plate_speed = 2.0 / (100 * 365 * 86400)
    get_u, get_L = velocity.corner_2d("X", "Z", plate_speed)
    H, W = 2e5, 1e6
    for tag, kw in (("i3_default", {}), ("i3_tight", {"rtol": 1e-10, "atol": 1e-12})):
        N = 200
        final_location = np.array([W, 0.0, -0.3 * H])
        ts, get_x = ref_pathlines.get_pathline(
            final_location, get_u, get_L,
            np.array([0.0, 0.0, -H]), np.array([W, 0.0, 0.0]),
            max_strain=10.0, regular_steps=50,
        )
        params = dict(
            DEFAULT_PARAMS,
            number_of_grains=N,
            phase_assemblage=(core.MineralPhase.olivine, core.MineralPhase.enstatite),
            phase_fractions=(0.7, 0.3),
        )
        olA = minerals.Mineral(0, 0, 4, n_grains=N, seed=8816)
        ens = minerals.Mineral(1, 5, 4, n_grains=N, seed=8816)
        # The last 12 of the 50 outer steps carry most of the strain (near the corner).
        ts_used = ts[-13:]
        # Dense samples of L(t) so the tests can rebuild the callable offline:
        # 33 Chebyshev-Lobatto nodes per outer interval.
        K = 33
        cheb = 0.5 * (1 - np.cos(np.pi * np.arange(K) / (K - 1)))
        tn = np.stack([t0 + (t1 - t0) * cheb for t0, t1 in zip(ts_used[:-1], ts_used[1:])])
        Ln = np.array([[get_L(t, get_x(t)) for t in row] for row in tn])
        F = run_reference([olA, ens], params, get_L, ts_used, get_x, **kw)
        out[tag + "_ts"] = ts_used
        out[tag + "_tnodes"] = tn
        out[tag + "_Lnodes"] = Ln
        out[tag + "_F"] = F
        out[tag + "_ol_o"] = np.stack(olA.orientations)
        out[tag + "_ol_f"] = np.stack(olA.fractions)
        out[tag + "_en_o"] = np.stack(ens.orientations)
        out[tag + "_en_f"] = np.stack(ens.fractions)

    # --- I4: other regimes and parameter variants, one or two steps each
    _, get_L = velocity.simple_shear_2d("Y", "X", 1.0)
    N = 64
    variants = [
        ("i4_diffusion", dict(regime=1), {}, {}),
        ("i4_minvisc", dict(regime=0), {}, {}),
        ("i4_yielding", dict(regime=6), {}, {}),
        ("i4_M0", dict(regime=4), {"gbm_mobility": 0}, {}),
        ("i4_chi0", dict(regime=4), {"gbs_threshold": 0.0}, {}),
        ("i4_chi09", dict(regime=4), {"gbs_threshold": 0.9}, {}),
        ("i4_fabB", dict(regime=4, fabric=1), {}, {}),
        ("i4_fabC", dict(regime=4, fabric=2), {}, {}),
        ("i4_fabD", dict(regime=4, fabric=3), {}, {}),
        ("i4_fabE", dict(regime=4, fabric=4), {}, {}),
        ("i4_exps", dict(regime=4), {"stress_exponent": 1.2, "deformation_exponent": 3.0,
                                     "nucleation_efficiency": 2.5}, {}),
    ]
    for tag, mkw, pkw, kw in variants:
        params = dict(DEFAULT_PARAMS, number_of_grains=N, **pkw)
        m = minerals.Mineral(n_grains=N, seed=99, **mkw)
        ts = np.linspace(0, 0.5, 4)
        F0 = np.eye(3) + 0.1 * np.arange(9).reshape(3, 3) / 9
        F = run_reference([m], params, get_L, ts, zeros, F0=F0, rtol=1e-10, atol=1e-12)
        out[tag + "_meta"] = np.array([m.regime, m.phase, m.fabric], dtype=np.int64)
        out[tag + "_params"] = np.array(
            [params["stress_exponent"], params["deformation_exponent"],
             params["nucleation_efficiency"], params["gbm_mobility"], params["gbs_threshold"]],
            dtype=float,
        )
        out[tag + "_ts"] = ts
        out[tag + "_L"] = get_L(0.0, np.zeros(3))
        out[tag + "_F0"] = F0
        out[tag + "_F"] = F
        out[tag + "_o"] = np.stack(m.orientations)
        out[tag + "_f"] = np.stack(m.fractions)
    np.savez_compressed(os.path.join(HERE, "integrate.npz"), **out)
    print("integrate.npz written")
    return out


def gen_diagnostics(integ):
    out = {"versions": VERSIONS}
    # Voigt averages over the I2 snapshots (two phases, 9 snapshots)
    class _M:  # minimal stand-in carrying what voigt_averages reads
        pass

    ol, en = _M(), _M()
    ol.phase, en.phase = core.MineralPhase.olivine, core.MineralPhase.enstatite
    ol.n_grains = en.n_grains = integ["i2_default_ol_o"].shape[1]
    ol.orientations = list(integ["i2_default_ol_o"][::4])
    ol.fractions = list(integ["i2_default_ol_f"][::4])
    en.orientations = list(integ["i2_default_en_o"][::4])
    en.fractions = list(integ["i2_default_en_f"][::4])
    C = minerals.voigt_averages(
        [ol, en], [core.MineralPhase.olivine, core.MineralPhase.enstatite], [0.7, 0.3]
    )
    out["voigt_ol_o"] = np.stack(ol.orientations)
    out["voigt_ol_f"] = np.stack(ol.fractions)
    out["voigt_en_o"] = np.stack(en.orientations)
    out["voigt_en_f"] = np.stack(en.fractions)
    out["voigt_C"] = C
    # custom (non-symmetric-looking) stiffness to pin the index maps
    rng = np.random.default_rng(5)
    S = rng.random((6, 6)) * 100
    T4 = tensors.voigt_to_elastic_tensor(S)
    R = Rotation.random(1, random_state=3).as_matrix()[0]
    out["voigt_custom_S"] = S
    out["voigt_custom_R"] = R
    out["voigt_custom_rot"] = tensors.elastic_tensor_to_voigt(tensors.rotate(T4, R))

    # M-index: textures of increasing strength, several lattice systems
    tex = {
        "uniform": Rotation.random(120, random_state=8816).as_matrix(),
        "evolved": integ["i1_default_o"][-1][:150],
        "spread10": (
            Rotation.from_rotvec(
                np.deg2rad(10) * np.random.default_rng(1).normal(size=(100, 3)) / 3
            )
        ).as_matrix(),
    }
    systems = {
        "orthorhombic": geometry.LatticeSystem.orthorhombic,
        "triclinic": geometry.LatticeSystem.triclinic,
        "monoclinic": geometry.LatticeSystem.monoclinic,
        "hexagonal": geometry.LatticeSystem.hexagonal,
        "tetragonal": geometry.LatticeSystem.tetragonal,
    }
    # NOTE: LatticeSystem.rhombohedral is omitted: the reference itself fails on it
    # (stats.py:198 `assert False` for bin edges above 104 degrees).
    for tname, o in tex.items():
        out[f"mindex_{tname}_o"] = o
        for sname, system in systems.items():
            if tname != "uniform" and sname not in ("orthorhombic", "hexagonal"):
                continue
            t0 = time.time()
            hist, edges = stats.misorientation_hist(o, system)
            M = diagnostics.misorientation_index(o, system)
            out[f"mindex_{tname}_{sname}_hist"] = hist
            out[f"mindex_{tname}_{sname}_M"] = np.array(M)
            print(f"  mindex {tname}/{sname}: M={M:.6f} ({time.time() - t0:.1f}s)")
    for sname, system in systems.items():
        tmax = stats._max_misorientation(system)
        out[f"theory_{sname}"] = np.array(
            [stats.misorientations_random(i, i + 1, system) for i in range(tmax)]
        )
        ops = geometry.symmetry_operations(system)
        # store every op as the 4x4 matrix acting on scalar-last quaternions, exactly as
        # stats.misorientation_hist applies it (stats.py:118-123; utils.py:365-371)
        mats = []
        for qs in ops:
            if qs.shape == (4, 4):
                mats.append(qs.astype(float))
            else:
                mats.append(
                    np.array([utils.quat_product(qs, e) for e in np.eye(4)]).T.astype(float)
                )
        out[f"symops_{sname}"] = np.array(mats)
    np.savez_compressed(os.path.join(HERE, "diagnostics.npz"), **out)
    print("diagnostics.npz written")


def gen_config1_full():
    """Summary numbers of config 1 at N=5000 (SURVEY.md §8c golden list)."""
    _, get_L = velocity.simple_shear_2d("Y", "X", 1.0)
    params = dict(DEFAULT_PARAMS)
    m = minerals.Mineral(n_grains=5000, seed=8816)
    ts = np.linspace(0, 1, 25)[:6]
    F = run_reference([m], params, get_L, ts, lambda t: np.zeros(3))
    C = minerals.voigt_averages([m], [core.MineralPhase.olivine], [1.0])
    o_res, _ = stats.resample_orientations(
        np.stack(m.orientations[-1:]), np.stack(m.fractions[-1:]), n_samples=300, seed=8816
    )
    M = diagnostics.misorientation_index(o_res[0], geometry.LatticeSystem.orthorhombic)
    out = {
        "versions": VERSIONS,
        "F": F,
        "o0": m.orientations[0],
        "o_final_head": m.orientations[-1][:16],
        "f_final_head": m.fractions[-1][:16],
        "f_final_minmax": np.array([m.fractions[-1].min(), m.fractions[-1].max()]),
        "o_final_abs_sum": np.abs(m.orientations[-1]).sum(axis=0),
        "f_final_sorted_quantiles": np.quantile(m.fractions[-1], [0.01, 0.1, 0.5, 0.9, 0.99]),
        "voigt_C": C,
        "resampled_o": o_res[0],
        "resampled_M": np.array(M),
    }
    np.savez_compressed(os.path.join(HERE, "config1_n5000.npz"), **out)
    print("config1_n5000.npz written; M(resampled 300) =", M)


if __name__ == "__main__":
    t0 = time.time()
    gen_derivatives()
    integ = gen_integrate()
    gen_diagnostics(integ)
    gen_config1_full()
    print("done in %.0f s" % (time.time() - t0))
