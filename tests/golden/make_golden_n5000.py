"""Golden vectors at the BENCHMARKED ensemble size (N = 5000 grains) from the UNMODIFIED reference.

Run in the build container only (`python tests/golden/make_golden_n5000.py`); imports the
reference from /root/reference/src through `_refharness`.  Writes `n5000.npz` next to this
file.  The GPU box never runs this script.

Why: every other integrated-state fixture has N <= 400, while the reference drivers this path
replaces run N = 5000 (tests/test_corner_flow_2d.py:91-188, tests/test_simple_shear_3d.py:112-141)
and the device integrator places its per-system data differently at that size.

Cases (reference file:line), each stored as full final arrays (float64):
- c1a  config 1 (tests/test_simple_shear_2d.py:255-272 set-up): olivine A, simple shear
       `simple_shear_2d("Y","X",1.0)`, seed 8816, 2 outer intervals of linspace(0,1,25),
       LSODA rtol 1e-10 / atol 1e-12 (parity mode A).
- c1b  same, 6 intervals at the reference's default tolerances (parity mode B) and, beside
       it, the same 6 intervals at tight tolerances (`c1t`): the reference's own error band.
- c2a / c2b  config 2 (tests/test_simple_shear_3d.py:112-141 intent): enstatite AB then
       olivine A, phase fractions 0.3 / 0.7, L1 = simple_shear_2d("X","Z",5e-7) switching to
       L2 = simple_shear_2d("Y","X",5e-7) at t = 2e5, 4 intervals of 1e5 s; tight / default.
- c3a / c3b  config 3 (tests/test_corner_flow_2d.py:107-148): corner flow, pathline ending at
       z = -0.3 H, olivine A, the last 6 of 50 outer intervals; tight (2 intervals) / default (6).
       L(t) along the pathline is stored at 33 Chebyshev-Lobatto nodes per interval so the tests
       can rebuild the callable without the reference.
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
from pydrex import core, minerals, velocity  # noqa: E402
from pydrex import pathlines as ref_pathlines  # noqa: E402

VERSIONS = json.dumps(
    {"numpy": np.__version__, "scipy": scipy.__version__, "numba": numba.__version__}
)
N = 5000
TIGHT = {"rtol": 1e-10, "atol": 1e-12}
PARAMS = {
    "phase_assemblage": (core.MineralPhase.olivine,),
    "phase_fractions": (1.0,),
    "stress_exponent": 1.5,
    "deformation_exponent": 3.5,
    "gbm_mobility": 125,
    "gbs_threshold": 0.3,
    "nucleation_efficiency": 5.0,
    "number_of_grains": N,
}


def run(mins, params, get_L, ts, get_x, **kw):
    F = np.eye(3)
    for t0, t1 in zip(ts[:-1], ts[1:]):
        F = minerals.update_all(mins, params, F, get_L, (t0, t1, get_x), **kw)
    return F


def main():
    out = {"versions": VERSIONS}
    t_all = time.time()
    zeros = lambda t: np.zeros(3)  # noqa: E731

    # ---- config 1
    _, get_L = velocity.simple_shear_2d("Y", "X", 1.0)
    ts = np.linspace(0, 1, 25)
    out["c1_L"] = get_L(0.0, np.zeros(3))
    out["c1_ts"] = ts[:7]
    for tag, n_int, kw in (("c1a", 2, TIGHT), ("c1b", 6, {}), ("c1t", 6, TIGHT)):
        t0 = time.time()
        m = minerals.Mineral(n_grains=N, seed=8816)
        F = run([m], PARAMS, get_L, ts[: n_int + 1], zeros, **kw)
        if tag == "c1a":
            out["c1_o0"] = m.orientations[0]
        out[tag + "_F"] = F
        out[tag + "_o"] = m.orientations[-1]
        out[tag + "_f"] = m.fractions[-1]
        print(f"{tag}: {time.time() - t0:.1f} s", flush=True)

    # ---- config 2
    _, L1 = velocity.simple_shear_2d("X", "Z", 5e-7)
    _, L2 = velocity.simple_shear_2d("Y", "X", 5e-7)
    t_switch = 2e5

    def get_L2(t, x):
        return L1(t, x) if t < t_switch else L2(t, x)

    params2 = dict(
        PARAMS,
        phase_assemblage=(core.MineralPhase.olivine, core.MineralPhase.enstatite),
        phase_fractions=(0.7, 0.3),
    )
    ts2 = np.linspace(0, 4e5, 5)
    out["c2_ts"] = ts2
    out["c2_L1"] = L1(0.0, np.zeros(3))
    out["c2_L2"] = L2(0.0, np.zeros(3))
    out["c2_tswitch"] = np.array(t_switch)
    for tag, kw in (("c2a", TIGHT), ("c2b", {})):
        t0 = time.time()
        ens = minerals.Mineral(1, 5, 4, n_grains=N, seed=245623)
        olA = minerals.Mineral(0, 0, 4, n_grains=N, seed=245623)
        F = run([ens, olA], params2, get_L2, ts2, lambda t: np.full(3, np.nan), **kw)
        if tag == "c2a":
            out["c2_o0"] = olA.orientations[0]
        out[tag + "_F"] = F
        out[tag + "_ol_o"] = olA.orientations[-1]
        out[tag + "_ol_f"] = olA.fractions[-1]
        out[tag + "_en_o"] = ens.orientations[-1]
        out[tag + "_en_f"] = ens.fractions[-1]
        print(f"{tag}: {time.time() - t0:.1f} s", flush=True)

    # ---- config 3
    plate_speed = 2.0 / (100 * 365 * 86400)
    get_u, get_Lc = velocity.corner_2d("X", "Z", plate_speed)
    H, W = 2e5, 1e6
    tsp, get_x = ref_pathlines.get_pathline(
        np.array([W, 0.0, -0.3 * H]), get_u, get_Lc,
        np.array([0.0, 0.0, -H]), np.array([W, 0.0, 0.0]),
        max_strain=10.0, regular_steps=50,
    )
    ts3 = tsp[-7:]
    K = 33
    cheb = 0.5 * (1 - np.cos(np.pi * np.arange(K) / (K - 1)))
    tn = np.stack([a + (b - a) * cheb for a, b in zip(ts3[:-1], ts3[1:])])
    out["c3_ts"] = ts3
    out["c3_tnodes"] = tn
    out["c3_Lnodes"] = np.array([[get_Lc(t, get_x(t)) for t in row] for row in tn])
    for tag, n_int, kw in (("c3a", 2, TIGHT), ("c3b", 6, {})):
        t0 = time.time()
        m = minerals.Mineral(n_grains=N, seed=8816)
        F = run([m], PARAMS, get_Lc, ts3[: n_int + 1], get_x, **kw)
        out[tag + "_F"] = F
        out[tag + "_o"] = m.orientations[-1]
        out[tag + "_f"] = m.fractions[-1]
        print(f"{tag}: {time.time() - t0:.1f} s", flush=True)

    np.savez_compressed(os.path.join(HERE, "n5000.npz"), **out)
    print("n5000.npz written in %.0f s" % (time.time() - t_all))


if __name__ == "__main__":
    main()
