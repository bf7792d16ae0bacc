"""Pin the CPU oracle (oracle/) against golden vectors generated from the reference.

These run without a GPU.  They are what lets the `-m gpu` parity tests trust the oracle.
Reference tests mirrored: tests/test_core.py (helpers and derivatives), tests/test_tensors.py
(Voigt index maps), tests/test_diagnostics.py (M-index known answers).
"""

import numpy as np
import pytest
from numpy import testing as nt

OLIVINE, ENSTATITE = 0, 1


def _cases(g):
    for c in range(int(g["n_cases"])):
        yield f"c{c:02d}"


def test_derivatives_match_reference(oracle, golden_derivatives):
    g = golden_derivatives
    for key in _cases(g):
        regime, phase, fabric = (int(x) for x in g[key + "_meta"])
        p, n, lam, M, vf = g[key + "_scal"]
        o, f = g[key + "_o"], g[key + "_f"]
        od, fd = oracle.derivatives(
            regime, phase, fabric, len(f), o, f, g[key + "_D"], g[key + "_L"],
            g[key + "_spin"], p, n, lam, M, vf)
        # reference is numba fastmath: ~1 ulp reassociation noise on O(1) quantities
        nt.assert_allclose(od, g[key + "_od"], rtol=1e-11, atol=1e-13, err_msg=key)
        nt.assert_allclose(fd, g[key + "_fd"], rtol=1e-11, atol=1e-16, err_msg=key)


def test_grain_helpers_match_reference(oracle, golden_derivatives):
    g = golden_derivatives
    for key in _cases(g):
        if key + "_inv" not in g:
            continue
        regime, phase, fabric = (int(x) for x in g[key + "_meta"])
        p, n, lam, M, vf = g[key + "_scal"]
        nt.assert_array_equal(oracle.get_crss(phase, fabric), g[key + "_crss"])
        for i in range(0, len(g[key + "_f"]), 5):
            pr = oracle.grain_probe(phase, fabric, g[key + "_o"][i], g[key + "_D"],
                                    g[key + "_L"], p, n, lam)
            nt.assert_allclose(pr["invariants"], g[key + "_inv"][i], rtol=1e-12, atol=1e-15)
            nt.assert_allclose(pr["slip_rates"], g[key + "_rates"][i], rtol=1e-11, atol=1e-15)
            nt.assert_allclose(pr["deformation_rate"], g[key + "_defr"][i], rtol=1e-11,
                               atol=1e-14)
            nt.assert_allclose(pr["slip_rate_softest"], g[key + "_soft"][i], rtol=1e-11,
                               atol=1e-14)
            nt.assert_allclose(pr["strain_energy"], g[key + "_energy"][i], rtol=1e-11,
                               atol=1e-15)


def test_get_crss_errors(oracle):
    with pytest.raises(ValueError):
        oracle.get_crss(OLIVINE, 5)
    with pytest.raises(ValueError):
        oracle.get_crss(ENSTATITE, 0)
    with pytest.raises(ValueError):
        oracle.get_crss(2, 0)


def test_unsupported_regimes_raise(oracle):
    o = np.eye(3)[None]
    f = np.ones(1)
    for regime in (2, 3, 5, 8, -1):
        with pytest.raises(ValueError):
            oracle.derivatives(regime, 0, 0, 1, o, f, np.eye(3), np.eye(3), np.eye(3),
                               1.5, 3.5, 5.0, 125.0, 1.0)


def _const_L(L):
    return lambda t, x: L


@pytest.mark.parametrize("tag,tol", [("i1_tight", 1e-8), ("i1_default", 1e-9)])
def test_update_orientations_config1(oracle, golden_integrate, tag, tol):
    """Oracle LSODA driver vs reference Mineral.update_orientations, same tolerances.

    Same scipy LSODA, RHS equal to ~1e-15: trajectories agree far below the solver
    tolerance.  (Exact equality is not expected: the step-size controller amplifies
    last-bit RHS differences.)
    """
    g = golden_integrate
    kw = {"rtol": 1e-10, "atol": 1e-12} if tag.endswith("tight") else {}
    params = {"phase_assemblage": (0,), "phase_fractions": (1.0,), "stress_exponent": 1.5,
              "deformation_exponent": 3.5, "gbm_mobility": 125, "gbs_threshold": 0.3,
              "nucleation_efficiency": 5.0}
    ts = g[tag + "_ts"]
    o, f, F = g[tag + "_o"][0], g[tag + "_f"][0], np.eye(3)
    for k, (t0, t1) in enumerate(zip(ts[:-1], ts[1:])):
        F, o, f = oracle.update_orientations(
            o, f, 4, 0, 0, params, F, _const_L(g[tag + "_L"]), (t0, t1, lambda t: np.zeros(3)),
            **dict(kw))
        nt.assert_allclose(o, g[tag + "_o"][k + 1], rtol=0, atol=tol)
        nt.assert_allclose(f, g[tag + "_f"][k + 1], rtol=tol * 10, atol=1e-12)
    nt.assert_allclose(F, g[tag + "_F"], rtol=1e-9, atol=1e-12)


def test_update_orientations_two_phase(oracle, golden_integrate):
    g = golden_integrate
    tag = "i2_tight"
    params = {"phase_assemblage": (0, 1), "phase_fractions": (0.7, 0.3), "stress_exponent": 1.5,
              "deformation_exponent": 3.5, "gbm_mobility": 125, "gbs_threshold": 0.3,
              "nucleation_efficiency": 5.0}
    L1, L2, tsw = g[tag + "_L1"], g[tag + "_L2"], float(g[tag + "_tswitch"])
    get_L = lambda t, x: L1 if t < tsw else L2  # noqa: E731
    ts = g[tag + "_ts"]
    oo, of = g[tag + "_ol_o"][0], g[tag + "_ol_f"][0]
    eo, ef = g[tag + "_en_o"][0], g[tag + "_en_f"][0]
    F = np.eye(3)
    pos = lambda t: np.full(3, np.nan)  # noqa: E731
    for k, (t0, t1) in enumerate(zip(ts[:-1], ts[1:])):
        _, eo, ef = oracle.update_orientations(eo, ef, 4, 1, 5, params, F, get_L, (t0, t1, pos),
                                               rtol=1e-10, atol=1e-12)
        F, oo, of = oracle.update_orientations(oo, of, 4, 0, 0, params, F, get_L, (t0, t1, pos),
                                               rtol=1e-10, atol=1e-12)
        nt.assert_allclose(oo, g[tag + "_ol_o"][k + 1], rtol=0, atol=2e-8)
        nt.assert_allclose(eo, g[tag + "_en_o"][k + 1], rtol=0, atol=2e-8)
        nt.assert_allclose(of, g[tag + "_ol_f"][k + 1], rtol=1e-7, atol=1e-12)
        # enstatite strain energy is identically zero: fractions only move through GBS
        nt.assert_allclose(ef, g[tag + "_en_f"][k + 1], rtol=1e-12, atol=1e-15)
    nt.assert_allclose(F, g[tag + "_F"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize(
    "tag", ["i4_diffusion", "i4_minvisc", "i4_yielding", "i4_M0", "i4_chi0", "i4_chi09",
            "i4_fabB", "i4_fabC", "i4_fabD", "i4_fabE", "i4_exps"])
def test_update_orientations_variants(oracle, golden_integrate, tag):
    g = golden_integrate
    regime, phase, fabric = (int(x) for x in g[tag + "_meta"])
    p, n, lam, M, chi = g[tag + "_params"]
    params = {"phase_assemblage": (0,), "phase_fractions": (1.0,), "stress_exponent": p,
              "deformation_exponent": n, "gbm_mobility": M, "gbs_threshold": chi,
              "nucleation_efficiency": lam}
    ts = g[tag + "_ts"]
    o, f, F = g[tag + "_o"][0], g[tag + "_f"][0], g[tag + "_F0"]
    for k, (t0, t1) in enumerate(zip(ts[:-1], ts[1:])):
        F, o, f = oracle.update_orientations(
            o, f, regime, phase, fabric, params, F, _const_L(g[tag + "_L"]),
            (t0, t1, lambda t: np.zeros(3)), rtol=1e-10, atol=1e-12)
        nt.assert_allclose(o, g[tag + "_o"][k + 1], rtol=0, atol=2e-8, err_msg=tag)
        nt.assert_allclose(f, g[tag + "_f"][k + 1], rtol=1e-7, atol=1e-12, err_msg=tag)
    nt.assert_allclose(F, g[tag + "_F"], rtol=1e-9, atol=1e-12)


def test_voigt_average_matches_reference(oracle, golden_diagnostics):
    g = golden_diagnostics
    S_ol = np.array([[320.71, 69.84, 71.22, 0, 0, 0], [69.84, 197.25, 74.8, 0, 0, 0],
                     [71.22, 74.8, 234.32, 0, 0, 0], [0, 0, 0, 63.77, 0, 0],
                     [0, 0, 0, 0, 77.67, 0], [0, 0, 0, 0, 0, 78.36]])
    S_en = np.array([[236.9, 79.6, 63.2, 0, 0, 0], [79.6, 180.5, 56.8, 0, 0, 0],
                     [63.2, 56.8, 230.4, 0, 0, 0], [0, 0, 0, 84.3, 0, 0],
                     [0, 0, 0, 0, 79.4, 0], [0, 0, 0, 0, 0, 80.1]])
    for i in range(len(g["voigt_C"])):
        C = np.zeros((6, 6))
        oracle.voigt_average(g["voigt_ol_o"][i], g["voigt_ol_f"][i], S_ol, 0.7, out=C)
        oracle.voigt_average(g["voigt_en_o"][i], g["voigt_en_f"][i], S_en, 0.3, out=C)
        nt.assert_allclose(C, g["voigt_C"][i], rtol=1e-12, atol=1e-10)
    V = oracle.rotate_voigt(g["voigt_custom_S"], g["voigt_custom_R"])
    nt.assert_allclose(V, g["voigt_custom_rot"], rtol=1e-12, atol=1e-11)


SYSTEMS = {"triclinic": (1, 1), "monoclinic": (2, 2), "orthorhombic": (2, 4),
           "tetragonal": (4, 8), "hexagonal": (6, 12)}


def test_symmetry_ops_and_theory(oracle, golden_diagnostics):
    g = golden_diagnostics
    for name, system in SYSTEMS.items():
        nt.assert_allclose(oracle.symmetry_ops(system), g[f"symops_{name}"], rtol=0, atol=1e-16)
        theory = g[f"theory_{name}"]
        mine = np.array([oracle.misorientations_random(i, i + 1, system)
                         for i in range(len(theory))])
        nt.assert_allclose(mine, theory, rtol=1e-12, atol=1e-15)


def test_misorientation_index_matches_reference(oracle, golden_diagnostics):
    g = golden_diagnostics
    n = 0
    for key in g.files:
        if not (key.startswith("mindex_") and key.endswith("_M")):
            continue
        _, tname, sname, _ = key.split("_")
        o = g[f"mindex_{tname}_o"]
        M, hist, counts = oracle.misorientation_index(o, SYSTEMS[sname], return_hist=True)
        ref_hist = g[f"mindex_{tname}_{sname}_hist"]
        npairs = counts.sum()
        # float32 acos/rad2deg chains differ by an ulp between libm and numba: allow a pair
        # or two to fall on the other side of a bin edge.
        assert np.abs(hist - ref_hist).max() * npairs <= 2.5, (tname, sname)
        assert abs(M - float(g[key])) <= 2.5 / npairs, (tname, sname)
        n += 1
    assert n >= 9


def test_resample_orientations_golden(oracle, golden_resample, golden_integrate):
    """stats.resample_orientations (stats.py:14-64): exact where the reference is defined
    independently of its argsort's tie order."""
    g = golden_resample
    o, f = g["d_orientations"], g["d_fractions"]
    for tag in ("default", "n2000", "n10"):
        n_samples, seed = (int(v) for v in g[f"d_{tag}_args"])
        ro, rf = oracle.resample_orientations(o, f, None if n_samples < 0 else n_samples, seed)
        assert np.array_equal(ro, g[f"d_{tag}_o"])
        assert np.array_equal(rf, g[f"d_{tag}_f"])
    i = golden_integrate
    eo, ef = i["i1_default_o"], i["i1_default_f"]
    ro, rf = oracle.resample_orientations(eo, ef, 1000, 8816)
    assert np.array_equal(rf, g["e_f"])  # tied grains carry the same fraction
    for t in range(len(ef)):
        vals, counts = np.unique(ef[t], return_counts=True)
        untied = np.isin(rf[t], vals[counts == 1])
        assert np.array_equal(ro[t][untied], g["e_o"][t][untied])


FLOW_NAMES = ("shear", "shear_zx", "cell", "cell_yx", "corner", "corner_zy")
PATH_CASES = {"corner": 50, "cell": 20, "shear": 10}


def test_flow_fields_match_reference(oracle, golden_pathlines):
    """velocity.py:17-97 closed forms (the reference is numba fastmath: ~1 ulp noise)."""
    g = golden_pathlines
    for name in FLOW_NAMES:
        desc, x = g[f"f_{name}_desc"], g[f"f_{name}_x"]
        for xi, v_ref, L_ref in zip(x, g[f"f_{name}_v"], g[f"f_{name}_L"]):
            v, L = oracle.flow_eval(desc, xi)
            nt.assert_allclose(v, v_ref, rtol=1e-13, atol=1e-300, err_msg=name)
            nt.assert_allclose(L, L_ref, rtol=1e-13, atol=1e-300, err_msg=name)


def test_pathlines_match_reference(oracle, golden_pathlines):
    """pathlines.get_pathline (pathlines.py:10-116).  The reference's positions carry LSODA's
    rtol = 1e-5 and its start time additionally a rectangle-rule strain sum and a crude
    boundary event, so the comparison is made at that accuracy: the same point of the same
    streamline at the same time."""
    from scipy.interpolate import CubicSpline

    g = golden_pathlines
    for name, steps in PATH_CASES.items():
        desc, box = g[f"p_{name}_desc"], g[f"p_{name}_box"]
        scale = np.abs(box).max()
        max_strain = float(g[f"p_{name}_max_strain"])
        for final, t_ref, x_ref in zip(g[f"p_{name}_final"], g[f"p_{name}_t"], g[f"p_{name}_x"]):
            ts, xs, strains = oracle.trace_pathline(desc, final, box[0], box[1], max_strain, steps)
            assert ts.shape == t_ref.shape and ts[-1] == 0.0
            print(name, final, "t_start oracle/reference - 1 =", ts[0] / t_ref[0] - 1)
            nt.assert_allclose(ts[0], t_ref[0], rtol=2e-2, err_msg=name)
            assert strains[0] == 0.0 and strains[-1] <= max_strain * (1 + 1e-9)
            # positions along the common part of the path, at the reference's own times
            common = t_ref >= ts[0]
            fine_t, fine_x, _ = oracle.trace_pathline(desc, final, box[0], box[1], max_strain, 2000)
            spline = CubicSpline(fine_t, fine_x)
            nt.assert_allclose(spline(t_ref[common]), x_ref[common], atol=2e-4 * scale, rtol=0,
                               err_msg=name)
