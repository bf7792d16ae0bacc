"""The reference's own known-answer tests, run on the CUDA path through the drop-in API.

Closed-form answers, no fixtures: they catch what random-grain comparisons cannot (slip-system
switching angles, exact ties, exact zeros).  Each test cites the reference test it mirrors
(paths relative to /root/reference/tests/); tolerances are the reference's, except that exact
zeros are compared with atol 1e-14 (fused multiply-adds leave 1e-17 where numpy leaves 0)."""

import numpy as np
import pytest
from numpy import testing as nt
from scipy.spatial.transform import Rotation

pytestmark = pytest.mark.gpu

SEED = 8816  # conftest.py:246-249
OLIVINE_SLIP_SYSTEMS = (([0, 1, 0], [1, 0, 0]), ([0, 0, 1], [1, 0, 0]), ([0, 1, 0], [0, 0, 1]),
                        ([1, 0, 0], [0, 0, 1]))  # minerals.py:25-30


@pytest.fixture(scope="module")
def drex():
    import pydrex_b200

    return pydrex_b200


def _derivatives(drex, orientations, L, gbm_mobility=0.0):
    n = len(orientations)
    D = (L + L.T) / 2
    return drex.derivatives(
        regime=drex.DeformationRegime.matrix_dislocation, phase=drex.MineralPhase.olivine,
        fabric=drex.MineralFabric.olivine_A, n_grains=n, orientations=orientations,
        fractions=np.full(n, 1.0 / n), strain_rate=D, velocity_gradient=L,
        deformation_gradient_spin=np.full((3, 3), np.nan), stress_exponent=1.5,
        deformation_exponent=3.5, nucleation_efficiency=5, gbm_mobility=gbm_mobility,
        volume_fraction=1.0)


@pytest.mark.parametrize("case", ["dvdx_010_100", "dudz_001_100", "dwdx_001_100", "dvdz_010_001"])
def test_olivine_a_single_grain_rotation_rates(drex, case):
    """test_core.py:110-548 (TestDislocationCreepOlivineA): analytic rotation rate of one A-type
    grain against its angle to the shear, four shear geometries; the dominant slip system and the
    private helpers are exercised on every tenth angle like the reference does on every angle."""
    from pydrex_b200 import core

    n = 3600 if case == "dvdx_010_100" else 360
    theta = np.mgrid[0 : 2 * np.pi : n * 1j]
    c, s, c2 = np.cos(theta), np.sin(theta), np.cos(2 * theta)
    z = np.zeros(n)
    if case == "dvdx_010_100":
        L = np.array([[0.0, 0, 0], [2, 0, 0], [0, 0, 0]])
        o = Rotation.from_rotvec(np.column_stack([z, z, theta])).as_matrix()
        target = np.array([[s * (1 + c2), c * (1 + c2), z], [-c * (1 + c2), s * (1 + c2), z], [z, z, z]])
        system = ([0, 1, 0], [1, 0, 0])
    elif case == "dudz_001_100":
        L = np.array([[0.0, 0, 2], [0, 0, 0], [0, 0, 0]])
        o = Rotation.from_rotvec(np.column_stack([z, theta, z])).as_matrix()
        target = np.array([[-s * (c2 - 1), z, c * (c2 - 1)], [z, z, z], [c * (1 - c2), z, s * (1 - c2)]])
        system = ([0, 0, 1], [1, 0, 0])
    elif case == "dwdx_001_100":
        L = np.array([[0.0, 0, 0], [0, 0, 0], [2, 0, 0]])
        o = Rotation.from_rotvec(np.column_stack([z, theta, z])).as_matrix()
        target = np.array([[-s * (1 + c2), z, c * (1 + c2)], [z, z, z], [-c * (1 + c2), z, -s * (1 + c2)]])
        system = ([0, 0, 1], [1, 0, 0])
    else:
        L = np.array([[0.0, 0, 0], [0, 0, 2], [0, 0, 0]])
        o = Rotation.from_rotvec(np.column_stack([theta, z, z])).as_matrix()
        target = np.array([[z, z, z], [z, -s * (1 + c2), -c * (1 + c2)], [z, c * (1 + c2), -s * (1 + c2)]])
        system = ([0, 1, 0], [0, 0, 1])
    target = np.moveaxis(target, -1, 0)
    od, fd = _derivatives(drex, o, L)
    nt.assert_allclose(od, target, rtol=1e-7, atol=1e-14)
    assert np.isclose(fd.sum(), 0.0)
    D = (L + L.T) / 2
    crss = core.get_crss(drex.MineralPhase.olivine, drex.MineralFabric.olivine_A)
    for g in range(0, n, 10):
        inv = core._get_slip_invariants(D, o[g])
        idx = np.argsort(np.abs(inv / crss))
        if np.abs(inv).max() > 1e-12:  # at the zero crossings every system is tied at 0
            assert OLIVINE_SLIP_SYSTEMS[idx[-1]] == system, (case, np.rad2deg(theta[g]))
        rates = core._get_slip_rates_olivine(inv, idx, crss, 3.5)
        G = core._get_deformation_rate(drex.MineralPhase.olivine, o[g], rates)
        assert np.all(np.isfinite(G)) and abs(np.trace(G)) < 1e-12


def test_enstatite_single_grain_rotation_rates(drex):
    """test_core.py:20-106 (TestDislocationCreepOPX), both shear geometries, 361 angles."""
    theta = np.linspace(0, 2 * np.pi, 361)
    c, s, c2 = np.cos(theta), np.sin(theta), np.cos(2 * theta)
    z = np.zeros_like(theta)
    for L, o, target in (
        (np.array([[0.0, 0, 2], [0, 0, 0], [0, 0, 0]]),
         Rotation.from_rotvec(np.column_stack([z, theta, z])).as_matrix(),
         (1 + c2)[:, None, None] * np.moveaxis(np.array([[s, z, -c], [z, z, z], [c, z, s]]), -1, 0)),
        (np.array([[0.0, 0, 0], [2, 0, 0], [0, 0, 0]]),
         Rotation.from_rotvec(np.column_stack([z, z, theta])).as_matrix(),
         np.moveaxis(np.array([[s, c, z], [-c, s, z], [z, z, z]]), -1, 0)),
    ):
        n = len(theta)
        od, fd = drex.derivatives(
            regime=drex.DeformationRegime.matrix_dislocation, phase=drex.MineralPhase.enstatite,
            fabric=drex.MineralFabric.enstatite_AB, n_grains=n, orientations=o,
            fractions=np.full(n, 1.0 / n), strain_rate=(L + L.T) / 2, velocity_gradient=L,
            deformation_gradient_spin=np.full((3, 3), np.nan), stress_exponent=1.5,
            deformation_exponent=3.5, nucleation_efficiency=5, gbm_mobility=0, volume_fraction=1.0)
        # the reference checks dudz with assert_allclose defaults and dvdx with atol 1e-15
        nt.assert_allclose(od, target, rtol=1e-7, atol=1e-14)
        assert np.isclose(fd.sum(), 0.0)


def test_recrystallisation_circle_inplane(drex):
    """test_core.py:591-685: 360000 A-type grains with a-axes on a circle in the shear plane;
    grain growth rates against the analytic strain energies, atol 1e-15."""
    n = 360000
    theta = np.mgrid[0 : 2 * np.pi : n * 1j]
    o = Rotation.from_rotvec(np.column_stack([np.zeros(n), np.zeros(n), theta])).as_matrix()
    L = np.array([[0.0, 0, 0], [2, 0, 0], [0, 0, 0]])
    od, fd = _derivatives(drex, o, L, gbm_mobility=125)
    rho = np.abs(np.cos(2 * theta)) ** (1.5 / 3.5)
    E = rho * np.exp(-5 * rho**2)
    nt.assert_allclose(fd, -125 / n * (E - E.mean()), atol=1e-15, rtol=0)
    nt.assert_allclose(np.linalg.norm(od[:, 0], axis=1), 1 + np.cos(2 * theta), atol=1e-12)


def test_recrystallisation_circle_shearplane_switching_angles(drex):
    """test_core.py:687-774: a-axes in the YZ plane, two slip systems alternate; the dominant one
    is checked every degree against the switching angles 64 / 117 / 244 / 297 degrees."""
    from pydrex_b200 import core

    D = np.array([[0.0, 1, 0], [1, 0, 0], [0, 0, 0]])
    crss = core.get_crss(drex.MineralPhase.olivine, drex.MineralFabric.olivine_A)
    for theta in np.mgrid[0 : 2 * np.pi : 360000j][::1000]:
        inv = core._get_slip_invariants(D, Rotation.from_euler("zx", [np.pi / 2, theta]).as_matrix())
        system = OLIVINE_SLIP_SYSTEMS[np.argsort(np.abs(inv / crss))[-1]]
        deg = np.rad2deg(theta)
        if 64 <= deg < 117 or 244 <= deg < 297:
            assert system == ([0, 0, 1], [1, 0, 0]), deg
        else:
            assert system == ([0, 1, 0], [1, 0, 0]), deg
    # and the whole circle through derivatives: finite, fractions conserved
    n = 36000
    theta = np.mgrid[0 : 2 * np.pi : n * 1j]
    o = Rotation.from_euler("zx", np.column_stack([np.full(n, np.pi / 2), theta])).as_matrix()
    od, fd = _derivatives(drex, o, np.array([[0.0, 0, 0], [2, 0, 0], [0, 0, 0]]), gbm_mobility=125)
    assert np.all(np.isfinite(od)) and abs(fd.sum()) < 1e-12


def test_mindex_known_answers(drex):
    """test_diagnostics.py:257-354 (TestMIndex): uniform 0.05, 10 degree spread 0.99, 45 degree
    spread 0.81, increasing with texture strength, girdle 0.67 (atol 1e-2 each)."""
    system = drex.LatticeSystem.orthorhombic
    M = lambda o: drex.misorientation_index(o, system)  # noqa: E731
    assert np.isclose(M(Rotation.random(1000, random_state=SEED).as_matrix()), 0.05, atol=1e-2, rtol=0)
    x = np.random.default_rng(seed=SEED).random(1000)

    def spread(factor, invert=False):
        ang = x * np.pi / factor - np.pi / factor / 2
        r = Rotation.from_rotvec(np.column_stack([np.zeros(1000), ang, ang]))
        return (r.inv() if invert else r).as_matrix()

    assert np.isclose(M(spread(18, invert=True)), 0.99, atol=1e-2, rtol=0)
    assert np.isclose(M(spread(2)), 0.81, atol=1e-2, rtol=0)
    vals = [M(spread(f)) for f in (2, 3, 4, 5)]
    assert np.all(np.diff(vals) >= 0), vals
    rng = np.random.default_rng(seed=SEED)
    cq, dq = rng.normal(0, 1.0, size=1000), rng.normal(0, 1.0, size=1000)
    girdle = Rotation.from_quat(np.column_stack([np.zeros(1000), np.zeros(1000), cq, dq])).as_matrix()
    assert np.isclose(M(girdle), 0.67, atol=1e-2)


def test_elasticity_components_browaeys_chevrot(drex):
    """test_diagnostics.py:16-69: symmetry decomposition of the olivine and enstatite tensors of
    Browaeys & Chevrot 2004."""
    for C, expected in (
        (np.array([[192, 66, 60, 0, 0, 0], [66, 160, 56, 0, 0, 0], [60, 56, 272, 0, 0, 0],
                   [0, 0, 0, 60, 0, 0], [0, 0, 0, 0, 62, 0], [0, 0, 0, 0, 0, 49.0]]),
         {"bulk_modulus": 109.8, "shear_modulus": 63.7, "percent_anisotropy": 20.7,
          "percent_monoclinic": 0, "percent_triclinic": 0}),
        (np.array([[225, 54, 72, 0, 0, 0], [54, 214, 53, 0, 0, 0], [72, 53, 178, 0, 0, 0],
                   [0, 0, 0, 78, 0, 0], [0, 0, 0, 0, 82, 0], [0, 0, 0, 0, 0, 76.0]]),
         {"bulk_modulus": 108.3, "shear_modulus": 76.4, "percent_anisotropy": 9.2,
          "percent_monoclinic": 0, "percent_triclinic": 0}),
    ):
        out = drex.elasticity_components([C])
        for k, v in expected.items():
            assert np.isclose(out[k][0], v, atol=0.1, rtol=0), f"{k}: {out[k]} != {v}"
        nt.assert_allclose(np.abs(out["hexagonal_axis"][0]), [0, 0, 1], atol=1e-12)


@pytest.mark.parametrize("gbm_mobility", [50, 100, 150])
def test_grainsize_median(drex, gbm_mobility):
    """test_simple_shear_2d.py:274-299: with M* in {50, 100, 150} and lambda* = 5 the median
    grain size decreases from the second step on."""
    params = drex.DefaultParams().as_dict()
    params["gbm_mobility"] = gbm_mobility
    params["nucleation_efficiency"] = 5
    timestamps = np.linspace(0, 1, 25)
    _, get_L = drex.velocity.simple_shear_2d("Y", "X", 1)
    m = drex.Mineral(phase=drex.MineralPhase.olivine, fabric=drex.MineralFabric.olivine_A,
                     regime=drex.DeformationRegime.matrix_dislocation,
                     n_grains=params["number_of_grains"], seed=SEED)
    F = np.eye(3)
    for t0, t1 in zip(timestamps[:-1], timestamps[1:]):
        F = m.update_orientations(params, F, get_L, (t0, t1, lambda t: np.zeros(3)))
    medians = np.array([np.median(f) for f in m.fractions])
    nt.assert_array_less(np.diff(medians)[1:], np.full(len(timestamps) - 2, 0))


def test_diffusion_creep_keeps_grain_sizes_and_symmetry(drex):
    """test_vortex_2d.py:389-455 (TestDiffusionCreep): in the matrix-diffusion regime the
    surrogate grain sizes do not change (atol 1e-16) and random / girdled / clustered textures
    keep their symmetry (P, G, R within 0.25 between steps; R, G, P > 0.9 respectively), around
    one revolution of the 2-D convection cell."""
    params = drex.DefaultParams().as_dict()
    params["gbm_mobility"] = 10
    N = params["number_of_grains"]
    rng = np.random.default_rng(seed=SEED)
    inits = [None,
             Rotation.from_euler("y", [[x * np.pi * 2] for x in rng.random(N)]).as_matrix(),
             Rotation.from_euler("y", [[x * np.pi / 8] for x in rng.random(N)]).as_matrix()]
    get_u, get_L = drex.velocity.cell_2d("X", "Z", 6.342e-10, 2e5)
    timestamps, get_x = drex.pathlines.get_pathline(
        np.array([1.5e4, 0.0, -8e4]), get_u, get_L, np.array([-1e5, 0.0, -1e5]),
        np.array([1e5, 0.0, 1e5]), max_strain=7.0, regular_steps=20)
    for i, o_init in enumerate(inits):
        m = drex.Mineral(phase=drex.MineralPhase.olivine, fabric=drex.MineralFabric.olivine_A,
                         regime=drex.DeformationRegime.matrix_diffusion, n_grains=N, seed=SEED,
                         orientations_init=o_init)
        F = np.eye(3)
        for t0, t1 in zip(timestamps[:-1], timestamps[1:]):
            F = m.update_orientations(params, F, get_L, (t0, t1, get_x))
            nt.assert_allclose(m.fractions[-1], m.fractions[-2], atol=1e-16, rtol=0)
            pgr = np.array(drex.symmetry_pgr(m.orientations[-1]))
            nt.assert_allclose(pgr, drex.symmetry_pgr(m.orientations[-2]), atol=0.25, rtol=0)
            assert pgr[(2, 1, 0)[i]] > 0.9, (i, pgr)
