"""Small-N run of every kernel of libdrex_b200.so, meant to be executed under
compute-sanitizer (scripts/sanitize.sh).  Sizes are tiny so that memcheck / racecheck /
synccheck finish in minutes; shapes are chosen to hit the ragged paths: odd N (no TMA
prefetch), N < 32, N not a multiple of 32, N = 5000, K = 9 / 33 / 65 velocity-gradient tables,
more systems than CTAs."""

import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__file__)))
import pydrex_b200 as drex  # noqa: E402
from pydrex_b200 import _lib  # noqa: E402

PARAMS = {"phase_assemblage": (0, 1), "phase_fractions": (0.7, 0.3), "stress_exponent": 1.5,
          "deformation_exponent": 3.5, "gbm_mobility": 125, "gbs_threshold": 0.3,
          "nucleation_efficiency": 5.0}
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def integrator():
    for P, N, K in ((3, 31, 1), (2, 333, 1), (2, 334, 9), (1, 5000, 9), (200, 64, 1), (2, 96, 33),
                    (2, 96, 65)):
        ens = drex.ParticleEnsemble(P, N, PARAMS, seed=3)
        for k in range(2):
            if K == 1:
                L = torch.zeros((P, 3, 3), dtype=torch.float64)
                L[:, 1, 0] = 2.0
                ens.update(0.05 * k, 0.05 * (k + 1), L)
            else:
                x = torch.as_tensor(drex.flow.chebyshev_lobatto(K))
                L = torch.zeros((P, K, 3, 3), dtype=torch.float64)
                L[:, :, 1, 0] = 2.0 + 0.5 * (k + x)[None, :]
                L[:, :, 0, 2] = 0.3 * x[None, :]
                ens.update(0.05 * k, 0.05 * (k + 1), L, kind=_lib.L_CHEBYSHEV)
        torch.cuda.synchronize()
        print("integrator", P, N, K, "ok", int(ens.grain_evals.sum()))
    ens = drex.ParticleEnsemble(2, 64, PARAMS, seed=3, regimes=(1, 1))
    L = torch.zeros((2, 3, 3), dtype=torch.float64)
    L[:, 1, 0] = 2.0
    ens.update(0.0, 0.1, L)
    torch.cuda.synchronize()


def others():
    from scipy.spatial.transform import Rotation

    N = 200
    o = Rotation.random(N, random_state=1).as_matrix()
    f = np.full(N, 1 / N)
    L = np.array([[0.0, 0, 0], [2, 0, 0], [0, 0, 0]])
    D = (L + L.T) / 2
    drex.derivatives(4, 0, 0, N, o, f, D, L, np.eye(3), 1.5, 3.5, 5.0, 125.0, 0.7)
    ens = drex.ParticleEnsemble(3, N, PARAMS, seed=3)
    ens.voigt_averages()
    ens.misorientation_indices(drex.LatticeSystem.orthorhombic)
    ens.misorientation_indices(drex.LatticeSystem.orthorhombic, n_samples=100, seeds=[1, 2, 3])
    drex.misorientation_index(o, drex.LatticeSystem.hexagonal)
    drex.resample_orientations(o[None], f[None], n_samples=77, seed=1)
    C = drex.voigt_averages([drex.Mineral(n_grains=N, seed=1)], [0], [1.0])
    drex.elasticity_components(C)
    drex.symmetry_pgr(o)
    drex.bingham_average(o)
    get_v, get_L = drex.velocity.corner_2d("X", "Z", 2.0 / (100 * 365 * 86400))
    drex.pathlines.get_pathline(np.array([1e6, 0.0, -6e4]), get_v, get_L, np.array([0.0, 0.0, -2e5]),
                                np.array([1e6, 0.0, 0.0]), max_strain=3.0, regular_steps=10)
    torch.cuda.synchronize()
    print("others ok")


if which in ("all", "integrator"):
    integrator()
if which in ("all", "others"):
    others()
