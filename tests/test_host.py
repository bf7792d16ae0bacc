"""CPU-only tests of the host side: the C ABI loads and exports what include/*.h declares,
the product fails loudly without a GPU (no CPU fallback), the L(t) tabulation, and the
particle sharding / gather logic over a world_size-2 gloo group."""

import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
from numpy import testing as nt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "drex_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drex_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pydrex_b200 import _lib

    _lib.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/drex_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    major, minor = ctypes.c_int(), ctypes.c_int()
    assert lib.drex_version(ctypes.byref(major), ctypes.byref(minor)) == 0
    assert (major.value, minor.value) == (0, 2)


def test_host_side_argument_validation_without_gpu():
    """Argument checks run before any launch, so they are testable without a device."""
    from pydrex_b200 import _lib

    lib = _lib.load()
    prm = _lib.DrexParams(1.5, 3.5, 5.0, 125.0, 0.3)
    one = np.zeros(1, dtype=np.int32)
    bad_regime = np.array([3], dtype=np.int32)
    rc = lib.drex_derivatives_f64(None, None, None, None, None, None, None, None, None,
                                  _lib.ptr(bad_regime), _lib.ptr(one), _lib.ptr(one),
                                  ctypes.byref(prm), 1, 8, None, None, None, 0, None)
    assert rc == _lib.DREX_E_UNSUPPORTED
    assert "not yet supported" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    theory = (ctypes.c_double * 180)()
    assert lib.drex_mindex_theory_host(2, 4, theory) == 120
    assert lib.drex_mindex_theory_host(3, 6, theory) == _lib.DREX_E_UNSUPPORTED
    assert lib.drex_mindex_theory_host(5, 5, theory) == _lib.DREX_E_UNSUPPORTED


def test_no_cpu_fallback():
    import torch

    import pydrex_b200 as drex

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        drex.derivatives(4, 0, 0, 1, np.eye(3)[None], [1.0], np.eye(3), np.eye(3), np.eye(3),
                         1.5, 3.5, 5.0, 125.0, 1.0)
    m = drex.Mineral(n_grains=8, seed=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.update_orientations(drex.DefaultParams().as_dict(), np.eye(3), lambda t, x: np.eye(3),
                              (0, 1, lambda t: np.zeros(3)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        drex.misorientation_index(np.eye(3)[None].repeat(4, 0), drex.LatticeSystem.orthorhombic)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under pydrex_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pydrex_b200")):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert "oracle" not in text.lower(), f"{name} mentions the oracle"


def test_api_surface_matches_reference_names():
    import inspect

    import pydrex_b200 as drex
    from pydrex_b200 import core

    sig = inspect.signature(drex.derivatives)
    assert list(sig.parameters) == [
        "regime", "phase", "fabric", "n_grains", "orientations", "fractions", "strain_rate",
        "velocity_gradient", "deformation_gradient_spin", "stress_exponent",
        "deformation_exponent", "nucleation_efficiency", "gbm_mobility", "volume_fraction"]
    sig = inspect.signature(drex.Mineral.update_orientations)
    assert list(sig.parameters)[:6] == ["self", "params", "deformation_gradient",
                                        "get_velocity_gradient", "pathline", "get_regime"]
    assert list(inspect.signature(drex.update_all).parameters)[:6] == [
        "minerals", "params", "deformation_gradient", "get_velocity_gradient", "pathline",
        "get_regime"]
    assert list(inspect.signature(drex.voigt_averages).parameters) == [
        "minerals", "phase_assemblage", "phase_fractions", "elastic_tensors"]
    assert list(inspect.signature(drex.misorientation_index).parameters) == [
        "orientations", "system", "bins"]
    for name in ("_get_slip_invariants", "_get_slip_rates_olivine", "_get_deformation_rate",
                 "_get_slip_rate_softest", "_get_orientation_change", "_get_strain_energy",
                 "_get_rotation_and_strain", "get_crss"):
        assert callable(getattr(core, name))
    assert [int(r) for r in drex.DeformationRegime] == list(range(8))
    nt.assert_array_equal(drex.get_crss(0, 2), [3, 2, np.inf, 1])
    with pytest.raises(ValueError):
        drex.get_crss(0, 5)
    with pytest.raises(ValueError):
        drex.get_crss(1, 0)
    with pytest.raises(ValueError):
        drex.get_crss(2, 0)
    m = drex.Mineral(n_grains=16, seed=8816)
    assert m.orientations[0].shape == (16, 3, 3) and np.allclose(m.fractions[0], 1 / 16)


def test_mineral_npz_roundtrip(tmp_path):
    import pydrex_b200 as drex

    m = drex.Mineral(n_grains=10, seed=3)
    path = str(tmp_path / "m.npz")
    m.save(path)
    m.save(path, postfix="again")
    a, b = drex.Mineral.from_file(path), drex.Mineral.from_file(path, postfix="again")
    for other in (a, b):
        nt.assert_array_equal(other.orientations[0], m.orientations[0])
        nt.assert_array_equal(other.fractions[0], m.fractions[0])
    data = np.load(path)
    assert data["meta"].dtype == np.uint8 and data["orientations"].shape == (1, 10, 3, 3)


def test_velocity_gradient_tabulation():
    from pydrex_b200 import _lib, flow

    pos = lambda t: np.zeros(3)  # noqa: E731
    L0 = np.array([[0.0, 0, 0], [2, 0, 0], [0, 0, 0]])
    kind, nodes = flow.tabulate_velocity_gradient(lambda t, x: L0, pos, 0.0, 1.0)
    assert kind == _lib.L_CONSTANT and nodes.shape == (1, 3, 3)
    kind, nodes = flow.tabulate_velocity_gradient(lambda t, x: L0 * (1 + t), pos, 0.0, 2.0)
    assert kind == _lib.L_LINEAR and nodes.shape == (2, 3, 3)
    nt.assert_allclose(nodes[1], 3 * L0)
    f = lambda t, x: L0 * np.exp(-t) + L0.T * np.sin(3 * t)  # noqa: E731
    kind, nodes = flow.tabulate_velocity_gradient(f, pos, -1.0, 1.0)
    assert kind == _lib.L_CHEBYSHEV and nodes.shape[0] in (9, 17, 33)
    for x in np.linspace(0, 1, 41):
        nt.assert_allclose(flow.barycentric_interpolate(nodes, x), f(-1 + 2 * x, None),
                           rtol=0, atol=1e-9)
    with pytest.raises(ValueError):
        flow.tabulate_velocity_gradient(lambda t, x: np.full((3, 3), np.nan), pos, 0.0, 1.0)


def test_shard_range_partitions_particles():
    from pydrex_b200 import shard_range

    for P in (1, 7, 8, 1000, 10_000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(P, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == P
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from pydrex_b200 import shard_range, gather_to_rank0
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
P = 7
lo, hi = shard_range(P, rank, world)
# each rank "computes" rows that encode the global particle id
local = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.ones(1, 36, dtype=torch.float64)
full = gather_to_rank0(local, P)
if rank == 0:
    assert full.shape == (P, 36), full.shape
    assert torch.equal(full[:, 0], torch.arange(P, dtype=torch.float64))
    print("GATHER_OK")
else:
    assert full is None
dist.destroy_process_group()
"""


def test_world_size_2_gloo_gather(tmp_path):
    """N > 1 path on CPU: two ranks shard 7 particles and gather result rows to rank 0."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29731", str(script), ROOT],
        capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "GATHER_OK" in res.stdout
