"""pytest configuration: the `gpu` marker and shared fixtures."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_derivatives():
    return _load("derivatives.npz")


@pytest.fixture(scope="session")
def golden_integrate():
    return _load("integrate.npz")


@pytest.fixture(scope="session")
def golden_diagnostics():
    return _load("diagnostics.npz")


@pytest.fixture(scope="session")
def golden_resample():
    return _load("resample.npz")


@pytest.fixture(scope="session")
def golden_config1():
    return _load("config1_n5000.npz")


@pytest.fixture(scope="session")
def golden_n5000():
    return _load("n5000.npz")


def cheb_interpolant(tnodes, Lnodes):
    """L(t) along a stored pathline: barycentric interpolation through the Chebyshev-Lobatto
    samples of each outer interval (tnodes [n_int, K], Lnodes [n_int, K, 3, 3]); this is how
    the fixtures carry the reference's get_velocity_gradient(t, get_position(t))."""
    tnodes, Lnodes = np.asarray(tnodes), np.asarray(Lnodes)
    K = tnodes.shape[1]
    w = np.where(np.arange(K) % 2 == 0, 1.0, -1.0)
    w[0] *= 0.5
    w[-1] *= 0.5

    def get_L(t, x=None):
        lo, hi = tnodes[:, 0], tnodes[:, -1]
        i = int(np.clip(np.searchsorted(hi, t, side="left"), 0, len(hi) - 1))
        if abs(hi[i] - lo[i]) == 0:
            return Lnodes[i, 0]
        d = t - tnodes[i]
        hit = np.abs(d) < 1e-15 * max(1.0, abs(hi[i] - lo[i]))
        if hit.any():
            return Lnodes[i, int(np.argmax(hit))]
        c = w / d
        return np.tensordot(c, Lnodes[i], axes=(0, 0)) / c.sum()

    return get_L


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as _oracle

    _oracle.build()
    return _oracle


@pytest.fixture(scope="session")
def golden_pathlines():
    return _load("pathlines.npz")
