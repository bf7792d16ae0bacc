"""Golden vectors for stats.resample_orientations (reference stats.py:14-64), generated from
the unmodified reference; run in the build container only.

Case "distinct": fractions without ties, so the reference's (unstable) argsort has one
answer.  Case "evolved": a texture integrated with grain boundary sliding, where many grains
sit on the same floor fraction: there only the returned FRACTIONS are defined independently
of the numpy build (tied grains are interchangeable)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refharness  # noqa: E402

_refharness.import_reference()
from pydrex import stats  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

rng = np.random.default_rng(1234)
T, N = 3, 700
o = Rotation.random(T * N, random_state=77).as_matrix().reshape(T, N, 3, 3)
f = rng.random((T, N)) ** 3
f /= f.sum(axis=1, keepdims=True)
assert all(len(np.unique(row)) == N for row in f)
out = {"d_orientations": o, "d_fractions": f}
for tag, n_samples, seed in (("default", None, 8816), ("n2000", 2000, 1), ("n10", 10, 0)):
    ro, rf = stats.resample_orientations(o, f, n_samples=n_samples, seed=seed)
    out[f"d_{tag}_o"], out[f"d_{tag}_f"] = ro, rf
    out[f"d_{tag}_args"] = np.array([-1 if n_samples is None else n_samples, seed])

g = np.load(os.path.join(HERE, "integrate.npz"))
eo, ef = g["i1_default_o"], g["i1_default_f"]  # 6 snapshots x 400 grains, GBS floor reached
print("evolved: unique fractions per snapshot", [len(np.unique(r)) for r in ef], "of", ef.shape[1])
ro, rf = stats.resample_orientations(eo, ef, n_samples=1000, seed=8816)
out["e_o"], out["e_f"] = ro, rf
np.savez_compressed(os.path.join(HERE, "resample.npz"), **out)
print({k: v.shape for k, v in out.items()})
