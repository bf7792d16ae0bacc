"""Golden vectors for the tensors.py helpers (reference tensors.py:14-21, 39-273), generated from the unmodified reference; run in the
build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refharness  # noqa: E402

_refharness.import_reference()
from pydrex import tensors  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

rng = np.random.default_rng(3)
B = 6
T = rng.normal(size=(B, 3, 3, 3, 3))
R = Rotation.random(B, random_state=4).as_matrix()
V = rng.normal(size=(B, 6, 6))
V = V + V.transpose(0, 2, 1)
M = rng.normal(size=(B, 3, 3)) + 2 * np.eye(3)
out = {
    "T": T, "R": R, "V": V, "M": M,
    "rotated": np.array([tensors.rotate(t, r) for t, r in zip(T, R)]),
    "v2t": np.array([tensors.voigt_to_elastic_tensor(v) for v in V]),
    "t2v": np.array([tensors.elastic_tensor_to_voigt(t) for t in T]),
    "t2v_roundtrip": np.array([tensors.elastic_tensor_to_voigt(tensors.voigt_to_elastic_tensor(v)) for v in V]),
}
# decomposition helpers (tensors.py:39-164, 206-251), incl. the matrices the reference's own
# tests use (tests/test_tensors.py:16-126) and a matrix with exact zeros in the upper triangle
X = rng.normal(size=(B, 21))
U = rng.normal(size=(B, 6, 6))
U[1, 0, 3] = 0.0
U3 = rng.normal(size=(B, 3, 3))
U3[2, 0, 1] = 0.0
dec = [tensors.voigt_decompose(v) for v in V]
out.update({
    "X": X, "U": U, "U3": U3,
    "dilat": np.array([d[0] for d in dec]), "deviat": np.array([d[1] for d in dec]),
    "mono": np.array([tensors.mono_project(x) for x in X]),
    "ortho": np.array([tensors.ortho_project(x) for x in X]),
    "tetr": np.array([tensors.tetr_project(x) for x in X]),
    "hex": np.array([tensors.hex_project(x) for x in X]),
    "m2v": np.array([tensors.voigt_matrix_to_vector(v) for v in V]),
    "v2m": np.array([tensors.voigt_vector_to_matrix(x) for x in X]),
    "sym6": np.array([tensors.upper_tri_to_symmetric(u) for u in U]),
    "sym3": np.array([tensors.upper_tri_to_symmetric(u) for u in U3]),
})
for left in (True, False):
    pd = [tensors.polar_decompose(m, left=left) for m in M]
    out[f"polar_rot_{int(left)}"] = np.array([p[0] for p in pd])
    out[f"polar_stretch_{int(left)}"] = np.array([p[1] for p in pd])
np.savez_compressed(os.path.join(HERE, "tensors.npz"), **out)
print({k: v.shape for k, v in out.items()})
