"""Golden vectors for the tensors.py helpers on the Voigt-average path (reference
tensors.py:14-21, 167-203, 254-273), generated from the unmodified reference; run in the
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
for left in (True, False):
    pd = [tensors.polar_decompose(m, left=left) for m in M]
    out[f"polar_rot_{int(left)}"] = np.array([p[0] for p in pd])
    out[f"polar_stretch_{int(left)}"] = np.array([p[1] for p in pd])
np.savez_compressed(os.path.join(HERE, "tensors.npz"), **out)
print({k: v.shape for k, v in out.items()})
