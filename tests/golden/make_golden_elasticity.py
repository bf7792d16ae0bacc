"""Golden vectors for diagnostics.elasticity_components (reference diagnostics.py:38-182),
generated from the unmodified reference; run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refharness  # noqa: E402

_refharness.import_reference()
from pydrex import diagnostics, tensors  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

g = np.load(os.path.join(HERE, "diagnostics.npz"))
c1 = np.load(os.path.join(HERE, "config1_n5000.npz"))
mats = list(g["voigt_C"]) + list(c1["voigt_C"])
# rotated single-crystal olivine: strongly anisotropic, well-conditioned symmetry axes
S_ol = np.array([[320.71, 69.84, 71.22, 0, 0, 0], [69.84, 197.25, 74.8, 0, 0, 0],
                 [71.22, 74.8, 234.32, 0, 0, 0], [0, 0, 0, 63.77, 0, 0],
                 [0, 0, 0, 0, 77.67, 0], [0, 0, 0, 0, 0, 78.36]])
T = tensors.voigt_to_elastic_tensor(S_ol)
for R in Rotation.random(6, random_state=2).as_matrix():
    mats.append(tensors.elastic_tensor_to_voigt(tensors.rotate(T, R)))
mats = np.array(mats)
out = diagnostics.elasticity_components(mats)
np.savez_compressed(os.path.join(HERE, "elasticity.npz"), voigt=mats,
                    **{k: np.asarray(v) for k, v in out.items()})
for k, v in out.items():
    print(k, np.asarray(v)[:3].round(4).tolist())
