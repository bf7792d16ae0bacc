"""Golden vectors for utils.extract_vars / apply_gbs / strain_increment / quat_product and
geometry.misorientation_angles (reference utils.py:144-190, 365-371, geometry.py:75-118),
generated from the unmodified reference; run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refharness  # noqa: E402

_refharness.import_reference()
from pydrex import geometry, utils  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

rng = np.random.default_rng(11)
N = 300
out = {}
y = np.concatenate([rng.normal(size=9), 1.2 * rng.uniform(-1, 1, size=9 * N), rng.normal(0.3, 0.5, size=N)])
F, o, f = utils.extract_vars(y.copy(), N)
out.update(ev_y=y, ev_F=F, ev_o=o, ev_f=f)
o_prev = Rotation.random(N, random_state=1).as_matrix()
o_new = Rotation.random(N, random_state=2).as_matrix()
fr = rng.random(N) ** 4
fr /= fr.sum()
for chi in (0.0, 0.3, 0.9):
    oo, ff = utils.apply_gbs(o_new.copy(), fr.copy(), chi, o_prev.copy(), N)
    out[f"gbs_{int(chi * 10)}_o"], out[f"gbs_{int(chi * 10)}_f"] = oo, ff
out.update(gbs_o_new=o_new, gbs_o_prev=o_prev, gbs_f=fr)
Ls = rng.normal(size=(8, 3, 3))
dts = rng.normal(size=8)
out.update(si_L=Ls, si_dt=dts, si=np.array([utils.strain_increment(dt, L) for dt, L in zip(dts, Ls)]))
q1, q2 = rng.normal(size=4), rng.normal(size=4)
out.update(qp_1=q1, qp_2=q2, qp=np.array(utils.quat_product(q1, q2)))
for tag, dtype in (("f32", np.float32), ("f64", np.float64)):
    a = Rotation.random(40 * 3, random_state=5).as_quat().reshape(40, 3, 4).astype(dtype)
    b = Rotation.random(40 * 5, random_state=6).as_quat().reshape(40, 5, 4).astype(dtype)
    out[f"ma_{tag}_a"], out[f"ma_{tag}_b"] = a, b
    out[f"ma_{tag}"] = geometry.misorientation_angles(a, b)
np.savez_compressed(os.path.join(HERE, "utils.npz"), **out)
print({k: v.shape for k, v in out.items()})
