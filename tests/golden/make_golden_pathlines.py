"""Golden vectors for pathlines.get_pathline (reference pathlines.py:10-116) and the analytic
flows of velocity.py:17-97, generated from the unmodified reference; run in the build
container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refharness  # noqa: E402

_refharness.import_reference()
from pydrex import pathlines, velocity  # noqa: E402

out = {}
rng = np.random.default_rng(7)

# --- field samples -------------------------------------------------------------------
fields = {
    "shear": (velocity.simple_shear_2d("Y", "X", 1.5), (0, 1, 0, 1.5, 0.0)),
    "shear_zx": (velocity.simple_shear_2d("Z", "X", 5e-7), (0, 2, 0, 5e-7, 0.0)),
    "cell": (velocity.cell_2d("X", "Z", 6.3e-10, 1e5), (1, 0, 2, 6.3e-10, 1e5)),
    "cell_yx": (velocity.cell_2d("Y", "X", 1.0), (1, 1, 0, 1.0, 2.0)),
    "corner": (velocity.corner_2d("X", "Z", 2.0 / (100 * 365 * 86400)), (2, 0, 2, 2.0 / (100 * 365 * 86400), 0.0)),
    "corner_zy": (velocity.corner_2d("Z", "Y", 1.0), (2, 2, 1, 1.0, 0.0)),
}
for name, ((u, L), desc) in fields.items():
    if name.startswith("cell"):
        half = desc[4] / 2
        x = rng.uniform(-half, half, size=(40, 3))
    elif name.startswith("corner"):
        x = rng.uniform(-2e5, 2e5, size=(40, 3)) if name == "corner" else rng.uniform(-2, 2, size=(40, 3))
        x[0] = 0.0  # the corner itself: NaN
    else:
        x = rng.normal(size=(40, 3))
    out[f"f_{name}_desc"] = np.array(desc, dtype=np.float64)
    out[f"f_{name}_x"] = x
    out[f"f_{name}_v"] = np.array([u(np.nan, xi) for xi in x])
    out[f"f_{name}_L"] = np.array([L(np.nan, xi) for xi in x])

# --- corner-flow pathlines (tests/test_corner_flow_2d.py:107-140) -----------------------
plate_speed = 2.0 / (100.0 * 365.0 * 86400.0)
H, W = 2.0e5, 1.0e6
u, L = velocity.corner_2d("X", "Z", plate_speed)
mn, mx = np.array([0.0, 0.0, -H]), np.array([W, 0.0, 0.0])
zs = np.array([-0.1, -0.3, -0.54, -0.78]) * H
ts, xs = [], []
for z in zs:
    t, pos = pathlines.get_pathline(np.array([W, 0.0, z]), u, L, mn, mx, 10, regular_steps=50)
    ts.append(t)
    xs.append(np.array([pos(ti) for ti in t]))
out["p_corner_desc"] = out["f_corner_desc"]
out["p_corner_final"] = np.array([[W, 0.0, z] for z in zs])
out["p_corner_box"] = np.array([mn, mx])
out["p_corner_max_strain"] = np.array(10.0)
out["p_corner_t"] = np.array(ts)
out["p_corner_x"] = np.array(xs)
# irregular (solver) time stamps
t, pos = pathlines.get_pathline(np.array([W, 0.0, zs[1]]), u, L, mn, mx, 10)
out["p_corner_solver_t"] = t
print("corner t_start (yr):", np.array(ts)[:, 0] / 3.15576e7, "solver steps:", len(t))
print("entry points:", np.array(xs)[:, 0])

# --- convection cell ----------------------------------------------------------------------
u, L = velocity.cell_2d("X", "Z", 1.0)
mn, mx = np.array([-1.0, 0.0, -1.0]), np.array([1.0, 0.0, 1.0])
finals = np.array([[0.5, 0.0, -0.75], [0.2, 0.0, 0.3], [-0.7, 0.0, 0.1]])
ts, xs = [], []
for fl in finals:
    t, pos = pathlines.get_pathline(fl, u, L, mn, mx, 2.0, regular_steps=20)
    ts.append(t)
    xs.append(np.array([pos(ti) for ti in t]))
out["p_cell_desc"] = np.array((1, 0, 2, 1.0, 2.0))
out["p_cell_final"] = finals
out["p_cell_box"] = np.array([mn, mx])
out["p_cell_max_strain"] = np.array(2.0)
out["p_cell_t"] = np.array(ts)
out["p_cell_x"] = np.array(xs)
print("cell t_start:", np.array(ts)[:, 0])

# --- simple shear leaving through the side of the box -----------------------------------
u, L = velocity.simple_shear_2d("X", "Z", 1.0)
mn, mx = np.array([-1.0, 0.0, -1.0]), np.array([1.0, 0.0, 1.0])
finals = np.array([[0.5, 0.0, 0.5], [0.5, 0.0, 0.01], [-0.5, 0.0, -0.8]])
ts, xs = [], []
for fl in finals:
    t, pos = pathlines.get_pathline(fl, u, L, mn, mx, 1.3, regular_steps=10)
    ts.append(t)
    xs.append(np.array([pos(ti) for ti in t]))
out["p_shear_desc"] = np.array((0, 0, 2, 1.0, 0.0))
out["p_shear_final"] = finals
out["p_shear_box"] = np.array([mn, mx])
out["p_shear_max_strain"] = np.array(1.3)
out["p_shear_t"] = np.array(ts)
out["p_shear_x"] = np.array(xs)
print("shear t_start:", np.array(ts)[:, 0], "x_start:", np.array(xs)[:, 0])
np.savez_compressed(os.path.join(HERE, "pathlines.npz"), **out)
