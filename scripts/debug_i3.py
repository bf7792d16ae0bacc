import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import pydrex_b200 as drex
from test_gpu_parity import _cheb_callable, TWO_PHASE
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "integrate.npz"))
N = g["i3_tight_ol_o"].shape[1]
get_L = _cheb_callable(g["i3_default_tnodes"], g["i3_default_Lnodes"])
olA = drex.Mineral(0, 0, 4, n_grains=N, orientations_init=g["i3_default_ol_o"][0].copy(), fractions_init=g["i3_default_ol_f"][0].copy())
ens = drex.Mineral(1, 5, 4, n_grains=N, orientations_init=g["i3_default_en_o"][0].copy(), fractions_init=g["i3_default_en_f"][0].copy())
F = np.eye(3)
ts = g["i3_default_ts"]
from pydrex_b200 import flow
for t0, t1 in zip(ts[:-1], ts[1:]):
    kind, nodes = flow.tabulate_velocity_gradient(get_L, lambda t: np.zeros(3), float(t0), float(t1))
    D = (nodes + np.swapaxes(nodes, -1, -2)) / 2
    s = np.abs(np.linalg.eigvalsh(D)).max(axis=-1)
    print("interval", t0, t1, "kind", kind, "K", len(nodes), "strain", s.mean() * (t1 - t0), "h_next", getattr(olA, "_h_next", None), getattr(ens, "_h_next", None))
    try:
        F = drex.update_all([olA, ens], TWO_PHASE, F, get_L, (t0, t1, lambda t: np.zeros(3)))
        print("   ok  n_rhs", olA.n_rhs, ens.n_rhs, "rej", olA.n_rejected, ens.n_rejected, "finite", np.isfinite(olA.orientations[-1]).all(), np.isfinite(ens.orientations[-1]).all())
    except Exception as e:
        print("   FAIL", e)
        break
