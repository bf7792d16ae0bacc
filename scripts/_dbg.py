import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pydrex_b200 as drex
dbg = torch.zeros(8 + 8 * 4096, dtype=torch.int64).pin_memory()
os.environ['DREX_DEBUG_HOSTPTR'] = hex(dbg.data_ptr())
from pydrex_b200 import _lib
TWO = {"phase_assemblage": (0, 1), "phase_fractions": (0.7, 0.3), "stress_exponent": 1.5,
       "deformation_exponent": 3.5, "gbm_mobility": 125, "gbs_threshold": 0.3, "nucleation_efficiency": 5.0}
cases = [(3, 333, 4, 0, 4), (2, 31, 3, 0, 4), (5, 1000, 3, 2, 4), (400, 96, 3, 0, 4), (3, 200, 2, 0, 1), (3, 334, 4, 0, 4), (64, 5000, 3, 0, 4)]
P, N, n, kind, regime = cases[int(sys.argv[1])]
e = drex.ParticleEnsemble(P, N, TWO, seed=5, regimes=(regime, regime))
for k in range(n):
    if kind == 0:
        L = torch.zeros((P, 3, 3), dtype=torch.float64)
        L[:, 1, 0] = torch.linspace(1.0, 3.0, P)
    else:
        x = 0.5 * (1.0 - torch.cos(torch.pi * torch.arange(5, dtype=torch.float64) / 4))
        L = torch.zeros((P, 5, 3, 3), dtype=torch.float64)
        L[:, :, 1, 0] = torch.linspace(1.0, 3.0, P)[:, None] * (1.0 + 0.5 * (k + x))[None, :]
        L[:, :, 0, 2] = 0.3 * x[None, :]
    try:
        e.update(0.05 * k, 0.05 * (k + 1), L, kind=kind)
        torch.cuda.synchronize()
    except Exception as ex:
        print("FAILED", str(ex)[:80])
        recs = [r for r in dbg[8:].reshape(-1, 8).tolist() if r[0] != 0]
        print("records", len(recs))
        import collections
        print(collections.Counter(r[0] for r in recs))
        for r in [r for r in recs if r[0] != 20][:48]:
            print("  where", r[0], "cta", r[1], "warp", r[2], "a", hex(r[3] & (2**64-1)), "b", hex(r[4] & (2**64-1)), "c", hex(r[5] & (2**64-1)), "n", r[6])
        sys.exit(1)
    print("case", sys.argv[1], "update", k, "ok", int(e.grain_evals.sum()), flush=True)
