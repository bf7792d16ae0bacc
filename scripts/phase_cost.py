"""Launch time of the integrator for olivine-only and enstatite-only ensembles of the bench
size (how the two phases share a two-phase launch)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench
import pydrex_b200 as drex

P, N, K = 2500, 5000, 10
dev = torch.device("cuda", 0)
kind, tables, t_edges = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "simple_shear", P, K, seed=1000)
for name, phases, fabrics in (("olivine", (0,), (0,)), ("enstatite", (1,), (5,))):
    ens = drex.ParticleEnsemble(P, N, bench.PARAMS, phases=phases, fabrics=fabrics, regimes=(4,), seed=8816, device=dev)
    tol = ens.tolerances()
    ms = []
    for k in range(K):
        args = (torch.as_tensor(t_edges[0][k]).to(dev), torch.as_tensor(t_edges[1][k]).to(dev),
                torch.as_tensor(np.ascontiguousarray(tables[k])).to(dev), kind, tol)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ens.update(*args, check=False)
        torch.cuda.synchronize()
        ms.append(1e3 * (time.perf_counter() - t0))
    rhs = float(ens.stats[1].sum()) / P
    print(f"{name}: {np.mean(ms[2:]):.3f} ms per launch of {P} systems, last interval {rhs:.2f} right-hand sides per system, "
          f"{np.mean(ms[2:]) * 148 / P * 1e3 / rhs:.1f} us of a CTA per right-hand side of a system")
