"""Launch time of the integrator for the bench workloads with one phase at a time (olivine A
only, enstatite only) and both: how the cost splits between the expensive and the cheap
right-hand side."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench
import pydrex_b200 as drex

P, N = 1250, 5000
workload = sys.argv[1] if len(sys.argv) > 1 else "corner_flow"
dev = torch.device("cuda", 0)
kind, tables, t_edges = bench.make_workload(workload, P, 12, seed=1000)
for name, phases, fabrics in (("olivine", (0,), (0,)), ("enstatite", (1,), (5,)), ("both", (0, 1), (0, 5))):
    ens = drex.ParticleEnsemble(P, N, bench.PARAMS, phases=phases, fabrics=fabrics, regimes=(4,) * len(phases),
                                seed=8816, device=dev)
    tol = ens.tolerances()
    ms, ev = [], []
    for k in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0, t1 = torch.as_tensor(t_edges[0][k]).to(dev), torch.as_tensor(t_edges[1][k]).to(dev)
        L = torch.as_tensor(np.ascontiguousarray(tables[k])).to(dev)
        torch.cuda.synchronize()
        e0.record()
        ens.update(t0, t1, L, kind, tol, check=False)
        e1.record()
        torch.cuda.synchronize()
        if k >= 4:
            ms.append(e0.elapsed_time(e1))
            ev.append(float(ens.grain_evals.sum()) / (ens.S * N))
    print(f"{workload} {name}: {np.mean(ms):.2f} ms per launch, {np.mean(ev):.2f} right-hand sides per grain and interval")
