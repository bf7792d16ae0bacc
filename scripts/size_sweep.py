"""Integrator throughput against the number of grains per system (simple shear, two phases,
about the same total number of grains per launch)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench
import pydrex_b200 as drex

dev = torch.device("cuda", 0)
for N in [int(x) for x in os.environ.get("SWEEP_N", "100,500,1000,2000,5000,10000,20000,50000").split(",")]:
    P = max(148, int(6.25e6 / N))
    kind, tables, t_edges = bench.make_workload("simple_shear", P, 9, seed=1000)
    ens = drex.ParticleEnsemble(P, N, bench.PARAMS, seed=8816, device=dev)
    tol = ens.tolerances()
    ms, evals = [], 0
    for k in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ens.update(torch.as_tensor(t_edges[0][k]).to(dev), torch.as_tensor(t_edges[1][k]).to(dev),
                   torch.as_tensor(np.ascontiguousarray(tables[k])).to(dev), kind, tol, check=False)
        e1.record()
        torch.cuda.synchronize()
        if k >= 4:
            ms.append(e0.elapsed_time(e1))
            evals += int(ens.grain_evals.sum())
    bad = int((ens.stats[0] != 0).sum())
    print(f"N={N:6d} P={P:6d}: {np.mean(ms):6.2f} ms/step, {evals / (sum(ms) * 1e-3):.3e} grain-evals/s, "
          f"{P * 2 * N * len(ms) / (sum(ms) * 1e-3):.3e} grain-steps/s, failed {bad}")
    del ens
    torch.cuda.empty_cache()
