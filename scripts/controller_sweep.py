"""Right-hand sides, rejected steps and launch time of the integrator over the first outer
intervals of both bench workloads; run once per library variant (PYDREX_B200_LIB) to compare
step-controller settings."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench
import pydrex_b200 as drex

P, N, K = 1250, 5000, int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device("cuda", 0)
for workload in ("simple_shear", "corner_flow"):
    kind, tables, t_edges = bench.make_workload(workload, P, K, seed=1000)
    ens = drex.ParticleEnsemble(P, N, bench.PARAMS, seed=8816, device=dev)
    tol = ens.tolerances()
    ms, rhs, rej = [], 0, 0
    for k in range(K):
        args = (torch.as_tensor(t_edges[0][k]).to(dev), torch.as_tensor(t_edges[1][k]).to(dev),
                torch.as_tensor(np.ascontiguousarray(tables[k])).to(dev), kind, tol)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ens.update(*args, check=False)
        torch.cuda.synchronize()
        ms.append(1e3 * (time.perf_counter() - t0))
        rhs += float(ens.grain_evals.sum()) / N
        rej += int(ens.stats[3].sum())
    f = ens.f.cpu().numpy()
    print(f"{os.environ.get('PYDREX_B200_LIB', 'default')[-14:]} {workload}: {np.mean(ms[2:]):.3f} ms/step, "
          f"rhs/grain/step {rhs / (2 * P * K):.3f}, rejected {rej}, failed {int((ens.stats[0] != 0).sum())}, "
          f"checksum {float(ens.o.sum()):.10f} {float(f.max()):.8e}", flush=True)
