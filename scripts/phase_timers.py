"""Where a CTA of the integrator spends its time: run with a library built with
-DDREX_PHASE_TIMERS (PYDREX_B200_LIB=<that .so>); thread 0's clock at the phase boundaries,
summed over all CTAs, is read back from the counters behind the work queue."""
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
kind, tables, t_edges = bench.make_workload(workload, P, 10, seed=1000)
ens = drex.ParticleEnsemble(P, N, bench.PARAMS, seed=8816, device=dev)
tol = ens.tolerances()
names = ["queue/bookkeeping", "initial slope", "stage flows (serial)", "phase 1 (own chunks)",
         "wait for slowest warp + phase 2", "accept/controller", "end-of-interval pass", "tail"]
for k in range(10):
    ens.update(torch.as_tensor(t_edges[0][k]).to(dev), torch.as_tensor(t_edges[1][k]).to(dev),
               torch.as_tensor(np.ascontiguousarray(tables[k])).to(dev), kind, tol, check=False)
    torch.cuda.synchronize()
    if k >= 6:
        c = ens._ws[64:128].view(torch.int64).cpu().numpy().astype(float)
        print(f"step {k}: " + ", ".join(f"{n} {100 * v / c.sum():.1f}%" for n, v in zip(names, c)))
