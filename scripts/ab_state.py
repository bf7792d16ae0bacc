"""State after eight outer intervals of both bench workloads (300 particles x 2 phases x 5000 grains) as
SHA-256 digests: run with two builds of the library (PYDREX_B200_LIB) to check that a change of the
integrator that should not change results does not -- e.g. the flow cache and the step counters of
round 2 against the kernel before them: every digest identical."""
import hashlib
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench, pydrex_b200 as drex
P, N = 300, 5000
dev = torch.device("cuda", 0)
out = {}
for workload in ("corner_flow", "simple_shear"):
    kind, tables, t_edges = bench.make_workload(workload, P, 8, seed=1000)
    ens = drex.ParticleEnsemble(P, N, bench.PARAMS, seed=8816, device=dev)
    tol = ens.tolerances()
    for k in range(8):
        ens.update(torch.as_tensor(t_edges[0][k]).to(dev), torch.as_tensor(t_edges[1][k]).to(dev),
                   torch.as_tensor(np.ascontiguousarray(tables[k])).to(dev), kind, tol)
    out[workload + "_o"] = ens.o.cpu().numpy()
    out[workload + "_f"] = ens.f.cpu().numpy()
    out[workload + "_F"] = ens.F.cpu().numpy()
    out[workload + "_evals"] = ens.grain_evals.cpu().numpy()
for k in sorted(out):
    print(k, hashlib.sha256(np.ascontiguousarray(out[k]).tobytes()).hexdigest())
