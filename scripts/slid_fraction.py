"""How many grains (and 32-grain chunks) sit on the grain-boundary-sliding floor after each
outer step of the bench workload?  (Decides whether carrying the FSAL slope across outer
intervals could pay: a chunk with a slid grain has to be re-evaluated anyway.)"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench
import pydrex_b200 as drex

P, N = 296, 5000
dev = torch.device("cuda", 0)
kind, tables, t_edges = bench.make_workload("corner_flow", P, 36, seed=1000)
ens = drex.ParticleEnsemble(P, N, bench.PARAMS, seed=8816, device=dev)
tol = ens.tolerances()
for k in range(36):
    ens.update(torch.as_tensor(t_edges[0][k]).to(dev), torch.as_tensor(t_edges[1][k]).to(dev),
               torch.as_tensor(np.ascontiguousarray(tables[k])).to(dev), kind, tol, check=False)
    f = ens.fractions()[:, 0]  # olivine
    floor = f.min(dim=1, keepdim=True).values
    slid = f <= floor * (1 + 1e-12)
    pad = (-N) % 32
    chunks = torch.nn.functional.pad(slid, (0, pad)).reshape(P, -1, 32).any(dim=2)
    if k % 3 == 2:
        print(f"step {k + 1:2d}: grains on the floor {100 * slid.float().mean():5.1f} %, "
              f"chunks with one {100 * chunks.float().mean():5.1f} %")
