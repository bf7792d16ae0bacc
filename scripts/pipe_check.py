"""Development check of the pipelined integrator against the one-system kernel: agreement on
small ensembles (odd N, N not a multiple of 32, one and two phases) and timing at bench size."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench
import pydrex_b200 as drex

TWO = dict(bench.PARAMS)
for P, N in ((3, 333), (5, 1000), (2, 31), (7, 5000)):
    ens = [drex.ParticleEnsemble(P, N, TWO, seed=5) for _ in range(2)]
    L = torch.zeros((P, 3, 3), dtype=torch.float64)
    L[:, 1, 0] = torch.linspace(1.0, 3.0, P)
    tols = [e.tolerances() for e in ens]
    tols[0].method, tols[1].method = 0x200, 0x400
    for k in range(4):
        for e, tol in zip(ens, tols):
            e.update(0.05 * k, 0.05 * (k + 1), L, tol=tol)
    a, b = ens
    print(f"P={P} N={N}: max|do| {(a.o - b.o).abs().max().item():.3e} max|df|/f "
          f"{((a.f - b.f).abs() / a.f).max().item():.3e} max|dF| {(a.F - b.F).abs().max().item():.3e} "
          f"stats equal {bool((a.stats == b.stats).all())} rhs {a.stats[1].sum().item()} {b.stats[1].sum().item()}",
          flush=True)

if len(sys.argv) > 1:
    P, N = 1250, 5000
    dev = torch.device("cuda", 0)
    for workload in ("simple_shear", "corner_flow"):
        kind, tables, t_edges = bench.make_workload(workload, P, 10, seed=1000)
        keep = {}
        for name, method in (("single", 0x200), ("pipe", 0x400), ("pipe", 0x400)):
            ens = drex.ParticleEnsemble(P, N, bench.PARAMS, seed=8816, device=dev)
            tol = ens.tolerances()
            tol.method = method
            ms = []
            for k in range(10):
                args = (torch.as_tensor(t_edges[0][k]).to(dev), torch.as_tensor(t_edges[1][k]).to(dev),
                        torch.as_tensor(np.ascontiguousarray(tables[k])).to(dev), kind, tol)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ens.update(*args, check=False)
                torch.cuda.synchronize()
                ms.append(1e3 * (time.perf_counter() - t0))
            print(workload, name, " ".join(f"{m:.2f}" for m in ms), "rhs", ens.stats[1].sum().item(), flush=True)
            if name in keep:
                print("  repeat run bit-identical:", bool((keep[name][0] == ens.o).all()), bool((keep[name][1] == ens.f).all()))
            elif keep:
                ref = keep["single"]
                print(f"  vs single: max|do| {(ref[0] - ens.o).abs().max().item():.3e} max|df|/f {((ref[1] - ens.f).abs() / ref[1]).max().item():.3e}")
            keep[name] = (ens.o.clone(), ens.f.clone())
            if os.environ.get("DREX_TIMERS"):
                c = ens._ws[64:192].view(torch.int64).cpu().numpy().astype(float)
                print("  timers", c)
