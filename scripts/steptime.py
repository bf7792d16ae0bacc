"""Per-step integrator time and right-hand sides per system at bench size."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, bench
import pydrex_b200 as drex
dev = torch.device('cuda', 0)
P, N = 1250, 5000
kind, tables, (lo, hi) = bench.make_workload('corner_flow', P, 8, seed=1000)
L = [torch.as_tensor(np.ascontiguousarray(t)).to(dev) for t in tables]
lo, hi = torch.as_tensor(lo).to(dev), torch.as_tensor(hi).to(dev)
ens = drex.ParticleEnsemble(P, N, bench.PARAMS, seed=8816, device=dev)
tol = ens.tolerances()
for variant in ('full', 'noorder'):
    for k in range(8):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        if variant == 'full':
            ens.update(lo[k], hi[k], L[k], kind, tol, check=False)
        else:
            order = ens.order
            ens.update(lo[k], hi[k], L[k], kind, tol, check=False)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(variant, k, 'gpu ms %.2f' % e0.elapsed_time(e1), 'cpu ms %.2f' % ((t1 - t0) * 1e3),
              'rhs/grain %.2f' % (float(ens.grain_evals.sum()) / (ens.S * ens.N)),
              'max rhs/grain of a system %.1f' % (float(ens.grain_evals.max()) / ens.N))
