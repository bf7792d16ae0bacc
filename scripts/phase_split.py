"""Time the integrator on olivine-only and enstatite-only ensembles (same flow) to see how a
two-phase launch splits."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, bench
import pydrex_b200 as drex
dev = torch.device('cuda', 0)
P, N = 1250, 5000
kind, tables, (lo, hi) = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else 'corner_flow', P, 8, seed=1000)
L = [torch.as_tensor(np.ascontiguousarray(t)).to(dev) for t in tables]
lo, hi = torch.as_tensor(lo).to(dev), torch.as_tensor(hi).to(dev)
for name, ph, fab in (('olivine', (0,), (0,)), ('enstatite', (1,), (5,)), ('both', (0, 1), (0, 5))):
    ens = drex.ParticleEnsemble(P, N, bench.PARAMS, phases=ph, fabrics=fab, regimes=(4,) * len(ph), seed=8816, device=dev)
    tol = ens.tolerances()
    tot = 0.0; rhs = 0
    for k in range(8):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ens.update(lo[k], hi[k], L[k], kind, tol, check=False); e1.record()
        torch.cuda.synchronize()
        if k >= 3:
            tot += e0.elapsed_time(e1); rhs += int(ens.stats[1].sum())
    print(name, 'ms/step %.3f' % (tot / 5), 'evals/s %.3e' % (rhs * N / tot * 1e3), 'rhs/system/step %.2f' % (rhs / 5 / ens.S))
