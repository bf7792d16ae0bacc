"""Where the warps of the integrator spend their time: run with a library built with
-DDREX_PHASE_TIMERS (PYDREX_B200_LIB=<that .so>); every warp's clock at the boundaries of its
activities, summed over the launch, is read back from the counters behind the work queue."""
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
which = sys.argv[2] if len(sys.argv) > 2 else "both"
phases, fabrics = {"both": ((0, 1), (0, 5)), "olivine": ((0,), (0,)), "enstatite": ((1,), (5,))}[which]
ens = drex.ParticleEnsemble(P, N, bench.PARAMS, phases=phases, fabrics=fabrics, regimes=(4,) * len(phases),
                            seed=8816, device=dev)
tol = ens.tolerances()
names = ["chunk tickets / waiting for a system", "preparing systems", "waiting for prefetched rows",
         "chunk stages", "chunk results", "finishing systems", "stage flows", "error norm + controller"]
for k in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ens.update(torch.as_tensor(t_edges[0][k]).to(dev), torch.as_tensor(t_edges[1][k]).to(dev),
               torch.as_tensor(np.ascontiguousarray(tables[k])).to(dev), kind, tol, check=False)
    e1.record()
    torch.cuda.synchronize()
    if k >= 6:
        c = ens._ws[64:128].view(torch.int64).cpu().numpy().astype(float)
        ev = float(ens.grain_evals.sum())
        ms = e0.elapsed_time(e1)
        warp_cycles = ms * 1e-3 * 1.965e9 * 148 * 16
        print(f"step {k}: {ms:.2f} ms, {ev / 32 / 1e6:.2f} M warp-evals, "
              f"{c[3] / (ev / 32):.0f} cycles per warp-eval in the stages; accounted "
              f"{100 * c.sum() / warp_cycles:.0f}% of warp-time: "
              + ", ".join(f"{n} {100 * v / warp_cycles:.1f}%" for n, v in zip(names, c) if n != "-"))
