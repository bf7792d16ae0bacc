"""Diagnostic: why is the first timed step after a synchronize slower than the following ones?"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
import pydrex_b200 as drex

P, N = 1250, 5000
dev = torch.device("cuda", 0)
kind, tables, t_edges = bench.make_workload("simple_shear", P, 24, seed=1000)
L_dev = [torch.as_tensor(np.ascontiguousarray(t)).to(dev) for t in tables]
t_lo, t_hi = (torch.as_tensor(np.ascontiguousarray(t)).to(dev) for t in t_edges)
ens = drex.ParticleEnsemble(P, N, bench.PARAMS, seed=8816, device=dev)
tol = ens.tolerances()
k = 0
for mode in ("back-to-back", "sync-before-block", "sleep-20ms-before-block", "dummy-kernel-before-block"):
    for _ in range(3):
        ens.update(t_lo[k], t_hi[k], L_dev[k], kind, tol, check=False); k += 1
    torch.cuda.synchronize()
    if mode == "sleep-20ms-before-block":
        time.sleep(0.02)
    if mode == "dummy-kernel-before-block":
        x = torch.zeros(1 << 26, device=dev); x += 1
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    cpu = []
    ev[0].record()
    for i in range(3):
        t = time.perf_counter()
        ens.update(t_lo[k], t_hi[k], L_dev[k], kind, tol, check=False); k += 1
        cpu.append((time.perf_counter() - t) * 1e3)
        ev[i + 1].record()
    torch.cuda.synchronize()
    print(mode, "gpu ms", [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(3)], "cpu enqueue ms", [round(c, 2) for c in cpu])
