"""Diagnostic for the host-buffer arm: copy-out bandwidth alone, and update_host with different
group counts, at bench size."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench
import pydrex_b200 as drex

P, N = 1250, 5000
dev = torch.device("cuda", 0)
workload = sys.argv[1] if len(sys.argv) > 1 else "corner_flow"
kind, tables, t_edges = bench.make_workload(workload, P, 14, seed=1000)
ens = drex.ParticleEnsemble(P, N, bench.PARAMS, seed=1, device=dev)
S = ens.S
h_o = ens.orientations().reshape(S, N, 9).cpu().pin_memory()
h_f = ens.f.cpu().pin_memory()
h_F = ens.F.cpu().pin_memory()
h_L = [torch.as_tensor(np.ascontiguousarray(t)).pin_memory() for t in tables]
h_lo, h_hi = (torch.as_tensor(np.ascontiguousarray(t)).pin_memory() for t in t_edges)
d = torch.empty((S, N, 9), dtype=torch.float64, device=dev)
for _ in range(2):
    h_o.copy_(d, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    h_o.copy_(d, non_blocking=True)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 5 * 1e3
print(f"D2H 0.9 GB in one copy: {ms:.2f} ms = {d.numel() * 8 / ms / 1e6:.1f} GB/s")
# the same bytes as update_host moves out, cut the same way (8 groups x 3 copies), copies only
s_out = torch.cuda.Stream(device=dev)
per = -(-P // 8) * 2
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s_out):
        for c in range(8):
            s0, s1 = c * per, min(S, (c + 1) * per)
            h_o[s0:s1].copy_(d[s0:s1], non_blocking=True)
            h_f[s0:s1].copy_(ens.f[s0:s1], non_blocking=True)
            h_F[s0:s1].copy_(ens.F[s0:s1], non_blocking=True)
    s_out.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
print(f"D2H of the whole state in 24 copies on a side stream: {ms:.2f} ms")
for n_chunks in (2, 3, 4, 8):
    for k in range(6, 8):
        ens.update_host(h_o, h_f, h_F, h_lo[k], h_hi[k], h_L[k], kind, n_chunks=n_chunks)
    t0 = time.perf_counter()
    for k in range(8, 14):
        ens.update_host(h_o, h_f, h_F, h_lo[k], h_hi[k], h_L[k], kind, n_chunks=n_chunks)
    ms = (time.perf_counter() - t0) / 6 * 1e3
    print(f"update_host, {n_chunks} groups, no state upload: {ms:.2f} ms per step")
ens.mark_host_dirty()
for k in range(2):
    ens.update_host(h_o, h_f, h_F, h_lo[k], h_hi[k], h_L[k], kind, n_chunks=8, upload=True)
t0 = time.perf_counter()
for k in range(2, 8):
    ens.update_host(h_o, h_f, h_F, h_lo[k], h_hi[k], h_L[k], kind, n_chunks=8, upload=True)
print(f"update_host, 8 groups, state uploaded every step: {(time.perf_counter() - t0) / 6 * 1e3:.2f} ms per step")
