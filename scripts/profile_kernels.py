"""One pass over every kernel of the path at bench size, as the target command for ncu
(scripts/profile_all.sh): a few integrator launches on the corner-flow workload, then the
standalone derivative kernels, the Voigt average, the M-index, the pathline tracer, the
resampler and the FP64 peak probe."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench
import pydrex_b200 as drex
from pydrex_b200 import _lib

P, N = 1250, 5000
n_updates = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
kind, tables, t_edges = bench.make_workload("corner_flow", P, n_updates, seed=1000)
ens = drex.ParticleEnsemble(P, N, bench.PARAMS, seed=8816, device=dev)
for k in range(n_updates):
    ens.update(torch.as_tensor(t_edges[0][k]).to(dev), torch.as_tensor(t_edges[1][k]).to(dev),
               torch.as_tensor(np.ascontiguousarray(tables[k])).to(dev), kind, check=False)
torch.cuda.synchronize()
print("integrator evals of the last launch:", int(ens.grain_evals.sum()))
lib, st, Pp = _lib.load(), _lib.stream_ptr(torch), _lib.ptr
S = ens.S
D = torch.zeros((S, 9), dtype=torch.float64, device=dev)
D[:, 1] = D[:, 3] = 1.0
Ls = torch.zeros((S, 9), dtype=torch.float64, device=dev)
Ls[:, 3] = 2.0
spin = torch.zeros((S, 9), dtype=torch.float64, device=dev)
do, df = torch.empty_like(ens.o), torch.empty_like(ens.f)
prm = _lib.DrexParams(1.5, 3.5, 5.0, 125.0, 0.3)
nws = lib.drex_derivatives_workspace_bytes(S, N)
ws = torch.empty(nws, dtype=torch.uint8, device=dev)
m, mh = ens.meta, ens.meta_host
for _ in range(2):
    _lib.check(lib.drex_derivatives_f64(
        Pp(ens.o), Pp(ens.f), Pp(D), Pp(Ls), Pp(spin), Pp(m[3]), Pp(m[1]), Pp(m[2]), Pp(ens.vol_frac),
        Pp(mh[3]), Pp(mh[1]), Pp(mh[2]), ctypes.byref(prm), S, N, Pp(do), Pp(df), Pp(ws), nws, st))
    ens.voigt_averages()
sub = drex.ParticleEnsemble(64, N, bench.PARAMS, seed=2, device=dev)
sub.o.copy_(ens.o[: sub.S])
sub.f.copy_(ens.f[: sub.S])
for _ in range(2):
    sub.misorientation_indices(drex.LatticeSystem.orthorhombic, phase_slot=0)
    sub.resample(1000, list(range(64)), 0)
bench.trace_pathlines_on_device(drex, torch, P)
flops = ctypes.c_double()
_lib.check(lib.drex_fp64_peak_probe(0, ctypes.byref(flops)))
torch.cuda.synchronize()
print("done; fp64 probe", flops.value / 1e12, "TFLOP/s")
