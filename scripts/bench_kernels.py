"""Micro-benchmarks of the non-integrator kernels at config sizes (config 5 of BASELINE.json):
standalone core.derivatives, Voigt average, M-index.  CUDA-event timings, inputs larger than L2
where the kernel streams (derivatives, Voigt); prints one JSON line per kernel."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pydrex_b200 as drex  # noqa: E402
from pydrex_b200 import _lib  # noqa: E402

PARAMS = {"phase_assemblage": (0, 1), "phase_fractions": (0.7, 0.3), "stress_exponent": 1.5,
          "deformation_exponent": 3.5, "gbm_mobility": 125, "gbs_threshold": 0.3,
          "nucleation_efficiency": 5.0}


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    n_mindex = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    ens = drex.ParticleEnsemble(P, N, PARAMS, seed=1, device=dev)
    L = torch.zeros((P, 3, 3), dtype=torch.float64)
    L[:, 1, 0] = 2.0
    for k in range(3):  # evolve a little so the textures are not uniform
        ens.update(0.1 * k, 0.1 * (k + 1), L)
    S = ens.S
    st = _lib.stream_ptr(torch)
    Pp = _lib.ptr

    # --- standalone derivatives (B_alg = 160 B per grain-evaluation)
    D = torch.zeros((S, 9), dtype=torch.float64, device=dev)
    D[:, 1] = D[:, 3] = 1.0
    Ls = torch.zeros((S, 9), dtype=torch.float64, device=dev)
    Ls[:, 3] = 2.0
    spin = torch.zeros((S, 9), dtype=torch.float64, device=dev)
    do = torch.empty_like(ens.o)
    df = torch.empty_like(ens.f)
    prm = _lib.DrexParams(1.5, 3.5, 5.0, 125.0, 0.3)
    nws = lib.drex_derivatives_workspace_bytes(S, N)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    m, mh = ens.meta, ens.meta_host

    def deriv():
        _lib.check(lib.drex_derivatives_f64(
            Pp(ens.o), Pp(ens.f), Pp(D), Pp(Ls), Pp(spin), Pp(m[3]), Pp(m[1]), Pp(m[2]),
            Pp(ens.vol_frac), Pp(mh[3]), Pp(mh[1]), Pp(mh[2]), ctypes.byref(prm), S, N, Pp(do), Pp(df),
            Pp(ws), nws, st))

    ms = timed(deriv)
    evals = S * N
    print(json.dumps({"kernel": "derivatives (grain + gbm kernels)", "ms": ms, "grain_evals_per_s": evals / ms * 1e3,
                      "roofline": {"bound": "hbm", "achieved": 160.0 * evals / ms * 1e-6, "peak": hbm,
                                   "unit": "GB/s", "frac": 160.0 * evals / ms * 1e-6 / hbm,
                                   "algorithmic_bytes_per_eval": 160}}))

    # --- Voigt average (80 B per grain)
    ms = timed(lambda: ens.voigt_averages())
    print(json.dumps({"kernel": "voigt_average", "ms": ms, "grains_per_s": S * N / ms * 1e3,
                      "roofline": {"bound": "hbm", "achieved": 80.0 * S * N / ms * 1e-6, "peak": hbm,
                                   "unit": "GB/s", "frac": 80.0 * S * N / ms * 1e-6 / hbm}}))

    # --- M-index, orthorhombic, N grains per texture
    sub = drex.ParticleEnsemble(n_mindex, N, PARAMS, seed=2, device=dev)
    sub.o.copy_(ens.o[: sub.S])
    system = drex.LatticeSystem.orthorhombic
    ms = timed(lambda: sub.misorientation_indices(system, phase_slot=0), reps=3, warm=1)
    pairs = n_mindex * N * (N - 1) / 2
    flop = pairs * 343.0
    print(json.dumps({"kernel": "mindex (quats + pairs + finalize)", "ms": ms, "textures": n_mindex,
                      "textures_per_s": n_mindex / ms * 1e3, "pairs_per_s": pairs / ms * 1e3,
                      "fp32_tflops_algorithmic": flop / ms * 1e-9}))


if __name__ == "__main__":
    main()
