#!/usr/bin/env python
"""bench.py -- throughput of the CPO-integration hot path on B200 (and the CPU baseline).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the reference itself
                                                                      # (baseline/_ref), else the port

Workload (BASELINE.json configs[3]'s per-GPU share, weak scaling): each GPU holds
`--particles` pathline particles (default 1250 = 10^4 / 8), each carrying an olivine A-type
and an enstatite ensemble of `--grains` (5000) grains, GBM and GBS enabled (M* = 125,
chi = 0.3), advanced through outer CPO intervals along synthetic pathlines:
  simple_shear  every particle sees a randomly oriented simple shear with its own strain
                rate, strain increment 0.02-0.08 per outer step;
  corner_flow   particles ride 2-D corner-flow pathlines (analytic Batchelor solution,
                plate speed 2 cm/yr, exit depths spread over the domain); L(t) is
                tabulated per outer interval at 9 Chebyshev nodes.

A "step" is one outer CPO interval of every particle: ONE launch of the persistent
integrator kernel.  metric "grain-steps/s" = grains x outer intervals advanced per second
(the integrator-independent unit of useful work); grain-derivative evaluations/s and
internal integrator steps/s are reported next to it and feed the roofline.

value  : inputs resident in HBM, device-timed with CUDA events (max over ranks)
e2e    : the same through the host-buffer API -- every step copies that step's
         velocity-gradient tables from pinned host memory, integrates, converts layout on
         device and copies the new state (orientations + fractions + F, reference AoS layout)
         and the per-system counters back; the state itself is uploaded once.
The line also carries the other workload (`workloads`), the rates of the other kernels of the
path (`kernels`), and a parity check of the timed configuration against the CPU arm on the same
textures (`parity_check`).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_ALG = 254.0  # FP64 flop per grain-derivative evaluation
B_ALG_DERIV = 160.0  # bytes per evaluation of the standalone derivative kernel
PARAMS = {"phase_assemblage": (0, 1), "phase_fractions": (0.7, 0.3), "stress_exponent": 1.5,
          "deformation_exponent": 3.5, "gbm_mobility": 125, "gbs_threshold": 0.3,
          "nucleation_efficiency": 5.0}


# ----------------------------------------------------------------------------- workloads


def simple_shear_tables(P, n_steps, seed):
    """Per-particle constant velocity gradients: a simple shear of rate r_p in a random
    frame.  Returns (kind, L[n_steps][P,1,3,3], (t_lo[n_steps][P], t_hi[n_steps][P]))."""
    rng = np.random.default_rng(seed)
    from scipy.spatial.transform import Rotation

    R = Rotation.random(P, random_state=seed).as_matrix()
    rate = rng.uniform(0.5, 2.0, size=P)
    L0 = np.zeros((3, 3))
    L0[1, 0] = 2.0
    L = np.einsum("pij,jk,plk->pil", R, L0, R) * rate[:, None, None]
    dt = 0.04
    t_edges = np.arange(n_steps + 1)[:, None] * dt * np.ones(P)[None]
    return 0, [L[:, None].copy() for _ in range(n_steps)], (t_edges[:-1], t_edges[1:])


def corner_flow_tables(P, n_steps, seed, K=9, n_path_steps=50):
    """2-D corner-flow pathlines (config 3/4): particles exit at x = 1e6 m, depths spread
    over 5-95 % of the 200 km box, plate speed 2 cm/yr (geometry of the reference's
    tests/test_corner_flow_2d.py:107-113; velocity field = the analytic Batchelor corner
    solution of velocity._corner_2d*).  Each pathline is traced backwards with a vectorised
    RK4 (host, numpy) until it leaves the box (strain target 10 is never reached first),
    cut into 50 equal-time outer intervals like `get_pathline(..., regular_steps=50)`
    (pathlines.py:113-116), and L(t) = grad v(x(t)) is tabulated at K Chebyshev-Lobatto
    nodes per interval.  Strain per outer interval ranges from 0.004 to 0.5: strongly
    different work per particle.  Bench step k uses interval k mod 50."""
    U = 2.0 / (100 * 365 * 86400)
    H, W = 2e5, 1e6
    z_exit = -np.linspace(0.05, 0.95, P) * H

    def vel(x, z):
        r2 = x * x + z * z
        pre = 2 * U / np.pi
        return pre * (np.arctan2(x, -z) + x * z / r2), pre * z * z / r2

    def grad(x, z):
        pre = 4 * U / (np.pi * (x * x + z * z) ** 2)
        L = np.zeros(x.shape + (3, 3))
        L[..., 0, 0] = -pre * x * x * z
        L[..., 0, 2] = pre * x**3
        L[..., 2, 0] = -pre * x * z * z
        L[..., 2, 2] = pre * x * x * z
        return L

    def strain_rate(x, z):
        pre = 4 * U / (np.pi * (x * x + z * z) ** 2)
        return np.hypot(pre * x * x * z, 0.5 * pre * (x**3 - x * z * z))

    n_fine, d_eps = 4000, 10.0 / 4000
    x, z, t = np.full(P, W), z_exit.copy(), np.zeros(P)
    alive = np.ones(P, dtype=bool)
    T, X, Z = [t.copy()], [x.copy()], [z.copy()]
    for _ in range(n_fine):
        vx, vz = vel(x, z)
        dt = -np.minimum(d_eps / strain_rate(x, z), 2e3 / np.hypot(vx, vz))
        k1 = (vx, vz)
        k2 = vel(x + 0.5 * dt * k1[0], z + 0.5 * dt * k1[1])
        k3 = vel(x + 0.5 * dt * k2[0], z + 0.5 * dt * k2[1])
        k4 = vel(x + dt * k3[0], z + dt * k3[1])
        xn = x + dt / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
        zn = z + dt / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
        alive &= (xn > 0) & (zn > -H) & (zn < 0)
        x, z, t = np.where(alive, xn, x), np.where(alive, zn, z), np.where(alive, t + dt, t)
        T.append(t.copy())
        X.append(x.copy())
        Z.append(z.copy())
    T, X, Z = np.array(T)[::-1], np.array(X)[::-1], np.array(Z)[::-1]  # forward in time
    cheb = 0.5 * (1 - np.cos(np.pi * np.arange(K) / (K - 1)))
    edges = T[0][None] + np.linspace(0.0, 1.0, n_path_steps + 1)[:, None] * (0.0 - T[0])[None]
    t_nodes = edges[:-1, None, :] + cheb[None, :, None] * np.diff(edges, axis=0)[:, None, :]
    L_path = np.empty((n_path_steps, P, K, 3, 3))
    for p in range(P):
        keep = np.concatenate([[True], np.diff(T[:, p]) > 0])  # frozen tail after leaving
        xp = np.interp(t_nodes[:, :, p], T[keep, p], X[keep, p])
        zp = np.interp(t_nodes[:, :, p], T[keep, p], Z[keep, p])
        L_path[:, p] = grad(xp, zp)
    tables = [L_path[k % n_path_steps] for k in range(n_steps)]
    t_lo = np.stack([edges[k % n_path_steps] for k in range(n_steps)])
    t_hi = np.stack([edges[k % n_path_steps + 1] for k in range(n_steps)])
    return 2, tables, (t_lo, t_hi)



def trace_pathlines_on_device(drex, torch, P, K=9, n_path_steps=50):
    """the same particle cloud traced by the library's device tracer
    (drex_pathlines_f64: pathlines + the L(t) tables of all 50 intervals in one launch).
    Reported beside the bench, not inside the timed region: the timed workload keeps the
    host-generated tables so that the GPU arm and the CPU arm integrate identical inputs."""
    U, H, W = 2.0 / (100 * 365 * 86400), 2e5, 1e6
    finals = np.zeros((P, 3))
    finals[:, 0], finals[:, 2] = W, -np.linspace(0.05, 0.95, P) * H
    u, L = drex.velocity.corner_2d("X", "Z", U)
    x = torch.as_tensor(finals).cuda()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = []
    for _ in range(4):
        ev[0].record()
        paths = drex.pathlines.get_pathlines(x, u, L, [0, 0, -H], [W, 0, 0], 10.0, n_path_steps,
                                             n_nodes=K, check=False)
        ev[1].record()
        torch.cuda.synchronize()
        ms.append(ev[0].elapsed_time(ev[1]))
    return {"particles": P, "intervals": n_path_steps, "nodes_per_interval": K,
            "ms_per_trace": min(ms[1:]), "failed": int((paths.status != 0).sum()),
            "what": "drex_pathlines_f64: backward Dormand-Prince trace + L tables, one launch"}


def make_workload(name, P, n_steps, seed):
    if name == "simple_shear":
        return simple_shear_tables(P, n_steps, seed)
    if name == "corner_flow":
        return corner_flow_tables(P, n_steps, seed)
    raise SystemExit(f"unknown workload {name}")


# ----------------------------------------------------------------------------- clocks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    The sampler runs from before the warm-up (nvidia-smi takes ~0.1 s to start) at 20 ms
    intervals; only samples time-stamped inside [mark_start, mark_end] are reported."""

    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t_start = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi attaching to the driver stalls running kernels for several ms: wait
            # until it is past that (first sample printed) before any timed work starts
            t_wait = time.time()
            while not self.lines and time.time() - t_wait < 5.0:
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark_start(self):
        self.t_start = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        import datetime

        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        self.thread.join(timeout=2)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]), float(parts[3]), parts[4:8]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t_start - 0.01 <= r[0] <= self.t_end + 0.01]
        window = "timed region"
        if not inside and rows:  # region shorter than the sampling interval: nearest samples
            mid = 0.5 * (self.t_start + self.t_end)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            window = "nearest samples (timed region shorter than the sampling interval)"
        reasons = set()
        for r in inside:
            for name, val in zip(names, r[4]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None,
                "sm_max_mhz": inside[0][2] if inside else None,
                "power_w_max": max(r[3] for r in inside) if inside else None,
                "reasons": sorted(reasons), "samples": len(inside), "window": window,
                "region_ms": 1e3 * (self.t_end - self.t_start)}


# ----------------------------------------------------------------------------- GPU arm


def particle_seed(rank, p):
    """Seed of particle p of a rank: phase slot k of the particle starts from
    scipy Rotation.random(N, random_state=seed + k) -- in BOTH arms."""
    return 8816 + 100003 * rank + 2 * p


def initial_textures(rank, P, N):
    """[P, 2, N, 3, 3] starting orientations, the same scipy draws the CPU arm makes
    (Mineral.__post_init__, minerals.py:256-285)."""
    from scipy.spatial.transform import Rotation

    o = np.empty((P, 2, N, 3, 3))
    for p in range(P):
        for k in range(2):
            o[p, k] = Rotation.random(N, random_state=particle_seed(rank, p) + k).as_matrix()
    return o


def time_resident(torch, ens, kind, tables, t_edges, K, W, barrier, max_over_ranks, sampler=None):
    """`value` arm: state resident in HBM, one launch per step, CUDA events on the stream."""
    dev = ens.device
    L_dev = [torch.as_tensor(np.ascontiguousarray(t)).to(dev) for t in tables]
    t_lo, t_hi = (torch.as_tensor(np.ascontiguousarray(t)).to(dev) for t in t_edges)
    tol = ens.tolerances()  # the reference's default tolerances
    counters = torch.zeros((4,), dtype=torch.int64, device=dev)

    def count():  # grain right-hand sides; chunk-summed n_rhs, n_accept, n_reject (torch, tiny)
        counters[0] += ens.grain_evals.sum()
        counters[1:] += ens.stats[1:4].sum(dim=1)

    # the warm-up runs exactly what a timed step runs, counter bookkeeping included: CUDA loads
    # kernels lazily, and the first use of torch's small reduction kernel costs ~8 ms of host
    # time that would otherwise land inside the first timed step
    for k in range(W):
        ens.update(t_lo[k], t_hi[k], L_dev[k], kind, tol, check=False)
        count()
    counters.zero_()
    barrier()
    if sampler is not None:
        sampler.mark_start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record()
    for k in range(W, W + K):
        ens.update(t_lo[k], t_hi[k], L_dev[k], kind, tol, check=False)
        count()
        ev[k - W + 1].record()
    barrier()
    if sampler is not None:
        sampler.mark_end()
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[K]))
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    failed = int((ens.stats[0] != 0).sum().item())
    g_evals, n_rhs, n_acc, n_rej = (int(c) for c in counters.cpu())
    return {"total_ms": total_ms, "step_ms": step_ms, "failed": failed, "grain_evals": g_evals,
            "chunk_accepts": n_acc, "chunk_rejects": n_rej}


def parity_check(args, drex, torch, dev, rank, kind, tables, t_edges, o0, n_particles=4, n_steps=6):
    """The benchmark's own parity evidence: the first particles of this rank go through the SAME
    intervals on the GPU (public ensemble API, default tolerances) and through the oracle port on
    the host (C right-hand side + scipy LSODA), from the same textures -- once with LSODA
    converged (rtol 1e-10, atol 1e-12) and once at the reference's default tolerances.  The
    bands of SURVEY.md section 7 (mode B) apply against the converged run; the default run is
    itself only that close to it (`reference_default_vs_converged`)."""
    import multiprocessing as mp

    from oracle import oracle

    oracle.build()
    N = args.grains
    n_particles = min(n_particles, args.particles)
    ens = drex.ParticleEnsemble(n_particles, N, PARAMS, orientations=o0[:n_particles], device=dev)
    for k in range(n_steps):
        ens.update(torch.as_tensor(t_edges[0][k][:n_particles]).to(dev),
                   torch.as_tensor(t_edges[1][k][:n_particles]).to(dev),
                   torch.as_tensor(np.ascontiguousarray(tables[k][:n_particles])).to(dev), kind)
    o_dev = ens.orientations().cpu().numpy()
    f_dev = ens.fractions().cpu().numpy()
    jobs = [(particle_seed(rank, p), N, n_steps, kind, [t[p] for t in tables[:n_steps]],
             (t_edges[0][:n_steps, p], t_edges[1][:n_steps, p])) for p in range(n_particles)]
    tasks = [(j, {"rtol": 1e-10, "atol": 1e-12}) for j in jobs] + [(j, {}) for j in jobs]
    with mp.get_context("fork").Pool(min(len(tasks), len(os.sched_getaffinity(0)))) as pool:
        out = pool.map(_parity_worker, tasks, chunksize=1)
    converged, default = out[:n_particles], out[n_particles:]

    def distance(get_a, get_b):
        d_o, d_f, flips, n = [], [], 0, 0
        for p in range(n_particles):
            for k in range(2):
                (oa, fa), (ob, fb) = get_a(p, k), get_b(p, k)
                d = np.abs(oa - ob).reshape(N, -1).max(axis=1)
                same = d < 1e-2  # a flipped sliding decision puts the grain back to its old orientation
                flips += int((~same).sum())
                n += N
                d_o.append(d[same])
                d_f.append((np.abs(fa - fb) / fb)[same])
        d_o, d_f = np.concatenate(d_o), np.concatenate(d_f)
        return {"median_do": float(np.median(d_o)), "p99_do": float(np.quantile(d_o, 0.99)),
                "p99_df_rel": float(np.quantile(d_f, 0.99)), "mask_mismatch": flips / n}

    dev_state = lambda p, k: (o_dev[p, k], f_dev[p, k])  # noqa: E731
    vs_conv = distance(dev_state, lambda p, k: converged[p][k])
    bands = {"median_do": 1e-4, "p99_do": 1e-3, "p99_df_rel": 2e-2, "mask_mismatch": 1e-3}
    out = {"particles": n_particles, "intervals": n_steps, "grains_compared": 2 * N * n_particles,
           "against": "oracle port (C right-hand side + scipy LSODA) on the same textures and "
                      "velocity-gradient tables",
           "vs_converged": vs_conv,
           "vs_default": distance(dev_state, lambda p, k: default[p][k]),
           "reference_default_vs_converged": distance(lambda p, k: default[p][k],
                                                      lambda p, k: converged[p][k]),
           "bands": bands}
    out["ok"] = bool(all(vs_conv[key] <= bands[key] for key in bands))
    return out


def kernel_rates(drex, _lib, torch, ens, hbm_peak, fp64_peak):
    """Rates of the other kernels of the path at bench size, with the roofline each is bound by
    (ncu captures of the same kernels: profiles/)."""
    import ctypes

    lib, dev = _lib.load(), ens.device
    S, N, P = ens.S, ens.N, ens.P
    st, Pp = _lib.stream_ptr(torch), _lib.ptr

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
        return float(np.mean([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))

    out = {}
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

    def deriv():
        _lib.check(lib.drex_derivatives_f64(
            Pp(ens.o), Pp(ens.f), Pp(D), Pp(Ls), Pp(spin), Pp(m[3]), Pp(m[1]), Pp(m[2]),
            Pp(ens.vol_frac), Pp(mh[3]), Pp(mh[1]), Pp(mh[2]), ctypes.byref(prm), S, N, Pp(do), Pp(df),
            Pp(ws), nws, st))

    ms = timed(deriv)
    gbs = B_ALG_DERIV * S * N / ms * 1e-6
    out["derivatives"] = {"ms": ms, "grain_evals_per_s": S * N / ms * 1e3, "bound": "hbm",
                          "achieved_gbs": gbs, "frac": gbs / hbm_peak, "bytes_per_eval": B_ALG_DERIV}
    ms = timed(ens.voigt_averages)
    tf = 1000.0 * S * N / ms * 1e-9  # ~500 FMA per grain (6x6 congruence)
    out["voigt_average"] = {"ms": ms, "grains_per_s": S * N / ms * 1e3, "bound": "fp64",
                            "achieved_tflops": tf, "frac": tf / fp64_peak if fp64_peak else None,
                            "hbm_gbs": 80.0 * S * N / ms * 1e-6}
    system = drex.LatticeSystem.orthorhombic
    seeds = list(range(P))
    ms = timed(lambda: ens.misorientation_indices(system, phase_slot=0), reps=2, warm=1)
    pairs = P * N * (N - 1) / 2
    out["mindex_full_texture"] = {"ms": ms, "textures": P, "pairs_per_s": pairs / ms * 1e3,
                                  "bound": "fp32", "fp32_tflops_algorithmic": 343.0 * pairs / ms * 1e-9}
    ms = timed(lambda: ens.misorientation_indices(system, phase_slot=0, n_samples=1000, seeds=seeds),
               reps=2, warm=1)
    out["mindex_resampled_1000"] = {"ms": ms, "textures": P,
                                    "what": "resample_orientations + M-index, as the reference's drivers"}
    ms = timed(lambda: ens.resample(1000, seeds, 0, 0, min(P, 256)), reps=3, warm=1)
    out["resample_orientations"] = {"ms": ms, "textures": min(P, 256)}
    C = ens.voigt_averages()
    out["elasticity_components"] = {"ms": timed(lambda: drex.elasticity_components(C)), "matrices": P}
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import pydrex_b200 as drex
    from pydrex_b200 import _lib
    from pydrex_b200.ensemble import bind_host_to_gpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # one process per GPU: keep each rank (and the pinned buffers of its host-buffer arm) on
    # the GPU's own NUMA node
    numa_cores = bind_host_to_gpu(local_rank) if world > 1 and not args.no_numa_bind else []
    P, N, K, W = args.particles, args.grains, args.steps, args.warmup
    n_total = K + W
    kind, tables, t_edges = make_workload(args.workload, P, n_total, seed=1000 + rank)
    o0 = initial_textures(rank, P, N)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- resident-state arm: `value` and the roofline -----------------------
    ens = drex.ParticleEnsemble(P, N, PARAMS, orientations=o0, device=dev)
    S = ens.S
    # state exceeds L2 (126 MB) by far, so consecutive steps cannot hit in L2
    state_bytes = S * N * 10 * 8
    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clock_sampler:
        sampler.start()
    res = time_resident(torch, ens, kind, tables, t_edges, K, W, barrier, max_over_ranks,
                        sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    total_ms, step_ms = res["total_ms"], res["step_ms"]
    failed = int(sum_over_ranks(res["failed"]))
    g_evals = res["grain_evals"]
    evals = sum_over_ranks(g_evals)
    internal = sum_over_ranks(res["chunk_accepts"] * N / ens.n_chunks)  # steps are per chunk of grains
    outer = float(world * S * N * K)
    secs = total_ms * 1e-3
    value = outer / secs
    # kernel duration: the integrator launch (plus the restore of sliding grains that follows it
    # and a 256 B memset) is the whole step on the stream: per-step CUDA-event time
    kernel_ms = float(np.mean(step_ms))
    evals_per_launch = g_evals / K
    fp64_peak = None
    if rank == 0:
        import ctypes

        flops = ctypes.c_double()
        _lib.check(_lib.load().drex_fp64_peak_probe(local_rank, ctypes.byref(flops)))
        fp64_peak = flops.value / 1e12
    barrier()

    # ---------------- host-buffer arm: `e2e` ------------------------------------------------
    e2e = run_e2e(args, drex, _lib, torch, dev, kind, tables, t_edges, barrier, max_over_ranks,
                  world, rank, o0)
    # ---------------- the other workload of the north star, resident arm only ---------------
    other_name = "simple_shear" if args.workload == "corner_flow" else "corner_flow"
    o_kind, o_tables, o_edges = make_workload(other_name, P, n_total, seed=1000 + rank)
    ens2 = drex.ParticleEnsemble(P, N, PARAMS, orientations=o0, device=dev)
    res2 = time_resident(torch, ens2, o_kind, o_tables, o_edges, K, W, barrier, max_over_ranks)
    evals2 = sum_over_ranks(res2["grain_evals"])
    failed += int(sum_over_ranks(res2["failed"]))
    del ens2

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    assert failed == 0, f"{failed} systems failed to integrate"
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    parity = None if args.no_parity_check else parity_check(args, drex, torch, dev, rank, kind,
                                                            tables, t_edges, o0)
    assert parity is None or parity["ok"], f"parity check failed: {parity}"
    pathline_info = trace_pathlines_on_device(drex, torch, P)
    kernels = kernel_rates(drex, _lib, torch, ens, hbm_peak, fp64_peak)
    kernels["pathlines"] = pathline_info
    achieved_tf = F_ALG * evals_per_launch / (kernel_ms * 1e-3) / 1e12
    # algorithmic bytes of the integrator: one read and one write of the state per grain and
    # OUTER interval (80 B each way), whatever the number of internal steps
    hbm_gbs = 160.0 * S * N / (kernel_ms * 1e-3) / 1e9
    # DRAM traffic of the integrator from the committed ncu capture (profiles/), scaled by the
    # number of systems: it does not depend on the number of steps any more
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r02_integrate_traffic.json")))
        if (cap["particles_per_gpu"], cap["grains"]) == (P, N):
            traffic = cap["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        pass
    cpu = None if args.no_cpu_baseline else cpu_baseline(args)
    secs2 = res2["total_ms"] * 1e-3
    line = {
        "metric": "grain-steps/s", "value": value, "unit": "grain-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: {P} particles/GPU x (olivine A + enstatite AB) x "
                        f"{N} grains, GBM M*=125 + GBS chi=0.3, reference default tolerances",
            "step": "one outer CPO interval of every particle = one integrator launch (+ the restore "
                    "of sliding grains)",
            "window": f"outer intervals {W}..{W + K - 1} of each pathline (launch times vary along "
                      "the path: see step_ms)",
            "particles_per_gpu": P, "grains": N, "phases": 2,
            "cache": f"inputs larger than L2: {state_bytes / 1e6:.0f} MB state per GPU vs 126 MB L2",
            "parallelism": f"particles sharded over {world} GPU(s), no data-path collective",
            "host_binding": (f"rank 0 pinned to {len(numa_cores)} cores near its GPU (NVML affinity)"
                             if numa_cores else "none"),
            "textures": "scipy Rotation.random(N, random_state=seed_p + phase): identical in the GPU "
                        "arm, the CPU arm and the parity check",
        },
        "grain_evals_per_s": evals / secs,
        "grain_internal_steps_per_s": internal / secs,
        "rhs_per_grain_and_interval": g_evals / (S * N * K),
        "rejected_chunk_steps": res["chunk_rejects"],
        "failed_systems": failed,
        "parity_check": parity,
        # per step: integrate_kernel, gbs_restore_kernel, next_order_kernel
        "gpu_launches": 3 * K,
        "e2e": e2e,
        "roofline": {
            "kernel": "drex::integrate_kernel", "bound": "fp64",
            "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved_tf / fp64_peak if fp64_peak else None,
            "peak_source": "measured here: drex_fp64_peak_probe (DFMA loop; ncu capture of the probe "
                           "in profiles/); MEASURED_PEAKS.json has no FP64 entry",
            "algorithmic_flop_per_eval": F_ALG, "evals_per_launch": evals_per_launch,
            "launch_ms": kernel_ms,
            "note": "the decoupled integrator needs ~25 % fewer right-hand sides per interval than "
                    "one step sequence per system would; the fraction counts evaluations done",
            "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_gbs / hbm_peak, "bytes_per_grain_interval": 160.0,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
            "traffic": traffic,
            "traffic_source": "profiles/r02_integrate_traffic.json (ncu --set full, dram__bytes_read + "
                              "write of one launch at this size)" if traffic else None,
        },
        "step_ms": step_ms,
        "workloads": {
            other_name: {"value": float(world * S * N * K) / secs2, "unit": "grain-steps/s",
                         "ms_per_step": res2["total_ms"] / K,
                         "grain_evals_per_s": evals2 / secs2,
                         "roofline_frac": (F_ALG * evals2 / world / secs2 / 1e12 / fp64_peak)
                         if fp64_peak else None,
                         "rhs_per_grain_and_interval": res2["grain_evals"] / (S * N * K)},
        },
        "kernels": kernels,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(line))


def run_e2e(args, drex, _lib, torch, dev, kind, tables, t_edges, barrier, max_over_ranks, world,
            rank, o0):
    """Host-buffer path: the particle state lives in pinned host memory in the reference's AoS
    layout, as in a geodynamics code that calls once per model time step.  Every step the
    velocity-gradient tables of that step go in and the whole new state (orientations, fractions,
    deformation gradients) and the counters come back.  The state itself is uploaded once: the
    device copy stays current, and a host copy nobody touched is not sent again
    (ParticleEnsemble.mark_host_dirty)."""
    P, N, K, W = args.particles, args.grains, args.steps, args.warmup
    ens = drex.ParticleEnsemble(P, N, PARAMS, orientations=o0, device=dev)
    S = ens.S
    tol = ens.tolerances()
    h_o = torch.as_tensor(o0.reshape(S, N, 9)).pin_memory()  # AoS like Mineral.orientations
    h_f = ens.f.cpu().pin_memory()
    h_F = ens.F.cpu().pin_memory()
    h_stats = torch.zeros((4, S), dtype=torch.int32).pin_memory()
    h_L = [torch.as_tensor(np.ascontiguousarray(t)).pin_memory() for t in tables]
    h_lo, h_hi = (torch.as_tensor(np.ascontiguousarray(t)).pin_memory() for t in t_edges)
    n_chunks = args.e2e_chunks

    def step(k):
        # public API: pinned host state in, pinned host state out, pipelined over 3 streams;
        # returns after the new state is on the host
        ens.update_host(h_o, h_f, h_F, h_lo[k], h_hi[k], h_L[k], kind, tol, h_stats=h_stats,
                        n_chunks=n_chunks)

    for k in range(W):
        step(k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record()
    for k in range(W, W + K):
        step(k)
    e1.record()
    barrier()
    wall = time.perf_counter() - t_wall
    # the pipeline runs on side streams: the wall clock around the synchronous calls is the
    # honest number (the events on the main stream bracket it too)
    ms = max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3))
    state = S * N * 10 * 8 + S * 9 * 8
    h2d = h_L[0].numel() * 8 + 2 * P * 8
    d2h = state + 4 * S * 4
    if int(h_stats[0].abs().sum()) != 0:
        raise RuntimeError("e2e arm: integration failed")
    return {"value": world * S * N * K / (ms * 1e-3), "unit": "grain-steps/s",
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": ms / K, "gpu_launches_per_step": 4 * n_chunks,  # per group: next_order, integrate, gbs_restore, unpack_soa
            "state_upload_bytes_once": int(state),
            "api": "ParticleEnsemble.update_host",
            "note": "per step: this step's L tables and interval bounds from pinned host memory in, "
                    "the whole new state (orientations AoS + fractions + F) and the counters back "
                    f"to pinned host memory; {n_chunks} particle groups pipelined over compute / "
                    "copy-out streams.  The state is uploaded once (warm-up): the device copy "
                    "stays current and an untouched host copy is not sent again"}


# ----------------------------------------------------------------------------- CPU arm


def _interp_cheb(values, x):  # Chebyshev-Lobatto interpolant on [0, 1]
    K = values.shape[0]
    d = x - 0.5 * (1.0 - np.cos(np.pi * np.arange(K) / (K - 1)))
    if (np.abs(d) < 1e-15).any():
        return values[int(np.argmin(np.abs(d)))]
    w = np.where(np.arange(K) % 2 == 1, -1.0, 1.0)
    w[0] *= 0.5
    w[-1] *= 0.5
    w = w / d
    return np.tensordot(w, values, axes=(0, 0)) / w.sum()


def _make_get_L(kind, nodes, edges, k):
    if kind == 0:
        return lambda t, x: nodes[k][0]
    t0, t1 = edges[0][k], edges[1][k]
    return lambda t, x: _interp_cheb(nodes[k], (t - t0) / (t1 - t0))


def _cpu_particle(job, oracle, **lsoda):
    """One particle, both phases, `n_steps` outer intervals with the oracle port
    (C right-hand side + scipy LSODA), the structure of the reference's per-particle run.
    Returns (seconds, nfev, nsteps, states)."""
    from scipy.spatial.transform import Rotation

    seed, N, n_steps, kind, nodes, edges = job
    states = []
    for phase, fabric in ((0, 0), (1, 5)):
        o = Rotation.random(N, random_state=seed + phase).as_matrix()
        states.append([phase, fabric, o, np.full(N, 1.0 / N)])
    counter = oracle.RhsCounter()
    t_start = time.perf_counter()
    F = np.eye(3)
    for k in range(n_steps):
        get_L = _make_get_L(kind, nodes, edges, k)
        for stt in states:  # update_all: every phase starts from the same F (minerals.py:674-684)
            F_new, stt[2], stt[3] = oracle.update_orientations(
                stt[2], stt[3], 4, stt[0], stt[1], PARAMS, F, get_L,
                (edges[0][k], edges[1][k], lambda t: np.zeros(3)), counter=counter, **lsoda)
        F = F_new
    return time.perf_counter() - t_start, counter.nfev, counter.nsteps, states


def _parity_worker(task):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import oracle

    job, lsoda = task
    states = _cpu_particle(job, oracle, **lsoda)[3]
    return [(st[2], st[3]) for st in states]


def _port_worker(job):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import oracle

    return _cpu_particle(job, oracle)[:3]


def import_reference():
    """The UNMODIFIED reference package, from baseline/_ref (pip-installed there from
    /root/reference, git-ignored; travels to the GPU box) or from /root/reference/src.  Its
    plotting / mesh imports (matplotlib, h5py, meshio, cmcrameri, gmsh, pyvista: absent here and
    not on the CPO path) are replaced by inert stubs.  Returns None when it cannot be imported."""
    import types
    from unittest.mock import MagicMock

    if "pydrex" in sys.modules:
        return sys.modules["pydrex"]
    roots = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"]
    root = next((r for r in roots if os.path.isdir(os.path.join(r, "pydrex"))), None)
    if root is None:
        return None
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")

    def stub(name):
        mod = types.ModuleType(name)
        mod.__getattr__ = lambda attr: MagicMock(name=f"{name}.{attr}")
        mod.__path__ = []
        return mod

    names = ["matplotlib", "matplotlib.axes", "matplotlib.projections", "matplotlib.pyplot",
             "matplotlib.collections", "matplotlib.legend_handler", "matplotlib.transforms",
             "matplotlib.colors", "matplotlib.cm", "matplotlib.figure", "matplotlib.tri",
             "h5py", "meshio", "cmcrameri", "cmcrameri.cm", "gmsh", "pyvista"]
    added = []
    try:
        for name in names:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = stub(name)
                added.append(name)
        mpl = sys.modules["matplotlib"]
        if "matplotlib" in added:
            for name in names:
                if name.startswith("matplotlib."):
                    setattr(mpl, name.split(".", 1)[1], sys.modules[name])

            class Axes:  # subclassed at src/pydrex/axes.py:12
                name = "stub"

            sys.modules["matplotlib.axes"].Axes = Axes
            sys.modules["matplotlib.pyplot"].rcParams = {"figure.figsize": (6.4, 4.8)}
        sys.path.insert(0, root)
        import logging

        import pydrex  # noqa: F401

        logging.getLogger("pydrex").setLevel(logging.ERROR)  # it logs every Mineral it creates

        return sys.modules["pydrex"]
    except Exception as exc:  # numba / scipy missing or incompatible on this box
        sys.stderr.write(f"reference not importable ({exc!r}); falling back to the oracle port\n")
        return None


def _reference_worker(job):
    """The same particle through the reference's own Mineral.update_orientations / update_all
    (numba core.derivatives + scipy LSODA), NUMBA_NUM_THREADS=1, as
    examples/standalone/cornerflow_simple.py:129-151 runs it per particle."""
    os.environ["NUMBA_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    pydrex = import_reference()
    from pydrex import core as rcore
    from pydrex import minerals as rminerals

    seed, N, n_steps, kind, nodes, edges = job
    params = dict(PARAMS, phase_assemblage=(rcore.MineralPhase.olivine, rcore.MineralPhase.enstatite))
    mins = [rminerals.Mineral(rcore.MineralPhase.olivine, rcore.MineralFabric.olivine_A,
                              rcore.DeformationRegime.matrix_dislocation, n_grains=N, seed=seed),
            rminerals.Mineral(rcore.MineralPhase.enstatite, rcore.MineralFabric.enstatite_AB,
                              rcore.DeformationRegime.matrix_dislocation, n_grains=N, seed=seed + 1)]
    nfev = [0]

    t_start = time.perf_counter()
    F = np.eye(3)
    for k in range(n_steps):
        inner = _make_get_L(kind, nodes, edges, k)

        def get_L(t, x, inner=inner):
            nfev[0] += 1
            return inner(t, x)

        F = rminerals.update_all(mins, params, F, get_L, (edges[0][k], edges[1][k], lambda t: np.zeros(3)))
    # every right-hand side asks for L once (minerals.py:391-392); update_orientations asks 3 more
    # times per call for its debug log (minerals.py:492-507)
    rhs = nfev[0] - 3 * 2 * n_steps
    return time.perf_counter() - t_start, rhs, 0


def cpu_run(args, n_particles, n_steps, cores, use_reference):
    """Time the CPU path on `cores` worker processes (one particle per task, like the
    reference's Pool.imap over particles); returns (grain-steps/s, evals/s, wall, kind).
    The pool is started and warmed (imports, numba JIT, library load) before the clock starts."""
    import multiprocessing as mp

    from oracle import oracle

    oracle.build()
    N = args.grains
    worker = _port_worker
    kind_name = "port"
    if use_reference and import_reference() is not None:
        worker, kind_name = _reference_worker, "reference"
    kind, tables, t_edges = make_workload(args.workload, n_particles, n_steps, seed=1000)
    jobs = [(particle_seed(0, p), N, n_steps, kind, [t[p] for t in tables],
             (t_edges[0][:, p], t_edges[1][:, p])) for p in range(n_particles)]
    warm = [(1, 64, 1, kind, [t[p % n_particles] for t in tables],
             (t_edges[0][:, p % n_particles], t_edges[1][:, p % n_particles]))
            for p in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(worker, warm, chunksize=1)
        t0 = time.perf_counter()
        out = pool.map(worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    nfev = sum(o[1] for o in out)
    grain_steps = n_particles * 2 * N * n_steps
    return grain_steps / wall, nfev * N / wall, wall, kind_name


def cpu_baseline(args):
    """The CPU path beside the GPU number, on a bounded sample of the same workload: the
    reference itself when it can be imported on this box, else the oracle port."""
    cores = len(os.sched_getaffinity(0))
    n_steps = 12
    use_ref = import_reference() is not None
    n_particles = (2 if use_ref else 16) * cores  # ~10-30 s of wall time on all cores
    v, evals, wall, kind = cpu_run(args, n_particles, n_steps, cores, use_ref)
    what = ("the unmodified reference (numba core.derivatives + scipy LSODA, NUMBA_NUM_THREADS=1, "
            "one untimed warm-up call per worker)" if kind == "reference"
            else "oracle port (C right-hand side + scipy LSODA)")
    return {"value": v, "unit": "grain-steps/s", "cores": cores, "kind": kind,
            "grain_evals_per_s": evals,
            "sample": f"{n_particles} particles x 2 phases x {args.grains} grains x {n_steps} outer "
                      f"steps of the {args.workload} workload (same textures and tables as the GPU "
                      f"arm's first particles), {what}, reference default tolerances, "
                      f"multiprocessing Pool({cores}); {wall:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    K, W = args.steps, args.warmup
    use_ref = import_reference() is not None
    n_particles = (2 if use_ref else 8) * cores
    # bounded sample: each "step" advances n_particles particles through one outer interval;
    # the warm-up is the pool warm-up inside cpu_run (imports, JIT)
    n_steps = min(K, 12 if use_ref else 30)
    v, evals, wall, kind = cpu_run(args, n_particles, n_steps, cores, use_ref)
    what = ("unmodified reference from baseline/_ref: Mineral.update_all (numba core.derivatives + "
            "scipy LSODA), NUMBA_NUM_THREADS=1" if kind == "reference"
            else "oracle port (C restatement of core.derivatives + scipy LSODA)")
    sample = (f"{n_particles} particles x 2 phases x {args.grains} grains x {n_steps} outer steps "
              f"per timed run (bounded sample of the GPU arm's workload, same textures and tables), "
              f"{what}, Pool({cores})")
    line = {
        "impl": "reference", "metric": "grain-steps/s", "value": v, "unit": "grain-steps/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": K, "warmup": W,
        "ms_per_step": wall * 1e3 / n_steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: bounded CPU sample, {sample}"},
        "grain_evals_per_s": evals,
        "cpu_baseline": {"value": v, "unit": "grain-steps/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": v, "unit": "grain-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=1250, help="particles per GPU")
    ap.add_argument("--grains", type=int, default=5000)
    ap.add_argument("--workload", default="corner_flow", choices=["simple_shear", "corner_flow"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostic: do not run nvidia-smi")
    ap.add_argument("--no-numa-bind", action="store_true", help="diagnostic: do not pin ranks to their GPU's cores")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="pipeline depth of the host-buffer arm")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
