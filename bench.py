#!/usr/bin/env python
"""bench.py -- throughput of the CPO-integration hot path on B200 (and the CPU baseline).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (oracle port)

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
e2e    : the same through the host-buffer API -- every step copies the particles' state
         (orientations + fractions + F, reference AoS layout) and velocity-gradient tables
         from pinned host memory, converts layout on device, integrates, and copies the new
         state and the per-system counters back.
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


def time_diagnostics(drex, torch, ens):
    """BASELINE config 5 on this GPU's particles after the timed steps (outside the timed
    region): Voigt averages over both phases, elasticity decomposition, and the olivine
    M-index of every particle -- on the full 5000-grain texture and on the reference-style
    volume-weighted resample of 1000 grains (tests/test_simple_shear_3d.py:194-202)."""
    system = drex.LatticeSystem.orthorhombic
    seeds = list(range(ens.P))

    def timed(fn):
        fn()  # warm-up (lazy kernel loading, workspace allocation)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out

    ms_voigt, C = timed(ens.voigt_averages)
    ms_elast, _ = timed(lambda: drex.elasticity_components(C))
    ms_full, M_full = timed(lambda: ens.misorientation_indices(system, phase_slot=0))
    ms_res, M_res = timed(lambda: ens.misorientation_indices(system, phase_slot=0, n_samples=1000,
                                                               seeds=seeds))
    pairs = ens.P * ens.N * (ens.N - 1) / 2
    return {"particles": ens.P, "voigt_averages_ms": ms_voigt, "elasticity_components_ms": ms_elast,
            "mindex_full_texture_ms": ms_full, "mindex_pairs_per_s": pairs / (ms_full * 1e-3),
            "mindex_resampled_1000_ms": ms_res,
            "mindex_mean_full": float(M_full.mean()), "mindex_mean_resampled": float(M_res.mean())}


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
    L_dev = [torch.as_tensor(np.ascontiguousarray(t)).to(dev) for t in tables]
    t_lo, t_hi = (torch.as_tensor(np.ascontiguousarray(t)).to(dev) for t in t_edges)

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
    ens = drex.ParticleEnsemble(P, N, PARAMS, seed=8816 + rank, device=dev)
    tol = ens.tolerances()  # the reference's default tolerances
    S = ens.S
    # state + scratch exceed L2 (126 MB) by far, so consecutive steps cannot hit in L2
    state_bytes = S * N * 10 * 8
    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clock_sampler:
        sampler.start()
    # the warm-up runs exactly what a timed step runs, counter bookkeeping included: CUDA loads
    # kernels lazily, and the first use of torch's small reduction kernel costs ~8 ms of host
    # time that would otherwise land inside the first timed step
    counters = torch.zeros((4,), dtype=torch.int64, device=dev)

    def count():  # grain right-hand sides; chunk-summed n_rhs, n_accept, n_reject (torch, tiny)
        counters[0] += ens.grain_evals.sum()
        counters[1:] += ens.stats[1:4].sum(dim=1)

    for k in range(W):
        ens.update(t_lo[k], t_hi[k], L_dev[k], kind, tol, check=False)
        count()
    counters.zero_()
    step_ms = []
    barrier()
    sampler.mark_start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record()
    for k in range(W, n_total):
        ens.update(t_lo[k], t_hi[k], L_dev[k], kind, tol, check=False)
        count()
        ev[k - W + 1].record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[K]))
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    status_bad = int((ens.stats[0] != 0).sum().item())
    g_evals, n_rhs, n_acc, n_rej = (int(c) for c in counters.cpu())
    evals = sum_over_ranks(g_evals)
    internal = sum_over_ranks(n_acc * N / ens.n_chunks)  # steps are per chunk of 32 grains
    outer = float(world * S * N * K)
    secs = total_ms * 1e-3
    value = outer / secs

    # kernel-only duration: the integrator launch is the whole step on the stream, so the
    # per-step CUDA-event time IS the kernel's launch duration (plus a 256 B memset)
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
                  world, rank)
    e2e_resident = run_e2e_resident(args, drex, torch, dev, kind, tables, t_edges, barrier,
                                    max_over_ranks, world, rank)

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    pathline_info = trace_pathlines_on_device(drex, torch, P) if args.workload == "corner_flow" else None
    diagnostics_info = time_diagnostics(drex, torch, ens)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    achieved_tf = F_ALG * evals_per_launch / (kernel_ms * 1e-3) / 1e12
    # bytes the integrator moves per grain per attempted step (csrc/drex_integrate.cuh):
    # orientation + FSAL slope in, both out (4 x 72 B), 3 stage energies, 10 fraction-stage
    # doubles for olivine; enstatite skips the fraction stages.  Averaged over both phases.
    hbm_gbs = 160.0 * S * N / (kernel_ms * 1e-3) / 1e9
    # DRAM traffic of the integrator from the committed ncu capture (profiles/), scaled by
    # right-hand sides per launch: the capture's launch and this run's launches differ only
    # in how many evaluations the step-size control needed
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r01_integrate_traffic.json")))
        if (cap["workload"], cap["particles_per_gpu"], cap["grains"]) == (args.workload, P, N):
            traffic = cap["dram_bytes_per_launch"] / cap["evals_per_launch"] * evals_per_launch
    except (OSError, KeyError, ValueError):
        pass
    cpu = None if args.no_cpu_baseline else cpu_baseline(args)
    line = {
        "metric": "grain-steps/s", "value": value, "unit": "grain-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: {P} particles/GPU x (olivine A + enstatite AB) x "
                        f"{N} grains, GBM M*=125 + GBS chi=0.3, reference default tolerances",
            "step": "one outer CPO interval of every particle = one integrator launch",
            "particles_per_gpu": P, "grains": N, "phases": 2,
            "cache": f"inputs larger than L2: {state_bytes / 1e6:.0f} MB state + "
                     f"{ens._ws.numel() / 1e6:.0f} MB scratch per GPU vs 126 MB L2",
            "parallelism": f"particles sharded over {world} GPU(s), no data-path collective",
            "host_binding": (f"rank 0 pinned to {len(numa_cores)} cores near its GPU (NVML affinity)"
                             if numa_cores else "none"),
        },
        "grain_evals_per_s": evals / secs,
        "grain_internal_steps_per_s": internal / secs,
        "rhs_per_outer_step": g_evals / (S * N * K), "rejected_chunk_steps": n_rej,
        "failed_systems": status_bad,
        "gpu_launches": K,
        "e2e": e2e,
        "e2e_resident_state": e2e_resident,
        "roofline": {
            "kernel": "drex::integrate_kernel", "bound": "fp64",
            "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved_tf / fp64_peak if fp64_peak else None,
            "peak_source": "measured here: drex_fp64_peak_probe (DFMA loop); MEASURED_PEAKS.json "
                           "has no FP64 entry",
            "algorithmic_flop_per_eval": F_ALG, "evals_per_launch": evals_per_launch,
            "launch_ms": kernel_ms,
            "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_gbs / hbm_peak, "bytes_per_grain_interval": 160.0,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)"
                    if peaks else "fallback 6650 GB/s"},
            "traffic": traffic,
            "traffic_source": "profiles/r01_integrate_traffic.json (ncu --set full, dram__bytes_read+write "
                              "per launch) scaled by evaluations per launch" if traffic else None,
        },
        "step_ms": step_ms,
        "pathlines": pathline_info,
        "diagnostics": diagnostics_info,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(line))


def run_e2e(args, drex, _lib, torch, dev, kind, tables, t_edges, barrier, max_over_ranks, world,
            rank):
    """Host-buffer path: state lives in pinned host memory in the reference's AoS layout and
    round-trips every step, as a geodynamics code calling once per model time step would."""
    import ctypes

    P, N, K, W = args.particles, args.grains, args.steps, args.warmup
    ens = drex.ParticleEnsemble(P, N, PARAMS, seed=8816 + rank, device=dev)
    S = ens.S
    tol = ens.tolerances()
    f64 = torch.float64
    h_o = ens.orientations().reshape(S, N, 9).cpu().pin_memory()  # AoS like Mineral.orientations
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
    h2d = state + h_L[0].numel() * 8 + 2 * P * 8
    d2h = state + 4 * S * 4
    return {"value": world * S * N * K / (ms * 1e-3), "unit": "grain-steps/s",
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": ms / K, "gpu_launches_per_step": 3 * n_chunks,
            "api": "ParticleEnsemble.update_host",
            "note": "state (orientations AoS + fractions + F) and L tables copied from pinned "
                    f"host memory every step, new state and counters copied back; {n_chunks} "
                    "particle groups pipelined over copy-in / compute / copy-out streams"}


def run_e2e_resident(args, drex, torch, dev, kind, tables, t_edges, barrier, max_over_ranks, world,
                     rank):
    """The other way to drive the public API: the particle cloud stays resident on the GPU
    (ParticleEnsemble.update); per step only that step's inputs -- the velocity-gradient
    tables and interval bounds -- come from pinned host memory, and the step's results -- the
    per-system status/counters and the deformation gradients -- go back to the host."""
    P, N, K, W = args.particles, args.grains, args.steps, args.warmup
    ens = drex.ParticleEnsemble(P, N, PARAMS, seed=8816 + rank, device=dev)
    S = ens.S
    tol = ens.tolerances()
    h_L = [torch.as_tensor(np.ascontiguousarray(t)).pin_memory() for t in tables]
    h_lo, h_hi = (torch.as_tensor(np.ascontiguousarray(t)).pin_memory() for t in t_edges)
    h_stats = torch.zeros((4, S), dtype=torch.int32).pin_memory()
    h_F = torch.zeros((S, 9), dtype=torch.float64).pin_memory()

    def step(k):
        Lk = h_L[k].to(dev, non_blocking=True)
        t0 = h_lo[k].to(dev, non_blocking=True)
        t1 = h_hi[k].to(dev, non_blocking=True)
        ens.update(t0, t1, Lk, kind, tol, check=False)
        h_stats.copy_(ens.stats, non_blocking=True)
        h_F.copy_(ens.F, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if int(h_stats[0].abs().sum()) != 0:
            raise RuntimeError("integration failed")

    for k in range(W):
        step(k)
    barrier()
    t_wall = time.perf_counter()
    for k in range(W, W + K):
        step(k)
    barrier()
    ms = max_over_ranks((time.perf_counter() - t_wall) * 1e3)
    return {"value": world * S * N * K / (ms * 1e-3), "unit": "grain-steps/s",
            "h2d_bytes_per_step": int(h_L[0].numel() * 8 + 2 * P * 8),
            "d2h_bytes_per_step": int(4 * S * 4 + S * 9 * 8), "ms_per_step": ms / K,
            "api": "ParticleEnsemble.update",
            "note": "state resident in HBM; per step L tables + interval bounds in, status/counters "
                    "+ deformation gradients out; wall clock around synchronous steps"}


# ----------------------------------------------------------------------------- CPU arm


def _cpu_worker(job):
    """One particle, both phases, `n_steps` outer intervals with the oracle port
    (C right-hand side + scipy LSODA), the structure of the reference's per-particle run."""
    os.environ["OMP_NUM_THREADS"] = "1"
    from scipy.spatial.transform import Rotation

    from oracle import oracle

    seed, N, n_steps, kind, nodes, edges = job

    def barycentric_interpolate(values, x):  # Chebyshev-Lobatto interpolant on [0, 1]
        K = values.shape[0]
        d = x - 0.5 * (1.0 - np.cos(np.pi * np.arange(K) / (K - 1)))
        if (np.abs(d) < 1e-15).any():
            return values[int(np.argmin(np.abs(d)))]
        w = np.where(np.arange(K) % 2 == 1, -1.0, 1.0)
        w[0] *= 0.5
        w[-1] *= 0.5
        w = w / d
        return np.tensordot(w, values, axes=(0, 0)) / w.sum()

    def make_get_L(k):
        if kind == 0:
            return lambda t, x: nodes[k][0]
        t0, t1 = edges[0][k], edges[1][k]
        return lambda t, x: barycentric_interpolate(nodes[k], (t - t0) / (t1 - t0))

    states = []
    for phase, fabric in ((0, 0), (1, 5)):
        o = Rotation.random(N, random_state=seed + phase).as_matrix()
        states.append([phase, fabric, o, np.full(N, 1.0 / N)])
    counter = oracle.RhsCounter()
    t_start = time.perf_counter()
    F = np.eye(3)
    for k in range(n_steps):
        get_L = make_get_L(k)
        for stt in states:  # update_all: every phase starts from the same F (minerals.py:674-684)
            F_new, stt[2], stt[3] = oracle.update_orientations(
                stt[2], stt[3], 4, stt[0], stt[1], PARAMS, F, get_L,
                (edges[0][k], edges[1][k], lambda t: np.zeros(3)), counter=counter)
        F = F_new
    return time.perf_counter() - t_start, counter.nfev, counter.nsteps


def cpu_run(args, n_particles, n_steps, cores):
    """Time the CPU port on `cores` worker processes (one particle per task, like the
    reference's Pool.imap over particles); returns (grain-steps/s, evals/s, steps/s, wall).
    The pool is started and warmed (imports, library load) before the clock starts."""
    import multiprocessing as mp

    from oracle import oracle

    oracle.build()
    N = args.grains
    kind, tables, t_edges = make_workload(args.workload, n_particles, n_steps, seed=1000)
    jobs = [(8816 + 2 * p, N, n_steps, kind, [t[p] for t in tables],
             (t_edges[0][:, p], t_edges[1][:, p])) for p in range(n_particles)]
    warm = [(1, 64, 1, kind, [t[p % n_particles] for t in tables],
             (t_edges[0][:, p % n_particles], t_edges[1][:, p % n_particles]))
            for p in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_worker, warm, chunksize=1)
        t0 = time.perf_counter()
        out = pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    nfev = sum(o[1] for o in out)
    nsteps = sum(o[2] for o in out)
    grain_steps = n_particles * 2 * N * n_steps
    return grain_steps / wall, nfev * N / wall, nsteps * N / wall, wall


def cpu_baseline(args):
    cores = len(os.sched_getaffinity(0))
    n_particles = 16 * cores  # ~10-20 s of wall time on all cores
    n_steps = 12
    v, evals, internal, wall = cpu_run(args, n_particles, n_steps, cores)
    return {"value": v, "unit": "grain-steps/s", "cores": cores, "kind": "port",
            "grain_evals_per_s": evals, "grain_internal_steps_per_s": internal,
            "sample": f"{n_particles} particles x 2 phases x {args.grains} grains x {n_steps} outer "
                      f"steps of the {args.workload} workload, oracle port (C RHS + scipy LSODA, "
                      f"reference tolerances), multiprocessing Pool({cores}); {wall:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    K, W = args.steps, args.warmup
    n_particles = 8 * cores
    # bounded sample: each "step" advances n_particles particles through one outer interval;
    # the warm-up is the pool warm-up inside cpu_run
    n_steps = min(K, 30)
    v, evals, internal, wall = cpu_run(args, n_particles, n_steps, cores)
    sample = (f"{n_particles} particles x 2 phases x {args.grains} grains x {n_steps} outer steps "
              f"per timed run (bounded sample of the GPU arm's workload), oracle port "
              f"(C restatement of core.derivatives + scipy LSODA), Pool({cores})")
    line = {
        "impl": "reference", "metric": "grain-steps/s", "value": v, "unit": "grain-steps/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": K, "warmup": W,
        "ms_per_step": wall * 1e3 / n_steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: bounded CPU sample, {sample}"},
        "grain_evals_per_s": evals, "grain_internal_steps_per_s": internal,
        "cpu_baseline": {"value": v, "unit": "grain-steps/s", "cores": cores, "kind": "port",
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
