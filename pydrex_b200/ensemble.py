"""Batched pathline-particle ensembles: the fast path the B200 kernels are built for.

The reference parallelises over particles with one OS process each
(`Pool.imap(run, final_locations)`, examples/standalone/cornerflow_simple.py:129-151, or
Ray on a PBS cluster).  Here a whole cloud of particles, each carrying one grain ensemble
per mineral phase, stays resident in HBM in SoA layout and advances with ONE launch of
the persistent integrator per outer time step.  Particles are independent, so multi-GPU
runs shard particle index ranges over ranks with no data-path collective; only results
are gathered at the end (`gather_to_rank0`).

A "system" is one (particle, phase) grain ensemble; system index s = particle * n_phases +
phase_slot.
"""

import ctypes

import numpy as np

from . import _lib
from . import core as _core
from . import exceptions as _err
from . import stats as _stats
from .minerals import StiffnessTensors


def shard_range(n_particles, rank, world_size):
    """Contiguous particle range [lo, hi) owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(int(n_particles), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_to_rank0(local, n_particles, group=None):
    """Gather per-particle result rows (a tensor [P_local, ...]) onto rank 0 in particle
    order.  The only inter-rank traffic of a run; uses the default process group (NCCL for
    CUDA tensors, gloo for CPU tensors).  Returns the full tensor on rank 0, None elsewhere.
    """
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_particles, r, world) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros((max_rows, *local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def random_orientations(n_systems, n_grains, seed, device):
    """Haar-random rotation matrices [S, N, 3, 3] from normalised Gaussian quaternions,
    generated on `device` (what scipy's Rotation.random does, minerals.py:261-263, but with
    torch's generator: seeds are NOT interchangeable with the reference's)."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    q = torch.randn((n_systems, n_grains, 4), dtype=torch.float64, device=device, generator=gen)
    q = q / q.norm(dim=-1, keepdim=True)
    x, y, z, w = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(n_systems, n_grains, 3, 3)


def _on_own_device(method):
    """Run a ParticleEnsemble method with the ensemble's GPU as the current CUDA device: the
    library sizes its launches from cudaGetDevice() and the binding passes torch's current
    stream, both of which belong to the CURRENT device, not to the tensors' device."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        with self._torch.cuda.device(self.device):
            return method(self, *args, **kwargs)

    return wrapper


class ParticleEnsemble:
    """P particles x len(phases) mineral phases x N grains, resident on one GPU.

    - `phases`, `fabrics`, `regimes` — one entry per phase slot (MineralPhase / MineralFabric
      / DeformationRegime ordinals)
    - `params` — dict with the keys of `DefaultParams.as_dict()`; `phase_fractions` are
      matched to `phases` through `phase_assemblage` exactly as minerals.py:412-415 does
    - `orientations` — optional initial orientations [P, n_phases, N, 3, 3] (numpy or
      torch); default Haar-random from `seed`
    - `fractions` — optional initial fractions [P, n_phases, N]; default 1/N
    """

    def __init__(self, n_particles, n_grains, params, phases=(0, 1), fabrics=(0, 5),
                 regimes=(4, 4), orientations=None, fractions=None, seed=8816, device=None):
        torch = _lib.require_cuda()
        self._torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.P, self.N, self.n_phases = int(n_particles), int(n_grains), len(phases)
        self.S = self.P * self.n_phases
        self.params = dict(params)
        self.phases = [int(p) for p in phases]
        self.fabrics = [int(f) for f in fabrics]
        for ph, fab in zip(self.phases, self.fabrics):
            _core.get_crss(ph, fab)  # ValueError for unsupported combinations
        assemblage = list(self.params["phase_assemblage"])
        vol = [float(self.params["phase_fractions"][assemblage.index(ph)]) for ph in self.phases]
        f64, i32, dev = torch.float64, torch.int32, self.device
        P, K, N, S = self.P, self.n_phases, self.N, self.S
        meta = np.empty((4, S), dtype=np.int32)
        meta[0] = np.repeat(np.arange(P, dtype=np.int32), K)
        meta[1] = np.tile(np.asarray(self.phases, dtype=np.int32), P)
        meta[2] = np.tile(np.asarray(self.fabrics, dtype=np.int32), P)
        meta[3] = np.tile(np.asarray([int(r) for r in regimes], dtype=np.int32), P)
        self.meta_host = meta
        self.meta = torch.as_tensor(meta).to(dev)
        self.vol_frac = torch.tensor(vol, dtype=f64, device=dev).repeat(P)
        self.F = torch.eye(3, dtype=f64, device=dev).reshape(1, 9).repeat(S, 1).contiguous()
        self.o = torch.empty((S, 9, N), dtype=f64, device=dev)  # SoA
        if orientations is None:
            o_aos = random_orientations(S, N, seed, dev)
        else:
            o_aos = torch.as_tensor(orientations).to(device=dev, dtype=f64)
        self._pack(o_aos.reshape(S, N, 9).contiguous(), self.o)
        if fractions is None:
            self.f = torch.full((S, N), 1.0 / N, dtype=f64, device=dev)
        else:
            self.f = torch.as_tensor(fractions).to(device=dev, dtype=f64).reshape(S, N).contiguous()
        # Work-queue order (longest-processing-time first): an olivine right-hand side costs
        # about four enstatite ones; refined after every update from the measured counts.
        self._weight = torch.tensor([4.0 if ph == 0 else 1.0 for ph in self.phases],
                                    dtype=torch.float32, device=dev).repeat(P)
        self.order = torch.argsort(self._weight, descending=True, stable=True).to(i32)
        # the integrator steps chunks of grains independently: one carried step per chunk
        self.chunk_grains = _lib.chunk_grains()
        self.n_chunks = -(-N // self.chunk_grains)
        self.h = torch.zeros((S, self.n_chunks), dtype=f64, device=dev)
        # status, n_rhs, n_accept, n_reject per system, the counters summed over its chunks
        self.stats = torch.zeros((4, S), dtype=i32, device=dev)
        self.grain_evals = torch.zeros(S, dtype=torch.int64, device=dev)  # per-grain right-hand sides
        self._o_alt = None  # second orientation buffer: a launch reads one and writes the other
        self._ws = None
        self._host_dirty = True  # update_host: the host copy of the state has not been uploaded
        self.total_rhs = 0    # right-hand sides x grains, accumulated by update(check=True)
        self.total_steps = 0  # accepted steps summed over chunks

    # -- layout --------------------------------------------------------------------------
    @_on_own_device
    def _pack(self, aos, soa):
        lib, torch = _lib.load(), self._torch
        S, N = soa.shape[0], soa.shape[2]
        _lib.check(lib.drex_pack_soa_f64(_lib.ptr(aos), _lib.ptr(soa), S, N, _lib.stream_ptr(torch)))

    @_on_own_device
    def orientations(self):
        """Current orientations as [P, n_phases, N, 3, 3] (device tensor, AoS like the reference)."""
        lib, torch = _lib.load(), self._torch
        aos = torch.empty((self.S, self.N, 9), dtype=torch.float64, device=self.device)
        _lib.check(lib.drex_unpack_soa_f64(_lib.ptr(self.o), _lib.ptr(aos), self.S, self.N,
                                           _lib.stream_ptr(torch)))
        return aos.reshape(self.P, self.n_phases, self.N, 3, 3)

    def fractions(self):
        return self.f.reshape(self.P, self.n_phases, self.N)

    def deformation_gradients(self):
        """F per particle (the last phase's, as update_all returns it, minerals.py:684)."""
        return self.F.reshape(self.P, self.n_phases, 3, 3)[:, -1]

    # -- integration -----------------------------------------------------------------------
    def tolerances(self, rtol=1e-6, atol_abs=1e-4, atol_rel=1e-6, first_step_frac=0.0,
                   max_step_frac=0.0, max_steps=100000, atol_abs_F=None, atol_abs_f=None):
        """Error weights atol_abs + (atol_rel + rtol) |y| (defaults: minerals.py:523-524);
        `atol_abs` applies to the orientations and, unless given separately, to the
        deformation gradient (`atol_abs_F`) and the volume fractions (`atol_abs_f`)."""
        tol = _lib.DrexTolerances()
        tol.rtol, tol.atol_abs, tol.atol_rel = rtol, atol_abs, atol_rel
        tol.atol_abs_F = atol_abs if atol_abs_F is None else atol_abs_F
        tol.atol_abs_f = atol_abs if atol_abs_f is None else atol_abs_f
        tol.first_step_frac, tol.max_step_frac = first_step_frac, max_step_frac
        tol.max_steps, tol.method = max_steps, 0
        return tol

    @_on_own_device
    def update(self, t0, t1, L_nodes, kind=_lib.L_CONSTANT, tol=None, check=True):
        """Advance every particle from t0[p] to t1[p] (device tensors [P] or floats).

        `L_nodes` is a device tensor [P, K, 3, 3] (or [P, 3, 3] for a constant L) holding the
        velocity-gradient table of each particle for this interval; `kind` one of
        `_lib.L_CONSTANT / L_LINEAR / L_CHEBYSHEV / L_PIECEWISE`.
        One kernel launch; returns the per-system status tensor (raises `IterationError`
        on failure when `check`, which synchronises).
        """
        lib, torch = _lib.load(), self._torch
        dev, f64 = self.device, torch.float64
        if not torch.is_tensor(t0):
            t0 = torch.full((self.P,), float(t0), dtype=f64, device=dev)
        if not torch.is_tensor(t1):
            t1 = torch.full((self.P,), float(t1), dtype=f64, device=dev)
        L_nodes = L_nodes.to(device=dev, dtype=f64)
        if L_nodes.dim() == 3:
            L_nodes = L_nodes[:, None]
        K = int(L_nodes.shape[1])
        L_nodes = L_nodes.reshape(self.P, K, 9).contiguous()
        if tol is None:
            tol = self.tolerances()
        p = self.params
        prm = _lib.DrexParams(float(p["stress_exponent"]), float(p["deformation_exponent"]),
                              float(p["nucleation_efficiency"]), float(p["gbm_mobility"]),
                              float(p["gbs_threshold"]))
        self._workspace()
        Pp = _lib.ptr
        m, mh = self.meta, self.meta_host
        _lib.check(lib.drex_integrate_f64(
            self.S, self.P, self.N, Pp(m[0]), Pp(m[1]), Pp(m[2]), Pp(m[3]), Pp(self.vol_frac),
            Pp(mh[1]), Pp(mh[2]), Pp(mh[3]), Pp(self.F), Pp(self.o), Pp(self.f), None,
            Pp(self._o_alt), Pp(self.f), Pp(t0), Pp(t1), int(kind), K, Pp(L_nodes),
            ctypes.byref(prm), ctypes.byref(tol), Pp(self.h), Pp(self.order), Pp(self.stats[0]),
            Pp(self.stats[1]), Pp(self.stats[2]), Pp(self.stats[3]), Pp(self.grain_evals),
            Pp(self._ws), self._ws.numel(), _lib.stream_ptr(torch)))
        self.o, self._o_alt = self._o_alt, self.o
        # next launch: systems that needed the most right-hand sides go first
        self._next_order(0, self.S, self.order)
        if check:
            stats = self.stats.cpu().numpy()
            bad = np.nonzero(stats[0])[0]
            if bad.size:
                s = int(bad[0])
                raise _err.IterationError(
                    f"{bad.size} system(s) failed; first: particle {s // self.n_phases}, phase slot "
                    f"{s % self.n_phases}: {_lib.STATUS_MESSAGES.get(int(stats[0, s]))}")
            self.total_rhs += int(self.grain_evals.sum().item())
            self.total_steps += int(stats[2].sum())
        return self.stats[0]

    def _next_order(self, s0, s1, out):
        """Work-queue order of systems [s0, s1) for their next launch (indices relative to s0):
        most expensive first by the measured right-hand sides of the last one."""
        lib, torch = _lib.load(), self._torch
        n = s1 - s0
        if n <= 16384:
            _lib.check(lib.drex_integrate_next_order(
                _lib.ptr(self.grain_evals[s0:s1]), _lib.ptr(self.meta[1][s0:s1]), n, _lib.ptr(out),
                _lib.stream_ptr(torch)))
        else:
            cost = self.grain_evals[s0:s1].to(torch.float32) * self._weight[s0:s1]
            out.copy_(torch.argsort(cost, descending=True, stable=True).to(torch.int32))

    def _workspace(self):
        """Workspace and second orientation buffer of the integrator, allocated on first use."""
        lib, torch = _lib.load(), self._torch
        if self._ws is None:
            nbytes = lib.drex_integrate_workspace_bytes(self.S, self.N, self.device.index or 0, 0)
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        if self._o_alt is None:
            self._o_alt = torch.empty_like(self.o)

    def mark_host_dirty(self):
        """Tell `update_host` that the caller changed the host copy of the state (orientations,
        fractions or deformation gradients) since the last call: it is uploaded again before
        the next integration.  Without this the device copy, which `update_host` keeps current,
        is used and only the velocity-gradient tables cross the bus on the way in."""
        self._host_dirty = True

    @_on_own_device
    def update_host(self, h_o, h_f, h_F, t0, t1, L_nodes, kind=_lib.L_CONSTANT, tol=None,
                    h_stats=None, n_chunks=8, upload=None):
        """Advance particles whose state lives in (pinned) HOST memory in the reference's
        layout and write the new state back: the call a geodynamics code makes once per
        model time step (cf. the Fluidity coupling, examples/fluidity/corner2d).

        `h_o` [S, N, 3, 3] (AoS like `Mineral.orientations[-1]`), `h_f` [S, N], `h_F` [S, 3, 3]
        are CPU tensors updated IN PLACE; `t0`, `t1` [P] and `L_nodes` [P, K, 3, 3] are CPU or
        device tensors; `h_stats` optional int32 [4, S] CPU tensor receiving status / n_rhs /
        n_accept / n_reject.  The particles are cut into `n_chunks` groups and pipelined over
        three streams so the host->device copy of group i+1, the integration of group i and
        the device->host copy of group i-1 overlap (PCIe is full duplex).  Synchronises
        before returning.

        The device keeps the state it has just written back, so a host copy nobody touched is
        NOT uploaded again: `upload=None` (default) uploads on the first call, after
        `mark_host_dirty()`, and when different host buffers are passed; `upload=True/False`
        forces the choice.  The new state is always copied back to the host.
        """
        lib, torch = _lib.load(), self._torch
        dev, f64 = self.device, torch.float64
        S, N, K_ph = self.S, self.N, self.n_phases
        if tol is None:
            tol = self.tolerances()
        per = -(-self.P // n_chunks) * K_ph
        if getattr(self, "_pipe", None) is None or self._pipe["per"] < per:
            # separate double-buffered staging for arriving and leaving orientations: with
            # one shared pair the copy-in of group c+2 had to wait for the copy-out of group c
            self._pipe = {
                "per": per,
                "streams": [torch.cuda.Stream(device=dev) for _ in range(3)],
                "aos_in": [torch.empty((per, N, 9), dtype=f64, device=dev) for _ in range(2)],
                "aos_out": [torch.empty((per, N, 9), dtype=f64, device=dev) for _ in range(2)],
                "order": [torch.empty(per, dtype=torch.int32, device=dev) for _ in range(2)],
            }
        s_in, s_run, s_out = self._pipe["streams"]
        main = torch.cuda.current_stream(dev)
        t0 = torch.as_tensor(t0, dtype=f64).to(dev, non_blocking=True)
        t1 = torch.as_tensor(t1, dtype=f64).to(dev, non_blocking=True)
        L_nodes = torch.as_tensor(L_nodes, dtype=f64).to(dev, non_blocking=True)
        if L_nodes.dim() == 3:
            L_nodes = L_nodes[:, None]
        K = int(L_nodes.shape[1])
        L_nodes = L_nodes.reshape(self.P, K, 9).contiguous()
        ready = torch.cuda.Event()
        ready.record(main)
        for st in (s_in, s_run, s_out):
            st.wait_event(ready)
        p = self.params
        prm = _lib.DrexParams(float(p["stress_exponent"]), float(p["deformation_exponent"]),
                              float(p["nucleation_efficiency"]), float(p["gbm_mobility"]),
                              float(p["gbs_threshold"]))
        self._workspace()
        host_id = (h_o.data_ptr(), h_f.data_ptr(), h_F.data_ptr())
        if upload is None:
            upload = self._host_dirty or getattr(self, "_host_id", None) != host_id
        self._host_id = host_id
        h_o3 = h_o.reshape(S, N, 9)
        h_F2 = h_F.reshape(S, 9)
        Pp = _lib.ptr
        m, mh = self.meta, self.meta_host
        per_p = -(-self.P // n_chunks)
        out_done, packed = [None, None], [None, None]
        for c in range(n_chunks):
            p0, p1 = c * per_p, min(self.P, (c + 1) * per_p)
            if p0 >= p1:
                break
            s0, s1 = p0 * K_ph, p1 * K_ph
            n = s1 - s0
            aos = self._pipe["aos_in"][c % 2][:n]
            aos_out = self._pipe["aos_out"][c % 2][:n]
            if upload:
                with torch.cuda.stream(s_in):
                    if packed[c % 2] is not None:
                        s_in.wait_event(packed[c % 2])  # arrival buffer consumed by the pack kernel
                    aos.copy_(h_o3[s0:s1], non_blocking=True)
                    self.f[s0:s1].copy_(h_f[s0:s1], non_blocking=True)
                    self.F[s0:s1].copy_(h_F2[s0:s1], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(s_in)
            with torch.cuda.stream(s_run):
                sp = _lib.stream_ptr(torch)
                # longest systems of the group first (from the previous step's counts): a group
                # is only ~2 systems per CTA, so the order decides the tail of its launch
                order = self._pipe["order"][c % 2][:n]
                self._next_order(s0, s1, order)  # from the previous step's counts
                if upload:
                    s_run.wait_event(ev_in)
                    _lib.check(lib.drex_pack_soa_f64(Pp(aos), Pp(self.o[s0:s1]), n, N, sp))
                    packed[c % 2] = torch.cuda.Event()
                    packed[c % 2].record(s_run)
                _lib.check(lib.drex_integrate_f64(
                    n, self.P, N, Pp(m[0][s0:s1]), Pp(m[1][s0:s1]), Pp(m[2][s0:s1]),
                    Pp(m[3][s0:s1]), Pp(self.vol_frac[s0:s1]), Pp(mh[1][s0:s1]), Pp(mh[2][s0:s1]),
                    Pp(mh[3][s0:s1]), Pp(self.F[s0:s1]), Pp(self.o[s0:s1]), Pp(self.f[s0:s1]), None,
                    Pp(self._o_alt[s0:s1]), Pp(self.f[s0:s1]), Pp(t0), Pp(t1), int(kind), K,
                    Pp(L_nodes), ctypes.byref(prm), ctypes.byref(tol), Pp(self.h[s0:s1]), Pp(order),
                    Pp(self.stats[0][s0:s1]), Pp(self.stats[1][s0:s1]), Pp(self.stats[2][s0:s1]),
                    Pp(self.stats[3][s0:s1]), Pp(self.grain_evals[s0:s1]), Pp(self._ws),
                    self._ws.numel(), sp))
                if out_done[c % 2] is not None:
                    s_run.wait_event(out_done[c % 2])  # departure buffer copied out
                _lib.check(lib.drex_unpack_soa_f64(Pp(self._o_alt[s0:s1]), Pp(aos_out), n, N, sp))
                ev_run = torch.cuda.Event()
                ev_run.record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_run)
                h_o3[s0:s1].copy_(aos_out, non_blocking=True)
                h_f[s0:s1].copy_(self.f[s0:s1], non_blocking=True)
                h_F2[s0:s1].copy_(self.F[s0:s1], non_blocking=True)
                out_done[c % 2] = torch.cuda.Event()
                out_done[c % 2].record(s_out)
        if h_stats is not None:
            with torch.cuda.stream(s_out):
                h_stats.copy_(self.stats, non_blocking=True)
        s_out.synchronize()
        main.wait_stream(s_run)
        self.o, self._o_alt = self._o_alt, self.o  # the device copy stays current
        self._host_dirty = False

    # -- diagnostics -----------------------------------------------------------------------
    @_on_own_device
    def voigt_averages(self, elastic_tensors=None):
        """Voigt-averaged stiffness per particle, [P, 6, 6] (minerals.voigt_averages for the
        current snapshot of every particle)."""
        lib, torch = _lib.load(), self._torch
        dev, f64 = self.device, torch.float64
        tensors = list(elastic_tensors if elastic_tensors is not None else StiffnessTensors())
        assemblage = list(self.params["phase_assemblage"])
        stiff = np.stack([np.asarray(tensors[assemblage.index(ph)], dtype=np.float64).reshape(36)
                          for ph in self.phases])
        stiff_d = torch.as_tensor(np.tile(stiff, (self.P, 1))).to(dev)
        Cbar = torch.empty((self.S, 36), dtype=f64, device=dev)
        _lib.check(lib.drex_voigt_average_f64(_lib.ptr(self.o), _lib.ptr(self.f), _lib.ptr(stiff_d),
                                              _lib.ptr(self.vol_frac), self.S, self.N,
                                              _lib.ptr(Cbar), 0, _lib.stream_ptr(torch)))
        return Cbar.reshape(self.P, self.n_phases, 6, 6).sum(dim=1)

    @_on_own_device
    def resample(self, n_samples=None, seeds=None, phase_slot=0, lo=0, hi=None):
        """Volume-weighted resample of the current texture of particles [lo, hi) for one phase
        slot (stats.resample_orientations, stats.py:14-64, called once per particle with
        that particle's seed like the reference drivers do, tests/test_simple_shear_3d.py:
        194-202).  Returns device tensors ([n, n_samples, 3, 3], [n, n_samples])."""
        lib, torch = _lib.load(), self._torch
        hi = self.P if hi is None else hi
        n = hi - lo
        n_samples = self.N if n_samples is None else int(n_samples)
        if seeds is None:
            seeds = [None] * self.P
        uniforms = _stats.numpy_uniforms([seeds[i] for i in range(lo, hi)], n_samples)
        o_soa = self.o.reshape(self.P, self.n_phases, 9, self.N)[lo:hi, phase_slot].contiguous()
        f = self.f.reshape(self.P, self.n_phases, self.N)[lo:hi, phase_slot].contiguous()
        o_aos = torch.empty((n, self.N, 3, 3), dtype=torch.float64, device=self.device)
        out_o = torch.empty((n, n_samples, 3, 3), dtype=torch.float64, device=self.device)
        out_f = torch.empty((n, n_samples), dtype=torch.float64, device=self.device)
        st = _lib.stream_ptr(torch)
        _lib.check(lib.drex_unpack_soa_f64(_lib.ptr(o_soa), _lib.ptr(o_aos), n, self.N, st))
        nbytes = lib.drex_resample_workspace_bytes(n, self.N)
        ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=self.device)
        _lib.check(lib.drex_resample_orientations_f64(
            _lib.ptr(o_aos), _lib.ptr(f), _lib.ptr(uniforms), n, self.N, n_samples,
            _lib.ptr(out_o), _lib.ptr(out_f), _lib.ptr(ws), nbytes, st))
        return out_o, out_f

    @_on_own_device
    def misorientation_indices(self, system, phase_slot=0, batch=256, n_samples=None, seeds=None):
        """M-index of every particle's texture for one phase slot, [P] (device tensor): of
        the full texture, or -- with `n_samples` -- of the volume-weighted resample drawn
        with the per-particle `seeds`, which is what the reference's drivers feed to
        `misorientation_index`."""
        lib, torch = _lib.load(), self._torch
        a, b = system.value
        o = self.o.reshape(self.P, self.n_phases, 9, self.N)[:, phase_slot]
        out = torch.empty(self.P, dtype=torch.float64, device=self.device)
        st = _lib.stream_ptr(torch)
        for lo in range(0, self.P, batch):
            if n_samples is None:
                chunk, N = o[lo:lo + batch].contiguous(), self.N
            else:
                ro, _ = self.resample(n_samples, seeds, phase_slot, lo, min(lo + batch, self.P))
                N = int(ro.shape[1])
                chunk = torch.empty((ro.shape[0], 9, N), dtype=torch.float64, device=self.device)
                _lib.check(lib.drex_pack_soa_f64(_lib.ptr(ro), _lib.ptr(chunk), ro.shape[0], N, st))
            T = chunk.shape[0]
            nbytes = lib.drex_mindex_workspace_bytes(T, N, a, b)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            rc = lib.drex_mindex_f32(_lib.ptr(chunk), T, N, a, b, _lib.ptr(out[lo:lo + T]),
                                     None, _lib.ptr(ws), nbytes, st)
            _lib.check(rc)
        return out


def bind_host_to_gpu(device_index=None):
    """Pin this process to the CPU cores NVML reports as nearest to the CUDA device, so that
    pinned staging buffers allocated afterwards (`ParticleEnsemble.update_host`) live on the
    GPU's own NUMA node.  With one process per GPU on a multi-socket host the state otherwise
    crosses the socket interconnect twice per step.  Returns the cores chosen, or [] when NVML
    or the affinity call is unavailable (nothing is changed then)."""
    import os

    torch = _lib.require_cuda()
    try:
        import pynvml

        index = torch.cuda.current_device() if device_index is None else int(device_index)
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        near = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cores = sorted(near & os.sched_getaffinity(0))
        if cores:
            os.sched_setaffinity(0, cores)
        return cores
    except Exception:  # NVML missing, no permission, non-Linux: leave the affinity alone
        return []
