"""Fluidity coupling: advance a whole cloud of CPO particle attributes in one call.

The reference couples to Fluidity through a per-particle Python hook
(examples/fluidity/corner2d/corner2d.flml:144-175): every model time step, for every
particle, it unpacks the `CPO_` attribute array of `10 N + 22` doubles into a fresh
`Mineral`, calls `update_orientations` with a velocity gradient that is LINEAR in time
between the previous and the current model step (corner2d.py:22-41), and packs the result.
`advance_cpo` does the same for all particles at once: one host->device copy of the
attribute rows, a layout kernel, one launch of the integrator with the `DREX_L_LINEAR`
provider, a layout kernel, one copy back.
"""

import ctypes

import numpy as np

from . import _lib
from .ensemble import ParticleEnsemble


def cpo_row_length(n_grains):
    return 10 * int(n_grains) + 22


def advance_cpo(cpo_prev, position, velocity_gradient, t, dt, params, phase=0, fabric=0, regime=4,
                ensemble=None, **tolerances):
    """New `CPO_` rows [P, 10 N + 22] for all particles (numpy in, numpy out).

    - `cpo_prev` — previous attribute rows [P, 10 N + 22]
    - `position` [P, 3], `velocity_gradient` [P, 3, 3] — current model-step values
    - `t`, `dt` — current time and step: the CPO is integrated from `t - dt` to `t`
    - `params` — dict with the hot-path keys of `DefaultParams.as_dict()`
    - `ensemble` — optional `ParticleEnsemble` (P particles, one phase) to reuse device
      buffers between model steps
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    cpo_prev = np.ascontiguousarray(cpo_prev, dtype=np.float64)
    P, row = cpo_prev.shape
    N = (row - 22) // 10
    if cpo_row_length(N) != row:
        raise ValueError(f"attribute rows of length {row} are not 10 N + 22")
    prm = dict(params)
    prm.setdefault("phase_assemblage", (phase,))
    prm.setdefault("phase_fractions", (1.0,))
    if ensemble is None:
        ensemble = ParticleEnsemble(P, N, prm, phases=(phase,), fabrics=(fabric,), regimes=(regime,),
                                    fractions=np.full((P, 1, N), 1.0 / N),
                                    orientations=np.zeros((P, 1, N, 3, 3)))  # overwritten below
    ens = ensemble
    if (ens.P, ens.N, ens.n_phases) != (P, N, 1):
        raise ValueError("ensemble shape does not match the attribute rows")
    dev, f64 = ens.device, torch.float64
    st = _lib.stream_ptr(torch)
    Pp = _lib.ptr
    d_cpo = torch.as_tensor(cpo_prev).to(dev)
    L_prev = torch.empty((P, 9), dtype=f64, device=dev)
    strain = torch.empty(P, dtype=f64, device=dev)
    _lib.check(lib.drex_fluidity_unpack_f64(Pp(d_cpo), P, N, Pp(ens.o), Pp(ens.f), Pp(ens.F),
                                            Pp(L_prev), Pp(strain), st))
    L_new = torch.as_tensor(np.ascontiguousarray(velocity_gradient, dtype=np.float64).reshape(P, 9)).to(dev)
    pos = torch.as_tensor(np.ascontiguousarray(position, dtype=np.float64).reshape(P, 3)).to(dev)
    nodes = torch.stack([L_prev, L_new], dim=1).reshape(P, 2, 3, 3)  # linear in time
    tol = ens.tolerances(**tolerances) if tolerances else ens.tolerances()
    ens.h.zero_()
    ens.update(float(t) - float(dt), float(t), nodes, kind=_lib.L_LINEAR, tol=tol)
    _lib.check(lib.drex_fluidity_pack_f64(Pp(ens.o), Pp(ens.f), Pp(ens.F), Pp(strain), Pp(pos),
                                          Pp(L_new), float(dt), P, N, Pp(d_cpo), st))
    return d_cpo.cpu().numpy()
