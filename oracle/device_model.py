"""Numpy model of the DEVICE integrator's algorithm.  TEST INFRASTRUCTURE ONLY.

`pydrex_b200/csrc/drex_integrate.cuh` does not restate LSODA; it solves the same ODE
(/root/reference/src/pydrex/minerals.py:388-453) in a form that has no coupling between
grains inside an outer interval.  This file restates THAT algorithm in numpy on top of the
oracle's per-grain physics (drex_oracle.c: oracle_grain_batch), so the algorithm itself can
be checked against the LSODA oracle and the reference's golden vectors on the CPU, and so
the CUDA kernel has a line-by-line model to be compared with.  Only tests/ import it.

The algorithm
-------------
* The rotation rate and strain energy E_g of a grain depend on that grain and L(t) only
  (core.py:687-760).  The volume fractions obey (core.py:431-436, minerals.py:450-451)

      df_g/dt = s(t) vf M* fhat_g (Ebar - E_g),  fhat = clip(f)/sum clip(f),  Ebar = sum fhat E,

  a replicator equation: sum f is conserved, f stays >= 0, and its exact solution is

      fhat_g(t) = fhat_g(t0) exp(v_g(t)) / sum_j fhat_j(t0) exp(v_j(t)),   dv_g/dt = -s vf M* E_g.

  So each grain carries ten unknowns (nine direction cosines and v_g) that do not see any
  other grain until the end of the interval, where one normalisation recovers the fractions.
* Grains are integrated in chunks of 32 (one warp; `chunk`): every chunk walks from t0 to t1 with its
  OWN Bogacki-Shampine 3(2) steps (FSAL), controlled by the weighted max norm of ITS grains
  with the reference's weights (minerals.py:523-524): atol + rtol |y| per component, for the
  fraction components through f_g ~ f_g(t0) exp(v_g), err_f = f_g err_v.
* The deformation gradient (dF/dt = L F, minerals.py:426) is integrated on its own.
* End of the interval as minerals.py:464-474, 536-538: extract_vars, apply_gbs against the
  interval-start orientations, extract_vars.
"""

import numpy as np

from . import oracle

BS_A21, BS_A32 = 0.5, 0.75
BS_B = (2.0 / 9.0, 1.0 / 3.0, 4.0 / 9.0)
BS_E = (-5.0 / 72.0, 1.0 / 12.0, 1.0 / 9.0, -1.0 / 8.0)
SAFETY, STRETCH = 0.9, 1.10
CHUNK = 32


def plan_step(h_prop, remaining):
    """drex_integrate.cuh plan_step: the rest of the interval in equal steps."""
    ratio = np.abs(remaining) / (np.abs(h_prop) * STRETCH)
    n = np.ceil(np.maximum(ratio, 1.0))
    return np.where(ratio > 1.0, remaining / n, remaining)


def _grain_rhs(a, Ls, phase, fabric, p, n, lam):
    G = a.shape[0]
    a = np.ascontiguousarray(np.clip(a, -1.0, 1.0))  # utils.py:187
    Ls = np.ascontiguousarray(Ls)
    da, E = np.empty((G, 9)), np.empty(G)
    dp = oracle._dp
    rc = oracle.lib().oracle_grain_batch(
        int(phase), int(fabric), G, a.ctypes.data_as(dp), Ls.ctypes.data_as(dp), p, n, lam,
        da.ctypes.data_as(dp), E.ctypes.data_as(dp))
    if rc != 0:
        raise ValueError("unsupported phase/fabric")
    return da, E


def integrate_F(F0, get_L, t0, t1, rtol, atol_abs, atol_rel):
    """dF/dt = L(t) F with its own BS3(2) steps (one thread per system on the device)."""
    F, t, h = np.array(F0, dtype=float).reshape(3, 3), t0, t1 - t0
    k1 = get_L(t) @ F
    relw = atol_rel + rtol
    for _ in range(1000000):
        k2 = get_L(t + BS_A21 * h) @ (F + h * BS_A21 * k1)
        k3 = get_L(t + BS_A32 * h) @ (F + h * BS_A32 * k2)
        Fn = F + h * (BS_B[0] * k1 + BS_B[1] * k2 + BS_B[2] * k3)
        k4 = get_L(t + h) @ Fn
        err = h * (BS_E[0] * k1 + BS_E[1] * k2 + BS_E[2] * k3 + BS_E[3] * k4)
        emax = np.max(np.abs(err) / (atol_abs + relw * np.abs(Fn)))
        fac = min(5.0, max(0.2, SAFETY * emax ** (-1.0 / 3.0))) if emax > 0 else 5.0
        if emax <= 1.0:
            t, F, k1 = t + h, Fn, k4
            rem = t1 - t
            if abs(rem) <= 1e-13 * max(abs(t1), abs(t1 - t0)):
                return F
            h = float(plan_step(h * fac, rem))
        else:
            h *= min(1.0, fac)
    raise RuntimeError("integrate_F: too many steps")


def update_orientations(orientations, fractions, regime, phase, fabric, params,
                        deformation_gradient, get_velocity_gradient, pathline,
                        rtol=1e-6, atol_abs=1e-4, atol_rel=1e-6, max_steps=100000, stats=None,
                        h_carry=None, global_steps=False, chunk=CHUNK):
    """The device algorithm for ONE system (regimes 4 and 6).  Returns (F, o, f) like
    oracle.update_orientations."""
    assert regime in (4, 6)
    t0, t1, get_position = pathline
    N = len(fractions)
    o_prev = np.asarray(orientations, dtype=float).reshape(N, 9)
    f0 = np.maximum(np.asarray(fractions, dtype=float), 0.0)
    vf = params["phase_fractions"][list(params["phase_assemblage"]).index(phase)]
    p, n = float(params["stress_exponent"]), float(params["deformation_exponent"])
    lam, M = float(params["nucleation_efficiency"]), float(params["gbm_mobility"])
    rs = 0.3 if regime == 6 else 1.0  # core.py:459,464
    kappa = rs * vf * M if phase == 0 else 0.0  # enstatite strain energy is identically 0
    relw = atol_rel + rtol

    def get_L(t):
        return np.asarray(get_velocity_gradient(t, get_position(t)), dtype=float)

    n_chunks = (N + chunk - 1) // chunk  # the kernel's drex_integrate_chunk_grains()
    chunk_of = np.arange(N) // chunk

    def flows(tc):  # per-chunk times -> per-grain (L/s, s)
        Ls, s = np.empty((n_chunks, 9)), np.empty(n_chunks)
        for c in range(n_chunks):
            L = get_L(tc[c])
            s[c] = oracle.lib().oracle_strain_rate_max(
                np.ascontiguousarray(L).ctypes.data_as(oracle._dp))
            Ls[c] = (L / s[c]).ravel()
        return Ls[chunk_of], s[chunk_of]

    def rhs(a, tc):
        Ls, s = flows(tc)
        da, E = _grain_rhs(a, Ls, phase, fabric, p, n, lam)
        return da * (rs * s)[:, None], -kappa * s * E  # minerals.py:450-451

    a, v = o_prev.copy(), np.zeros(N)
    span = t1 - t0
    t = np.full(n_chunks, float(t0))
    h = np.full(n_chunks, float(span))
    def scale_at(tt):
        sv = oracle.lib().oracle_strain_rate_max(np.ascontiguousarray(get_L(tt)).ctypes.data_as(oracle._dp))
        return sv if (sv > 0 and np.isfinite(sv)) else 1.0

    if h_carry is not None and len(h_carry) == n_chunks:
        # a chunk's last proposal of the previous interval (the kernel's per-chunk h_io),
        # carried as a STRAIN increment: time step x strain-rate scale
        h_carry = np.asarray(h_carry) / scale_at(t0)
        ok = np.abs(h_carry) > 0
        hc = np.where(np.abs(h_carry) < abs(span), np.copysign(np.abs(h_carry), span), span)
        h = np.where(ok, plan_step(hc, np.full(n_chunks, float(span))), h)
    h_out = h.copy()
    done = np.zeros(n_chunks, dtype=bool)
    k1, kv1 = rhs(a, t)
    n_rhs = np.ones(n_chunks, dtype=np.int64)
    n_acc = np.zeros(n_chunks, dtype=np.int64)
    n_rej = np.zeros(n_chunks, dtype=np.int64)
    last_rej = np.zeros(n_chunks, dtype=bool)
    for _ in range(max_steps):
        if done.all():
            break
        hg = h[chunk_of][:, None]
        k2, kv2 = rhs(a + hg * BS_A21 * k1, t + BS_A21 * h)
        k3, kv3 = rhs(a + hg * BS_A32 * k2, t + BS_A32 * h)
        an = a + hg * (BS_B[0] * k1 + BS_B[1] * k2 + BS_B[2] * k3)
        vn = v + hg[:, 0] * (BS_B[0] * kv1 + BS_B[1] * kv2 + BS_B[2] * kv3)
        k4, kv4 = rhs(an, t + h)
        err = hg * (BS_E[0] * k1 + BS_E[1] * k2 + BS_E[2] * k3 + BS_E[3] * k4)
        errv = hg[:, 0] * (BS_E[0] * kv1 + BS_E[1] * kv2 + BS_E[2] * kv3 + BS_E[3] * kv4)
        ratio = np.max(np.abs(err) / (atol_abs + relw * np.abs(an)), axis=1)
        f_est = f0 * np.exp(vn)
        ratio = np.maximum(ratio, f_est * np.abs(errv) / (atol_abs + relw * f_est))
        emax = np.zeros(n_chunks)
        np.maximum.at(emax, chunk_of, ratio)
        if global_steps:  # one step sequence per system (what the first device integrator did)
            emax[:] = emax.max()
        if not np.all(np.isfinite(emax)):
            raise RuntimeError("non-finite state")
        with np.errstate(divide="ignore"):
            fac = np.where(emax > 0, SAFETY * emax ** (-1.0 / 3.0), 5.0)
        accept = (emax <= 1.0) & ~done
        reject = (emax > 1.0) & ~done
        fac = np.minimum(np.where(accept, np.where(last_rej, 1.0, 5.0), 1.0), np.maximum(0.2, fac))
        n_rhs += np.where(done, 0, 3)
        n_acc += accept
        n_rej += reject
        ga = accept[chunk_of]
        a[ga], v[ga], k1[ga], kv1[ga] = an[ga], vn[ga], k4[ga], kv4[ga]
        t = np.where(accept, t + h, t)
        rem = t1 - t
        fin = accept & (np.abs(rem) <= 1e-13 * max(abs(t1), abs(span)))
        done |= fin
        h_next = h * fac
        h_acc = np.where(np.abs(h_next) >= np.abs(rem) * (1 - 1e-10), rem, plan_step(h_next, np.where(rem == 0, span, rem)))
        h_out = np.where(accept | reject, h_next, h_out)
        h = np.where(done, h, np.where(accept, h_acc, h_next))
        last_rej = np.where(accept, False, np.where(reject, True, last_rej))
    else:
        raise RuntimeError("max_steps")
    if stats is not None:
        stats.update(n_rhs=n_rhs, n_acc=n_acc, n_rej=n_rej, h=h_out * scale_at(t1))

    # end of the interval: minerals.py:464-474, 536-538 (utils.py:159-190)
    u = f0 * np.exp(v)
    f = u / u.sum()
    o = np.clip(a, -1.0, 1.0).reshape(N, 3, 3)
    o, f = oracle.apply_gbs(o, f, params["gbs_threshold"], o_prev.reshape(N, 3, 3), N)
    f = np.maximum(f, 0.0)
    f /= f.sum()
    F = integrate_F(deformation_gradient, get_L, t0, t1, rtol, atol_abs, atol_rel)
    return F, np.clip(o, -1.0, 1.0), f
