"""State-vector helpers by name (drop-in for the parts of pydrex.utils on the CPO path,
reference utils.py:144-190, 365-371).  The integrator applies the same operations inside its
kernel; these entry points exist so that code and tests written against the reference's
names keep working.  numpy in, numpy out; torch CUDA tensors in, torch CUDA tensors out."""

import numpy as np

from . import _lib


def _dev(torch, x, dtype=None):
    dev = torch.device("cuda", torch.cuda.current_device())
    if not type(x).__module__.startswith("torch"):
        x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64))
    return x.to(device=dev, dtype=dtype or torch.float64).contiguous()


def extract_vars(y, n_grains):
    """Extract deformation gradient, orientation matrices and grain sizes from y
    (utils.py:183-190); y may carry a leading batch axis."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    is_torch = type(y).__module__.startswith("torch")
    yd = _dev(torch, y)
    single = yd.dim() == 1
    N = int(n_grains)
    yd = yd.reshape(-1, 9 + 10 * N)
    S = int(yd.shape[0])
    F = torch.empty((S, 3, 3), dtype=torch.float64, device=yd.device)
    o = torch.empty((S, N, 3, 3), dtype=torch.float64, device=yd.device)
    f = torch.empty((S, N), dtype=torch.float64, device=yd.device)
    _lib.check(lib.drex_extract_vars_f64(_lib.ptr(yd), S, N, _lib.ptr(F), _lib.ptr(o), _lib.ptr(f),
                                         _lib.stream_ptr(torch)))
    out = (F[0], o[0], f[0]) if single else (F, o, f)
    return out if is_torch else tuple(t.cpu().numpy() for t in out)


def apply_gbs(orientations, fractions, gbs_threshold, orientations_prev, n_grains):
    """Apply grain boundary sliding for small grains (utils.py:159-180).  Returns the new
    (orientations, fractions); numpy inputs are also updated in place like the reference."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    is_torch = type(orientations).__module__.startswith("torch")
    N = int(n_grains)
    o = _dev(torch, orientations).reshape(-1, N, 9).clone()
    f = _dev(torch, fractions).reshape(-1, N).clone()
    o_prev = _dev(torch, orientations_prev).reshape(-1, N, 9)
    S = int(f.shape[0])
    _lib.check(lib.drex_apply_gbs_f64(_lib.ptr(o), _lib.ptr(f), _lib.ptr(o_prev),
                                      float(gbs_threshold), S, N, _lib.stream_ptr(torch)))
    o = o.reshape(tuple(orientations.shape))
    f = f.reshape(tuple(fractions.shape))
    if is_torch:
        orientations.copy_(o)
        fractions.copy_(f)
        return orientations, fractions
    orientations[...] = o.cpu().numpy()
    fractions[...] = f.cpu().numpy()
    return orientations, fractions


def strain_increment(dt, velocity_gradient):
    """Tensorial strain increment |dt| max|eig((L + L^T) / 2)| (utils.py:144-156); `dt` and
    `velocity_gradient` may carry a leading batch axis."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    is_torch = type(velocity_gradient).__module__.startswith("torch")
    L = _dev(torch, velocity_gradient)
    single = L.dim() == 2
    L = L.reshape(-1, 9)
    out = torch.empty((L.shape[0],), dtype=torch.float64, device=L.device)
    _lib.check(lib.drex_strain_rate_max_f64(_lib.ptr(L), int(L.shape[0]), _lib.ptr(out),
                                            _lib.stream_ptr(torch)))
    if is_torch:
        res = torch.abs(torch.as_tensor(dt, device=out.device)) * out
        return res[0] if single else res
    res = np.abs(dt) * out.cpu().numpy()
    return float(res[0]) if single else res


def quat_product(q1, q2):
    """Quaternion product in scalar-last (x, y, z, w) format exactly as the reference writes
    it (utils.py:365-371): the cross-product term is cross(q1, q1) = 0, i.e. missing.  The
    M-index kernels build their symmetry-operator tables with the same expression; this
    four-number host helper only exists for code that calls it by name."""
    q1, q2 = np.asarray(q1), np.asarray(q2)
    return [
        *(q1[-1] * q2[:3] + q2[-1] * q1[:3] + np.cross(q1[:3], q1[:3])),
        q1[-1] * q2[-1] - np.dot(q1[:3], q2[:3]),
    ]
