"""Core D-Rex functions and enums: the drop-in face of `pydrex.core` for the CPO hot path.

Same names, argument meaning and error behaviour as the reference
(/root/reference/src/pydrex/core.py); the arithmetic runs in the sm_100a kernels of
libdrex_b200.so (csrc/drex_grain.cuh, csrc/drex_derivatives.cuh).  numpy in => numpy out;
torch CUDA tensors in => torch CUDA tensors out.  No CPU fallback.
"""

import ctypes
from dataclasses import asdict, dataclass
from enum import IntEnum, unique

import numpy as np

from . import _lib


#: Levi-Civita symbol (core.py:25-31); the kernels use the equivalent cross products
PERMUTATION_SYMBOL = np.array(
    [
        [[0.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, -1.0, 0.0]],
        [[0.0, 0.0, -1.0], [0.0, 0.0, 0.0], [1.0, 0.0, 0.0]],
        [[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 0.0]],
    ]
)


@unique
class MineralPhase(IntEnum):
    """Supported mineral phases (core.py:35-47)."""

    olivine = 0
    enstatite = 1


@unique
class MineralFabric(IntEnum):
    """Supported mineral fabrics (core.py:50-67)."""

    olivine_A = 0
    olivine_B = 1
    olivine_C = 2
    olivine_D = 3
    olivine_E = 4
    enstatite_AB = 5


@unique
class DeformationRegime(IntEnum):
    """Deformation regimes (core.py:70-118)."""

    min_viscosity = 0
    matrix_diffusion = 1
    boundary_diffusion = 2
    sliding_diffusion = 3
    matrix_dislocation = 4
    sliding_dislocation = 5
    frictional_yielding = 6
    max_viscosity = 7


@dataclass(frozen=True)
class DefaultParams:
    """Default simulation parameters, the same record as the reference's (core.py:125-312) so that
    `as_dict()` carries every key a driver written against it may read.  The CPO hot path reads
    the first nine (minerals.py:400-445, 470); the flow-law constants below are data only (the
    viscosity module is out of scope)."""

    phase_assemblage: tuple = (MineralPhase.olivine,)
    phase_fractions: tuple = (1.0,)
    stress_exponent: float = 1.5
    deformation_exponent: float = 3.5
    gbm_mobility: int = 125
    gbs_threshold: float = 0.3
    nucleation_efficiency: float = 5.0
    number_of_grains: int = 3500
    initial_olivine_fabric: MineralFabric = MineralFabric.olivine_A
    disl_Peierls_stress: float = 2.0
    disl_prefactors: tuple = (1e-16, 1e-17)
    diff_prefactors: tuple = (1e-10, 1e-10)
    disl_lowtemp_switch: float = 0.7
    disl_activation_energy: float = 460.0
    disl_activation_volume: float = 12.0
    diff_activation_energies: tuple = (430.0, 330)
    diff_activation_volumes: tuple = (4.0, 4.0)
    disl_coefficients: tuple = (4.4e8, -5.26e4, 2.11e-2, 1.74e-4, -41.8, 4.21e-2, -1.14e-5)

    def __post_init__(self):
        for k, v in self.__dataclass_fields__.items():  # core.py:297-300
            if v.type is not type(v.default):
                raise ValueError(f"Illegal type for {self.__class__.__qualname__}.{k}")

    def as_dict(self):
        """Return a mutable copy of the parameters as a dictionary."""
        return asdict(self)


_CRSS = {
    MineralFabric.olivine_A: (1, 2, 3, np.inf),
    MineralFabric.olivine_B: (3, 2, 1, np.inf),
    MineralFabric.olivine_C: (3, 2, np.inf, 1),
    MineralFabric.olivine_D: (1, 1, 3, np.inf),
    MineralFabric.olivine_E: (3, 1, 2, np.inf),
}


def get_crss(phase, fabric):
    """Critical resolved shear stresses for `phase` and `fabric` (core.py:315-342)."""
    if phase == MineralPhase.olivine:
        try:
            return np.array(_CRSS[MineralFabric(fabric)], dtype=np.float64)
        except (KeyError, ValueError):
            raise ValueError(f"unsupported olivine fabric: {fabric}") from None
    elif phase == MineralPhase.enstatite:
        if fabric == MineralFabric.enstatite_AB:
            return np.array([np.inf, np.inf, np.inf, 1])
        raise ValueError(f"unsupported enstatite fabric: {fabric}")
    raise ValueError(f"phase must be a valid `MineralPhase`, not {phase}")


def _params_struct(stress_exponent, deformation_exponent, nucleation_efficiency, gbm_mobility,
                   gbs_threshold=0.0):
    return _lib.DrexParams(float(stress_exponent), float(deformation_exponent),
                           float(nucleation_efficiency), float(gbm_mobility),
                           float(gbs_threshold))


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def derivatives(regime, phase, fabric, n_grains, orientations, fractions, strain_rate,
                velocity_gradient, deformation_gradient_spin, stress_exponent,
                deformation_exponent, nucleation_efficiency, gbm_mobility, volume_fraction):
    """Derivatives of orientation and volume distribution (core.py:347-474).

    Returns `(orientations_diff[N,3,3], fractions_diff[N])`.  Inputs may be any numeric
    dtype (the reference tests pass int64 arrays); arithmetic is float64 on the GPU.
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    want_torch = _is_torch(orientations)
    dev = torch.device("cuda", torch.cuda.current_device())
    N = int(n_grains)

    def dev_f64(x, shape):
        t = x if _is_torch(x) else torch.as_tensor(np.asarray(x, dtype=np.float64))
        t = t.to(device=dev, dtype=torch.float64).reshape(shape).contiguous()
        return t

    o_aos = dev_f64(orientations, (1, N, 9))
    f = dev_f64(fractions, (1, N))
    D = dev_f64(strain_rate, (1, 9))
    L = dev_f64(velocity_gradient, (1, 9))
    spin = dev_f64(deformation_gradient_spin, (1, 9))
    meta_host = np.array([int(regime), int(phase), int(fabric)], dtype=np.int32)
    meta = torch.as_tensor(meta_host).to(dev)
    vf = torch.tensor([float(volume_fraction)], dtype=torch.float64, device=dev)
    params = _params_struct(stress_exponent, deformation_exponent, nucleation_efficiency,
                            gbm_mobility)
    o_soa = torch.empty((1, 9, N), dtype=torch.float64, device=dev)
    do_soa = torch.empty_like(o_soa)
    do_aos = torch.empty((N, 3, 3), dtype=torch.float64, device=dev)
    df = torch.empty(N, dtype=torch.float64, device=dev)
    ws_bytes = lib.drex_derivatives_workspace_bytes(1, N)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr(torch)
    P = _lib.ptr
    _lib.check(lib.drex_pack_soa_f64(P(o_aos), P(o_soa), 1, N, st))
    _lib.check(lib.drex_derivatives_f64(
        P(o_soa), P(f), P(D), P(L), P(spin), P(meta[0:1]), P(meta[1:2]), P(meta[2:3]), P(vf),
        P(meta_host[0:1]), P(meta_host[1:2]), P(meta_host[2:3]), ctypes.byref(params), 1, N,
        P(do_soa), P(df), P(ws), ws_bytes, st))
    _lib.check(lib.drex_unpack_soa_f64(P(do_soa), P(do_aos), 1, N, st))
    if want_torch:
        return do_aos, df
    return do_aos.cpu().numpy(), df.cpu().numpy()


def _helper(op, in0, in1, in2=None, idx=None, scalars=(0.0, 0.0, 0.0, 0.0), n_out=1):
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())

    def up(x):
        if x is None:
            return None
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64).ravel()).to(dev)

    t0, t1, t2 = up(in0), up(in1), up(in2)
    tidx = None
    if idx is not None:
        tidx = torch.as_tensor(np.ascontiguousarray(idx, dtype=np.int32)).to(dev)
    out = torch.empty(n_out, dtype=torch.float64, device=dev)
    s = [float(x) for x in scalars]
    _lib.check(lib.drex_grain_helper_f64(op, _lib.ptr(t0), _lib.ptr(t1), _lib.ptr(t2),
                                         _lib.ptr(tidx), s[0], s[1], s[2], s[3],
                                         _lib.ptr(out), _lib.stream_ptr(torch)))
    return out.cpu().numpy()


def _get_slip_invariants(strain_rate, orientation):
    """Strain-rate invariants of the four slip systems (core.py:571-598)."""
    return _helper(0, strain_rate, orientation, n_out=4)


def _get_slip_rates_olivine(invariants, slip_indices, crss, deformation_exponent):
    """Relative slip rates of the active olivine slip systems (core.py:541-568)."""
    return _helper(1, invariants, crss, idx=slip_indices,
                   scalars=(deformation_exponent, 0, 0, 0), n_out=4)


def _get_deformation_rate(phase, orientation, slip_rates):
    """Schmid / deformation-rate tensor (core.py:477-501)."""
    return _helper(2, orientation, slip_rates, n_out=9).reshape(3, 3)


def _get_slip_rate_softest(deformation_rate, velocity_gradient):
    """Dimensionless strain rate on the softest slip system (core.py:504-538)."""
    return float(_helper(3, deformation_rate, velocity_gradient, n_out=1)[0])


def _get_orientation_change(orientation, velocity_gradient, deformation_rate, slip_rate_softest):
    """Rotation rate of a grain undergoing dislocation creep (core.py:601-642)."""
    return _helper(4, orientation, velocity_gradient, deformation_rate,
                   scalars=(slip_rate_softest, 0, 0, 0), n_out=9).reshape(3, 3)


def _get_strain_energy(crss, slip_rates, slip_indices, slip_rate_softest, stress_exponent,
                       deformation_exponent, nucleation_efficiency):
    """Strain energy due to dislocations (core.py:645-684); `slip_indices` is unused there too."""
    return float(_helper(5, crss, slip_rates,
                         scalars=(slip_rate_softest, stress_exponent, deformation_exponent,
                                  nucleation_efficiency), n_out=1)[0])


def _get_rotation_and_strain(phase, fabric, orientation, strain_rate, velocity_gradient,
                             stress_exponent, deformation_exponent, nucleation_efficiency):
    """Rotation rate and strain energy of one grain (core.py:687-760)."""
    pr = grain_probe(phase, fabric, np.asarray(orientation, dtype=np.float64)[None], strain_rate,
                     velocity_gradient, stress_exponent, deformation_exponent,
                     nucleation_efficiency)
    return pr["orientation_change"][0], float(pr["strain_energy"][0])


def grain_probe(phase, fabric, orientations, strain_rate, velocity_gradient, stress_exponent,
                deformation_exponent, nucleation_efficiency):
    """All per-grain intermediates of core._get_rotation_and_strain for G grains at once."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    o = torch.as_tensor(np.ascontiguousarray(orientations, dtype=np.float64)).to(dev)
    G = o.shape[0]
    D = torch.as_tensor(np.ascontiguousarray(strain_rate, dtype=np.float64)).to(dev)
    L = torch.as_tensor(np.ascontiguousarray(velocity_gradient, dtype=np.float64)).to(dev)
    params = _params_struct(stress_exponent, deformation_exponent, nucleation_efficiency, 0.0)
    mk = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)  # noqa: E731
    inv, rates, defr, soft, energy, da = mk(G, 4), mk(G, 4), mk(G, 3, 3), mk(G), mk(G), mk(G, 3, 3)
    P = _lib.ptr
    _lib.check(lib.drex_grain_probe_f64(P(o), P(D), P(L), int(phase), int(fabric),
                                        ctypes.byref(params), G, P(inv), P(rates), P(defr),
                                        P(soft), P(energy), P(da), _lib.stream_ptr(torch)))
    return {"invariants": inv.cpu().numpy(), "slip_rates": rates.cpu().numpy(),
            "deformation_rate": defr.cpu().numpy(), "slip_rate_softest": soft.cpu().numpy(),
            "strain_energy": energy.cpu().numpy(), "orientation_change": da.cpu().numpy()}
