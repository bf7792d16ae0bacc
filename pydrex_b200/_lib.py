"""ctypes binding of libdrex_b200.so (C ABI declared in include/drex_b200.h).

The shared library is built in-tree (`pydrex_b200/lib/libdrex_b200.so`) by `build()` /
`__graft_entry__.build()` with nvcc for sm_100a.  There is NO CPU fallback: if the
library is missing, or no CUDA device is visible when a compute entry point is called,
the call raises.
"""

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.environ.get("PYDREX_B200_LIB", os.path.join(LIB_DIR, "libdrex_b200.so"))
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

# every symbol include/drex_b200.h declares
EXPORTS = (
    "drex_version", "drex_last_error", "drex_device_info",
    "drex_pack_soa_f64", "drex_unpack_soa_f64",
    "drex_fluidity_unpack_f64", "drex_fluidity_pack_f64",
    "drex_derivatives_workspace_bytes", "drex_derivatives_f64", "drex_grain_probe_f64",
    "drex_grain_helper_f64",
    "drex_integrate_workspace_bytes", "drex_integrate_chunk_grains", "drex_integrate_f64",
    "drex_integrate_next_order",
    "drex_voigt_average_f64", "drex_elasticity_components_f64",
    "drex_axis_scatter_f64",
    "drex_flow_field_f64",
    "drex_pathlines_f64",
    "drex_extract_vars_f64",
    "drex_apply_gbs_f64",
    "drex_misorientation_angles",
    "drex_strain_rate_max_f64",
    "drex_tensor_ops_f64",
    "drex_polar_decompose_f64",
    "drex_resample_workspace_bytes",
    "drex_resample_orientations_f64", "drex_numpy_uniform_f64",
    "drex_mindex_workspace_bytes", "drex_mindex_theory_host", "drex_mindex_f32",
    "drex_misorientations_random_host",
    "drex_fp64_peak_probe",
    "drex_grain_rhs_ceiling_probe",
)

DREX_OK = 0
DREX_E_INVALID = -1
DREX_E_UNSUPPORTED = -2
DREX_E_CUDA = -3
DREX_E_WORKSPACE = -4
DREX_E_NODEVICE = -5

L_CONSTANT, L_LINEAR, L_CHEBYSHEV, L_PIECEWISE = 0, 1, 2, 3
STATUS_MESSAGES = {
    1: "maximum number of internal steps exceeded",
    2: "step size underflow",
    3: "non-finite state or derivative (is the velocity gradient zero or NaN?)",
}


class LibraryNotBuilt(RuntimeError):
    """libdrex_b200.so is missing; run `python -c 'import __graft_entry__ as g; g.build()'`."""


class DrexParams(ctypes.Structure):
    _fields_ = [
        ("stress_exponent", ctypes.c_double),
        ("deformation_exponent", ctypes.c_double),
        ("nucleation_efficiency", ctypes.c_double),
        ("gbm_mobility", ctypes.c_double),
        ("gbs_threshold", ctypes.c_double),
    ]


class DrexTolerances(ctypes.Structure):
    _fields_ = [
        ("rtol", ctypes.c_double),
        ("atol_abs", ctypes.c_double),
        ("atol_rel", ctypes.c_double),
        ("atol_abs_F", ctypes.c_double),
        ("atol_abs_f", ctypes.c_double),
        ("first_step_frac", ctypes.c_double),
        ("max_step_frac", ctypes.c_double),
        ("max_steps", ctypes.c_int32),
        ("method", ctypes.c_int32),
    ]


class DrexFlowField(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("i0", ctypes.c_int32),
        ("i1", ctypes.c_int32),
        ("a", ctypes.c_double),
        ("b", ctypes.c_double),
    ]


FLOW_SIMPLE_SHEAR, FLOW_CELL, FLOW_CORNER = 0, 1, 2
PATH_MESSAGES = {
    1: "the velocity field is not finite on the pathline",
    2: "pathline tracer exceeded its step budget",
    3: "empty pathline: the final location is outside the domain or there is no strain budget",
    4: "pathline reached the time limit before accumulating the target strain",
}


def sources():
    return sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))
    ) + [os.path.join(INCLUDE, "drex_b200.h")]


def build(force=False, verbose=False):
    """Compile csrc/drex_capi.cu for sm_100a into lib/libdrex_b200.so (nvcc, no GPU needed)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in sources())
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    cmd = ["nvcc", *NVCC_FLAGS, "-o", LIB_PATH, os.path.join(CSRC, "drex_capi.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def _declare(lib):
    c = ctypes
    vp, i64, i32, dbl, szt = c.c_void_p, c.c_int64, c.c_int, c.c_double, c.c_size_t
    lib.drex_version.argtypes = [c.POINTER(c.c_int), c.POINTER(c.c_int)]
    lib.drex_last_error.argtypes = [c.c_char_p, szt]
    lib.drex_device_info.argtypes = [i32] + [c.POINTER(c.c_int)] * 3
    lib.drex_pack_soa_f64.argtypes = [vp, vp, i64, i64, vp]
    lib.drex_unpack_soa_f64.argtypes = [vp, vp, i64, i64, vp]
    lib.drex_fluidity_unpack_f64.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp]
    lib.drex_fluidity_pack_f64.argtypes = [vp, vp, vp, vp, vp, vp, dbl, i64, i64, vp, vp]
    lib.drex_derivatives_workspace_bytes.restype = szt
    lib.drex_derivatives_workspace_bytes.argtypes = [i64, i64]
    lib.drex_derivatives_f64.argtypes = (
        [vp] * 9 + [vp] * 3 + [c.POINTER(DrexParams), i64, i64, vp, vp, vp, szt, vp]
    )
    lib.drex_grain_probe_f64.argtypes = (
        [vp, vp, vp, i32, i32, c.POINTER(DrexParams), i64] + [vp] * 6 + [vp]
    )
    lib.drex_grain_helper_f64.argtypes = [i32, vp, vp, vp, vp, dbl, dbl, dbl, dbl, vp, vp]
    lib.drex_integrate_next_order.argtypes = [vp, vp, i64, vp, vp]
    lib.drex_integrate_chunk_grains.restype = i32
    lib.drex_integrate_chunk_grains.argtypes = []
    lib.drex_integrate_workspace_bytes.restype = szt
    lib.drex_integrate_workspace_bytes.argtypes = [i64, i64, i32, i32]
    lib.drex_integrate_f64.argtypes = (
        [i64, i64, i64] + [vp] * 5 + [vp] * 3 + [vp] * 6 + [vp, vp, i32, i32, vp]
        + [c.POINTER(DrexParams), c.POINTER(DrexTolerances)] + [vp] * 7 + [vp, szt, vp]
    )
    lib.drex_voigt_average_f64.argtypes = [vp, vp, vp, vp, i64, i64, vp, i32, vp]
    lib.drex_elasticity_components_f64.argtypes = [vp, i64, vp, vp]
    lib.drex_axis_scatter_f64.argtypes = [vp, i64, i64, i32, vp, vp]
    lib.drex_flow_field_f64.argtypes = [c.POINTER(DrexFlowField), vp, i64, vp, vp, vp]
    lib.drex_pathlines_f64.argtypes = (
        [c.POINTER(DrexFlowField), vp, i64, c.POINTER(c.c_double * 3), c.POINTER(c.c_double * 3)]
        + [dbl] * 4 + [i64, i32, i32, i32] + [vp] * 7 + [i32, vp, vp, vp]
    )
    lib.drex_extract_vars_f64.argtypes = [vp, i64, i64, vp, vp, vp, vp]
    lib.drex_apply_gbs_f64.argtypes = [vp, vp, vp, dbl, i64, i64, vp]
    lib.drex_misorientation_angles.argtypes = [vp, vp, i32, i64, i32, i32, vp, vp]
    lib.drex_strain_rate_max_f64.argtypes = [vp, i64, vp, vp]
    lib.drex_tensor_ops_f64.argtypes = [i32, vp, vp, i64, vp, vp]
    lib.drex_polar_decompose_f64.argtypes = [vp, i64, i32, vp, vp, vp]
    lib.drex_resample_workspace_bytes.restype = szt
    lib.drex_resample_workspace_bytes.argtypes = [i64, i64]
    lib.drex_resample_orientations_f64.argtypes = [vp, vp, vp, i64, i64, i64, vp, vp, vp, szt, vp]
    lib.drex_numpy_uniform_f64.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.drex_mindex_workspace_bytes.restype = szt
    lib.drex_mindex_workspace_bytes.argtypes = [i64, i64, i32, i32]
    lib.drex_mindex_theory_host.argtypes = [i32, i32, c.POINTER(c.c_double)]
    lib.drex_misorientations_random_host.argtypes = [dbl, dbl, i32, i32, c.POINTER(c.c_double)]
    lib.drex_mindex_f32.argtypes = [vp, i64, i64, i32, i32, vp, vp, vp, szt, vp]
    lib.drex_fp64_peak_probe.argtypes = [i32, c.POINTER(c.c_double)]
    lib.drex_grain_rhs_ceiling_probe.argtypes = [i32, c.POINTER(c.c_double)]
    return lib


def load():
    """Return the loaded library; raise `LibraryNotBuilt` if the .so is absent."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise LibraryNotBuilt(
                    f"{LIB_PATH} not found. pydrex_b200 has no CPU fallback; build the CUDA "
                    "library first: python -c 'import __graft_entry__ as g; g.build()'"
                )
            _lib = _declare(ctypes.CDLL(LIB_PATH))
    return _lib


def chunk_grains():
    """Grains per integration chunk (one step sequence, one carried step each)."""
    return int(load().drex_integrate_chunk_grains())


def last_error():
    buf = ctypes.create_string_buffer(512)
    load().drex_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc):
    """Translate a C-ABI return code into the exception the reference would raise."""
    if rc >= 0:
        return rc
    msg = last_error()
    if rc in (DREX_E_INVALID, DREX_E_UNSUPPORTED):
        raise ValueError(msg)
    if rc == DREX_E_WORKSPACE:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def require_cuda():
    """Fail loudly when there is no GPU: this package never computes on the CPU."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "pydrex_b200 needs a CUDA device (built for B200 / sm_100a); there is no CPU "
            "fallback. Use the reference PyDRex on machines without a GPU."
        )
    load()
    return torch


def stream_ptr(torch):
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array as c_void_p; None -> NULL."""
    if t is None:
        return ctypes.c_void_p(0)
    if hasattr(t, "data_ptr"):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)
