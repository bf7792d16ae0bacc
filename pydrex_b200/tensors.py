"""Tensor helpers (drop-in for pydrex.tensors: the Voigt-average path, reference
tensors.py:14-21, 167-203, 254-273, and the Browaeys & Chevrot decomposition helpers that
`diagnostics.elasticity_components` is built from, tensors.py:39-164, 206-251).

Every function takes one item like the reference or a leading batch axis; the arithmetic
runs in the library's kernels (`drex_tensor_ops_f64`, `drex_polar_decompose_f64`).
numpy in, numpy out; torch CUDA tensors in, torch CUDA tensors out.
"""

import numpy as np

from . import _lib


def _run(op, in0, item_shape, out_shape, in1=None, in1_shape=None):
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    is_torch = type(in0).__module__.startswith("torch")

    def to_dev(x):
        if not type(x).__module__.startswith("torch"):
            x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64))
        return x.to(device=dev, dtype=torch.float64).contiguous()

    a = to_dev(in0)
    single = a.dim() == len(item_shape)
    a = a.reshape((-1,) + tuple(item_shape))
    B = int(a.shape[0])
    b = None
    if in1 is not None:
        b = to_dev(in1).reshape((-1,) + tuple(in1_shape))
        if b.shape[0] == 1 and B > 1:
            b = b.expand(B, *in1_shape).contiguous()
        if int(b.shape[0]) != B:
            raise ValueError("batch sizes of the two arguments differ")
    out = torch.empty((B,) + tuple(out_shape), dtype=torch.float64, device=dev)
    _lib.check(lib.drex_tensor_ops_f64(op, _lib.ptr(a), None if b is None else _lib.ptr(b), B,
                                       _lib.ptr(out), _lib.stream_ptr(torch)))
    out = out[0] if single else out
    return out if is_torch else out.cpu().numpy()


def rotate(tensor, rotation):
    """Rotate 4-th order tensor(s) [..., 3, 3, 3, 3] using 3x3 rotation matrices
    (tensors.py:254-273)."""
    return _run(0, tensor, (3, 3, 3, 3), (3, 3, 3, 3), rotation, (3, 3))


def voigt_to_elastic_tensor(matrix):
    """4-th order elastic tensor from the equivalent 6x6 Voigt matrix (tensors.py:167-184)."""
    return _run(1, matrix, (6, 6), (3, 3, 3, 3))


def elastic_tensor_to_voigt(tensor):
    """6x6 Voigt matrix from the equivalent 4-th order elastic tensor (tensors.py:187-203)."""
    return _run(2, tensor, (3, 3, 3, 3), (6, 6))


def polar_decompose(matrix, left=True):
    """Polar decomposition M = V R (left, returns (R, V)) or M = R U (right, returns (R, U))
    of nonsingular 3x3 matrices (tensors.py:14-21)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    is_torch = type(matrix).__module__.startswith("torch")
    m = matrix if is_torch else torch.as_tensor(np.ascontiguousarray(matrix, dtype=np.float64))
    m = m.to(device=dev, dtype=torch.float64).contiguous()
    single = m.dim() == 2
    m = m.reshape(-1, 3, 3)
    rot, stretch = torch.empty_like(m), torch.empty_like(m)
    _lib.check(lib.drex_polar_decompose_f64(_lib.ptr(m), int(m.shape[0]), int(bool(left)),
                                            _lib.ptr(rot), _lib.ptr(stretch),
                                            _lib.stream_ptr(torch)))
    if single:
        rot, stretch = rot[0], stretch[0]
    if is_torch:
        return rot, stretch
    return rot.cpu().numpy(), stretch.cpu().numpy()


def voigt_decompose(matrix):
    """The two independent contractions of the elastic tensor given as 6x6 Voigt matrix:
    dilatational d_ij = C_ijkk and deviatoric v_ij = C_ijkj stiffness (tensors.py:39-73)."""
    out = _run(3, matrix, (6, 6), (2, 3, 3))
    return (out[0], out[1]) if out.ndim == 3 else (out[:, 0], out[:, 1])


def mono_project(voigt_vector):
    """Project the 21-component Voigt vector onto monoclinic symmetry (tensors.py:76-89)."""
    return _run(4, voigt_vector, (21,), (21,))


def ortho_project(voigt_vector):
    """Project the 21-component Voigt vector onto orthorhombic symmetry (tensors.py:92-103)."""
    return _run(5, voigt_vector, (21,), (21,))


def tetr_project(voigt_vector):
    """Project the 21-component Voigt vector onto tetragonal symmetry (tensors.py:106-119)."""
    return _run(6, voigt_vector, (21,), (21,))


def hex_project(voigt_vector):
    """Project the 21-component Voigt vector onto hexagonal symmetry (tensors.py:122-140)."""
    return _run(7, voigt_vector, (21,), (21,))


def upper_tri_to_symmetric(arr):
    """Symmetric array from the upper triangle of a 3x3 or 6x6 array (or a batch of them)
    (tensors.py:143-164)."""
    n = int(np.shape(arr)[-1])
    if n not in (3, 6) or np.shape(arr)[-2] != n:
        raise ValueError("upper_tri_to_symmetric handles 3x3 and 6x6 arrays on the device")
    return _run(10 if n == 3 else 11, arr, (n, n), (n, n))


def voigt_matrix_to_vector(matrix):
    """The 21-component Voigt vector equivalent to the 6x6 Voigt matrix (tensors.py:206-219)."""
    return _run(8, matrix, (6, 6), (21,))


def voigt_vector_to_matrix(vector):
    """The 6x6 matrix representation of the 21-component Voigt vector (tensors.py:222-251)."""
    return _run(9, vector, (21,), (6, 6))
