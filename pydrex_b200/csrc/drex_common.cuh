// drex_common.cuh -- small device helpers shared by the kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace drex {

// ---- deterministic block reductions (fixed shuffle tree, fixed warp order) ----------

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum of two values over the block; result broadcast to all threads.
// `red` is shared scratch of at least 2 * 32 doubles.  Contains two __syncthreads.
template <int BLOCK>
__device__ __forceinline__ void block_sum2(double &a, double &b, double *red) {
    constexpr int NW = BLOCK / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();  // protects `red` against the previous reduction's readers
    if (lane == 0) {
        red[warp] = a;
        red[32 + warp] = b;
    }
    __syncthreads();
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {  // every thread adds in the same order
        sa += red[w];
        sb += red[32 + w];
    }
    a = sa;
    b = sb;
}

template <int BLOCK>
__device__ __forceinline__ double block_max(double v, double *red) {
    constexpr int NW = BLOCK / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // NaN must not be lost: fmax drops NaN, so carry it as +inf
    v = (v != v) ? INFINITY : v;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double m = red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = fmax(m, red[w]);
    return m;
}

// ---- 3x3 helpers -------------------------------------------------------------------

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3 (row-major).  A = V diag(w) V^T.
__device__ inline void jacobi_eig3(const double *A_in, double *w, double *V) {
    double A[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        A[i] = A_in[i];
        V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
        const double diag = fabs(A[0]) + fabs(A[4]) + fabs(A[8]);
        if (off <= 1e-300 || off <= 1e-22 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double apq = A[3 * p + q];
                if (apq == 0.0) continue;
                const double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[3 * k + p], akq = A[3 * k + q];
                    A[3 * k + p] = c * akp - s * akq;
                    A[3 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[3 * p + k], aqk = A[3 * q + k];
                    A[3 * p + k] = c * apk - s * aqk;
                    A[3 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[3 * k + p], vkq = V[3 * k + q];
                    V[3 * k + p] = c * vkp - s * vkq;
                    V[3 * k + q] = s * vkp + c * vkq;
                }
            }
    }
    w[0] = A[0];
    w[1] = A[4];
    w[2] = A[8];
}

// max |eigenvalue| of sym(L)  (scipy.linalg.eigvalsh call at minerals.py:421-422).
// Closed form for a symmetric 3x3 (trigonometric solution of the characteristic cubic):
// the extreme eigenvalues are q + 2p cos(phi) and q + 2p cos(phi + 2pi/3); absolute error
// ~1 ulp of the spectral radius, the same class as LAPACK's.  This sits on the serial
// per-step path of every CTA, so it must not iterate.
__device__ inline double strain_rate_max(const double *L) {
    const double d00 = L[0], d11 = L[4], d22 = L[8];
    const double d01 = 0.5 * (L[1] + L[3]), d02 = 0.5 * (L[2] + L[6]), d12 = 0.5 * (L[5] + L[7]);
    const double p1 = d01 * d01 + d02 * d02 + d12 * d12;
    const double q = (d00 + d11 + d22) / 3.0;
    const double b00 = d00 - q, b11 = d11 - q, b22 = d22 - q;
    const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
    const double p = sqrt(p2 / 6.0);
    if (!(p > 0.0)) return fabs(q);  // multiple of the identity (or NaN, which propagates)
    const double ip = 1.0 / p;
    const double c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip;
    const double c01 = d01 * ip, c02 = d02 * ip, c12 = d12 * ip;
    double r = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) +
                      c02 * (c01 * c12 - c11 * c02));
    r = fmin(1.0, fmax(-1.0, r));
    const double cphi = cos(acos(r) / 3.0);  // phi in [0, pi/3]
    const double sphi = sqrt(fmax(0.0, 1.0 - cphi * cphi));
    const double e_hi = q + 2.0 * p * cphi;
    // cos(phi + 2 pi/3) = -cos(phi)/2 - sqrt(3)/2 sin(phi)
    const double e_lo = q - p * (cphi + 1.7320508075688772935 * sphi);
    return fmax(fabs(e_hi), fabs(e_lo));
}

// Left stretch of M (tensors.polar_decompose(M)[1], tensors.py:14-19) = sqrt(M M^T)
__device__ inline void left_stretch(const double *M, double *Vout) {
    double G[9], w[3], U[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            G[3 * i + j] = M[3 * i] * M[3 * j] + M[3 * i + 1] * M[3 * j + 1] + M[3 * i + 2] * M[3 * j + 2];
    jacobi_eig3(G, w, U);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += U[3 * i + k] * sqrt(fmax(w[k], 0.0)) * U[3 * j + k];
            Vout[3 * i + j] = s;
        }
}

__device__ __forceinline__ void matmul3(const double *A, const double *B, double *C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

__device__ __forceinline__ double clip1(double x) { return fmin(1.0, fmax(-1.0, x)); }

}  // namespace drex
