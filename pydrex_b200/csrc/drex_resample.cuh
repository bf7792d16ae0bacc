// drex_resample.cuh -- volume-weighted resampling of textures (sm_100a).
//
// Replaces stats.resample_orientations (reference src/pydrex/stats.py:14-64): per texture,
// sort grains by volume fraction (ascending), take the running sum, force its last entry to
// 1, and for every uniform variate u pick the first grain whose cumulative fraction is
// >= u (numpy.searchsorted, side="left").  Bit parity with the reference needs
//   * the SAME order of equal fractions as the reference's argsort.  numpy's default
//     argsort is an unstable introsort / SIMD sort whose tie order depends on the numpy
//     build and the host CPU, so the reference itself is only defined up to the order of
//     ties; this kernel sorts by (fraction, grain index), i.e. numpy.argsort(kind="stable").
//     Wherever fractions are distinct the two agree exactly, and the returned FRACTIONS agree
//     exactly in every case (tied grains have the same fraction);
//   * the SAME rounding of the running sum as numpy.cumsum, which adds left to right: one
//     thread does the N dependent additions (~20 us at N = 5000; the sort around it is the
//     parallel part, and textures run side by side, one CTA each).
// The uniform variates come from the caller (numpy's PCG64 stream on the host: the random
// stream is an input of the path, not arithmetic of it).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace drex {

constexpr int RESAMPLE_THREADS = 1024;
constexpr int RESAMPLE_SMEM_NPAD_MAX = 8192;  // 20 B per padded grain -> 160 KB of shared memory

__host__ __device__ inline int64_t resample_npad(int64_t n) {
    int64_t p = 2;
    while (p < n) p <<= 1;
    return p;
}

// keys/cum/idx live in shared memory when npad <= RESAMPLE_SMEM_NPAD_MAX, else in `ws`
// (per texture: npad keys, npad running sums, npad indices).
__global__ void __launch_bounds__(RESAMPLE_THREADS)
resample_kernel(const double *__restrict__ o_aos, const double *__restrict__ f,
                const double *__restrict__ u, int64_t N, int64_t npad, int64_t n_samples,
                double *__restrict__ out_o, double *__restrict__ out_f, unsigned char *ws) {
    extern __shared__ __align__(16) unsigned char resample_smem[];
    const int64_t t = blockIdx.x;
    unsigned char *base = (npad <= RESAMPLE_SMEM_NPAD_MAX) ? resample_smem : ws + t * npad * 20;
    double *keys = reinterpret_cast<double *>(base);
    double *cum = keys + npad;
    int *idx = reinterpret_cast<int *>(cum + npad);
    const double *ft = f + t * N;

    for (int64_t i = threadIdx.x; i < npad; i += blockDim.x) {
        keys[i] = (i < N) ? ft[i] : INFINITY;
        idx[i] = (i < N) ? (int)i : 0x7fffffff;
    }
    __syncthreads();
    // bitonic network on the total order (fraction, index): the result is the stable sort
    for (int64_t k = 2; k <= npad; k <<= 1) {
        for (int64_t j = k >> 1; j > 0; j >>= 1) {
            for (int64_t i = threadIdx.x; i < npad; i += blockDim.x) {
                const int64_t p = i ^ j;
                if (p > i) {
                    const double ka = keys[i], kb = keys[p];
                    const int ia = idx[i], ib = idx[p];
                    const bool a_less = ka < kb || (ka == kb && ia < ib);
                    const bool ascending = (i & k) == 0;
                    if (a_less != ascending) {
                        keys[i] = kb; keys[p] = ka;
                        idx[i] = ib; idx[p] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        double c = 0.0;
        for (int64_t i = 0; i < N; ++i) {
            c += keys[i];
            cum[i] = c;
        }
        cum[N - 1] = 1.0;  // stats.py:58
    }
    __syncthreads();
    const double *ut = u + t * n_samples;
    const double *ot = o_aos + t * N * 9;
    double *oo = out_o + t * n_samples * 9;
    double *of = out_f + t * n_samples;
    for (int64_t e = threadIdx.x; e < n_samples * 9; e += blockDim.x) {
        const int64_t s = e / 9;
        const int c = (int)(e - 9 * s);
        const double us = ut[s];
        int64_t lo = 0, hi = N;  // first i with cum[i] >= us
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (cum[mid] < us) lo = mid + 1; else hi = mid;
        }
        if (lo > N - 1) lo = N - 1;  // u > 1 cannot come from a [0,1) generator; keep the load in range
        oo[e] = ot[(int64_t)idx[lo] * 9 + c];
        if (c == 0) of[s] = keys[lo];
    }
}

}  // namespace drex
