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
constexpr int RESAMPLE_SMEM_N_MAX = 11000;  // 20 B per grain: up to 215 KB of shared memory

// bytes per texture in the workspace: N keys, N running sums, N indices
__host__ __device__ inline int64_t resample_ws_stride(int64_t n) { return (n * 20 + 15) / 16 * 16; }

__host__ __device__ inline int64_t resample_npad(int64_t n) {
    int64_t p = 2;
    while (p < n) p <<= 1;
    return p;
}

// keys/cum/idx (N each) live in shared memory when N <= RESAMPLE_SMEM_N_MAX, else in `ws`.
// The sort is the bitonic network in its all-ascending form (the first step of every merge
// pairs i with its mirror image i ^ (k - 1) inside the block of k, the others i with i ^ j):
// padding up to the next power of two then stays at the end as virtual +infinity entries that
// never move, so nothing is stored or compared for them -- N = 5000 needs 100 KB (two textures
// per SM) instead of the 160 KB of 8192 padded entries, and 39 % fewer comparisons.
__global__ void __launch_bounds__(RESAMPLE_THREADS)
resample_kernel(const double *__restrict__ o_aos, const double *__restrict__ f,
                const double *__restrict__ u, int64_t N, int64_t npad, int64_t n_samples,
                double *__restrict__ out_o, double *__restrict__ out_f, unsigned char *ws) {
    extern __shared__ __align__(16) unsigned char resample_smem[];
    const int64_t t = blockIdx.x;
    unsigned char *base = (N <= RESAMPLE_SMEM_N_MAX) ? resample_smem : ws + t * resample_ws_stride(N);
    double *keys = reinterpret_cast<double *>(base);
    double *cum = keys + N;
    int *idx = reinterpret_cast<int *>(cum + N);
    const double *ft = f + t * N;
    const int n = (int)N;

    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        keys[i] = ft[i];
        idx[i] = i;
    }
    __syncthreads();
    // on the total order (fraction, index): the result is the stable sort
    auto exchange = [&](int i, int p) {
        const double ka = keys[i], kb = keys[p];
        const int ia = idx[i], ib = idx[p];
        if (kb < ka || (ka == kb && ib < ia)) {
            keys[i] = kb; keys[p] = ka;
            idx[i] = ib; idx[p] = ia;
        }
    };
    for (int k = 2; k <= (int)npad; k <<= 1) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int p = i ^ (k - 1);
            if (p > i && p < n) exchange(i, p);
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i && p < n) exchange(i, p);
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

// numpy.random.default_rng(seed).random(n) on the device, bit for bit: the reference draws the
// variates of a resample from that stream (stats.py:49, 59), one generator per call, and its
// drivers call it once per particle with the particle's seed -- at 1250 particles the host's
// generator set-up (~12 us each) was four times the device work of the whole resample.
//   SeedSequence(seed): the seed's 32-bit words hashed into a pool of four (numpy
//     bit_generator.pyx: hashmix / mix, constants below), then generate_state(4, uint64);
//   PCG64: 128-bit LCG (multiplier 0x2360ED051FC65DA44385DF649FCCF645, increment from the
//     seed), output XSL-RR of the NEW state; random() = (next >> 11) * 2^-53.
// One thread per stream walks its n steps; `offset` variates are skipped first with the
// O(log offset) jump of the LCG (pcg_advance_lcg_128), so that one seeded stream can be cut
// into per-texture pieces (stats.py:59 draws texture after texture from one generator).
__device__ __forceinline__ uint32_t seedseq_hashmix(uint32_t v, uint32_t &hash_const) {
    v ^= hash_const;
    hash_const *= 0x931e8875u;
    v *= hash_const;
    return v ^ (v >> 16);
}
__device__ __forceinline__ uint32_t seedseq_mix(uint32_t x, uint32_t y) {
    const uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
    return r ^ (r >> 16);
}
__global__ void __launch_bounds__(64)
numpy_uniform_kernel(const uint64_t *__restrict__ seeds, const uint64_t *__restrict__ offsets, int64_t T,
                     int64_t n, double *__restrict__ out) {
    typedef unsigned __int128 u128;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const uint64_t seed = seeds[t];
    // SeedSequence: entropy words (little endian, at least one), pool of four
    uint32_t ent[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const int n_ent = ent[1] ? 2 : 1;
    uint32_t hc = 0x43b0d7e5u, pool[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pool[i] = seedseq_hashmix(i < n_ent ? ent[i] : 0u, hc);
#pragma unroll
    for (int src = 0; src < 4; ++src)
#pragma unroll
        for (int dst = 0; dst < 4; ++dst)
            if (src != dst) pool[dst] = seedseq_mix(pool[dst], seedseq_hashmix(pool[src], hc));
    uint32_t w[8], hb = 0x8b51f9ddu;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t d = pool[i & 3] ^ hb;
        hb *= 0x58f38dedu;
        d *= hb;
        w[i] = d ^ (d >> 16);
    }
    const uint64_t v0 = w[0] | ((uint64_t)w[1] << 32), v1 = w[2] | ((uint64_t)w[3] << 32);
    const uint64_t v2 = w[4] | ((uint64_t)w[5] << 32), v3 = w[6] | ((uint64_t)w[7] << 32);
    const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
    const u128 initstate = ((u128)v0 << 64) | v1, initseq = ((u128)v2 << 64) | v3;
    const u128 inc = (initseq << 1) | 1u;
    u128 state = inc;  // pcg_setseq_128_srandom_r: 0 -> step -> + initstate -> step
    state += initstate;
    state = state * mult + inc;
    uint64_t delta = offsets ? offsets[t] : 0ull;
    if (delta) {
        u128 acc_mult = 1u, acc_plus = 0u, cur_mult = mult, cur_plus = inc;
        while (delta) {
            if (delta & 1ull) {
                acc_mult *= cur_mult;
                acc_plus = acc_plus * cur_mult + cur_plus;
            }
            cur_plus = (cur_mult + 1u) * cur_plus;
            cur_mult *= cur_mult;
            delta >>= 1;
        }
        state = acc_mult * state + acc_plus;
    }
    double *o = out + t * n;
    for (int64_t k = 0; k < n; ++k) {
        state = state * mult + inc;
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        const uint64_t x = hi ^ lo;
        const unsigned rot = (unsigned)(hi >> 58);
        const uint64_t r = (x >> rot) | (x << ((64u - rot) & 63u));
        o[k] = (double)(r >> 11) * (1.0 / 9007199254740992.0);
    }
}

}  // namespace drex
