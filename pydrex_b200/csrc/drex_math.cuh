// drex_math.cuh -- table-driven FP64 exp / log for the per-grain physics (sm_100a).
//
// The integrator is bound by issue slots, and every FP64 instruction costs two of them
// (16 FP64 lanes per scheduler).  CUDA's exp() and log() spend ~30 and ~45 FP64
// instructions on degree-11+ polynomials and special cases; with a 64-entry table in shared
// memory the polynomials shrink to degree 6 / 8 and the functions to ~12 FP64 instructions
// each, at the same ~1 ulp accuracy (the parity tests hold 1e-10 relative on derivatives
// that cancel, so full double accuracy is required -- lower-precision shortcuts are not an
// option here).
//
//   exp:  x = (64 m + j) ln2/64 + r, |r| <= ln2/128;  e^x = 2^m * 2^(j/64) * e^r
//   log:  x = 2^e * mant, mant in [1,2), j = top 6 mantissa bits, c_j = 1 + (j + 1/2)/64,
//         r = mant/c_j - 1, |r| <= 2^-7;  ln x = e ln2 + ln c_j + log1p(r)
//
// Tables (scripts/gen_math_tables.py, correctly rounded from 60-digit decimals) live in
// constant memory and are copied to shared memory by each CTA: the per-thread index
// diverges, which constant memory would serialise.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace drex {

#include "drex_math_tables.inc"

constexpr int MATH_TABLE_DOUBLES = 64 + 128;

// copy the tables into `dst` (shared memory, MATH_TABLE_DOUBLES doubles); caller syncs
__device__ __forceinline__ void load_math_tables(double *dst) {
    for (int i = threadIdx.x; i < MATH_TABLE_DOUBLES; i += blockDim.x)
        dst[i] = (i < 64) ? DREX_EXP2_TAB[i] : DREX_LOG_TAB[i - 64];
}

__device__ __forceinline__ double exp_tab(double x, const double *__restrict__ mt) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rint() in the low mantissa word
    const double t = fma(x, 92.33248261689366, MAGIC);  // x * 64/ln2
    const int n = __double2loint(t);
    const double nd = t - MAGIC;
    double r = fma(nd, -0.010830424696223417, x);  // ln2/64, high 36 bits: nd * hi is exact
    r = fma(nd, -2.572804622327669e-14, r);
    double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = p * r;  // e^r - 1
    const double tj = mt[n & 63];
    double v = fma(tj, p, tj);
    v = __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));  // * 2^m
    v = (x < -708.0) ? 0.0 : v;       // (sub)normal underflow: not needed to 1e-308
    v = (x > 709.0) ? INFINITY : v;
    return (x != x) ? x : v;
}

// natural log of x >= 0 (x = 0 and subnormals give -inf, inf/NaN pass through)
__device__ __forceinline__ double log_tab(double x, const double *__restrict__ mt) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const int j = (hi >> 14) & 63;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 tab = *reinterpret_cast<const double2 *>(mt + 64 + 2 * j);  // {1/c_j, ln c_j}
    const double r = fma(m, tab.x, -1.0);
    double q = fma(r, -1.0 / 8.0, 1.0 / 7.0);
    q = fma(q, r, -1.0 / 6.0);
    q = fma(q, r, 0.2);
    q = fma(q, r, -0.25);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -0.5);
    const double r2 = r * r;
    const double ed = (double)e;
    // e*ln2_hi is exact (ln2_hi has 42 significant bits, |e| < 2^11)
    double res = fma(ed, 0.6931471805598903, tab.y);
    res += fma(r2, q, r);
    res = fma(ed, 5.497923018708371e-14, res);
    res = (hi < 0x00100000) ? -INFINITY : res;
    return (hi >= 0x7ff00000) ? x : res;
}

// x^y for x >= 0 with C pow() conventions at 0: log(0) = -inf makes x = 0 come out right
// for y != 0 without a per-thread branch; y is a model exponent (uniform over the launch).
// exp(y log x) is accurate to ~|y log x| * 2^-52 relative, far inside the 1e-10 bound.
__device__ __forceinline__ double pow_nonneg(double x, double y, const double *__restrict__ mt) {
    return (y == 0.0) ? 1.0 : exp_tab(y * log_tab(x, mt), mt);
}

}  // namespace drex
