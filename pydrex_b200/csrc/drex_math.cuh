// drex_math.cuh -- table-driven FP64 exp / log / sqrt for the per-grain physics (sm_100a).
//
// The integrator is bound by issue slots, and every FP64 instruction costs two of them
// (16 FP64 lanes per scheduler).  CUDA's exp() and log() spend ~30 and ~45 FP64
// instructions on degree-11+ polynomials and special cases; with a 64-entry table in shared
// memory the polynomials shrink to degree 6 / 8 and the functions to ~12 FP64 instructions
// each, at the same ~1 ulp accuracy (the parity tests hold 1e-10 relative on derivatives
// that cancel, so full double accuracy is required -- lower-precision shortcuts are not an
// option here).  Polynomial coefficients are read from constant memory as instruction
// operands (no immediate-materialising moves); range handling is folded into integer
// clamps and one select per pow() instead of three selects per exp().
//
//   exp:  x = (64 m + j) ln2/64 + r, |r| <= ln2/128;  e^x = 2^m * 2^(j/64) * e^r
//   log:  x = 2^e * mant, mant in [1,2), j = top 6 mantissa bits, c_j = 1 + (j + 1/2)/64,
//         r = mant/c_j - 1, |r| <= 2^-7;  ln x = e ln2 + ln c_j + log1p(r)
//   sqrt: hardware rsqrt approximation + two Goldschmidt steps + one residual correction
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

// polynomial coefficients: exp (r^6 ... r^2 of e^r), log1p (r^8 ... r^2), split constants
__device__ __constant__ double DREX_EXP_C[5] = {1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};
__device__ __constant__ double DREX_LOG_C[7] = {-1.0 / 8.0, 1.0 / 7.0,  -1.0 / 6.0, 0.2,
                                                -0.25,      1.0 / 3.0, -0.5};
__device__ __constant__ double DREX_LN2[5] = {92.33248261689366,       // 64 / ln2
                                              -0.010830424696223417,   // -(ln2/64) high 36 bits
                                              -2.572804622327669e-14,  // -(ln2/64) low part
                                              0.6931471805598903,      // ln2 high 42 bits
                                              5.497923018708371e-14};  // ln2 low part

// copy the tables into `dst` (shared memory, MATH_TABLE_DOUBLES doubles); caller syncs
__device__ __forceinline__ void load_math_tables(double *dst) {
    for (int i = threadIdx.x; i < MATH_TABLE_DOUBLES; i += blockDim.x)
        dst[i] = (i < 64) ? DREX_EXP2_TAB[i] : DREX_LOG_TAB[i - 64];
}

// e^x for finite x < ~709 (and NaN, which propagates).  Arguments below -708 give a value
// of order 1e-308 instead of a denormal or 0 -- indistinguishable for every use here; the
// caller handles -inf (see pow_nonneg).
__device__ __forceinline__ double exp_tab(double x, const double *__restrict__ mt) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rint() in the low mantissa word
    const double t = fma(x, DREX_LN2[0], MAGIC);
    const int n = __double2loint(t);
    const double nd = t - MAGIC;
    double r = fma(nd, DREX_LN2[1], x);  // nd * hi is exact
    r = fma(nd, DREX_LN2[2], r);
    // Estrin: e^r - 1 = r + r^2 (1/2 + r/6) + r^4 (1/24 + r/120 + r^2/720), depth 3 instead of 6
    const double r2 = r * r;
    const double a0 = fma(r, DREX_EXP_C[3], DREX_EXP_C[4]);   // 1/2 + r/6
    const double a1 = fma(r, DREX_EXP_C[1], DREX_EXP_C[2]);   // 1/24 + r/120
    const double a2 = fma(r2, DREX_EXP_C[0], a1);             // + r^2/720
    const double r4 = r2 * r2;
    double p = fma(r2, a0, r);
    p = fma(r4, a2, p);  // e^r - 1
    const double tj = mt[n & 63];
    const double v = fma(tj, p, tj);
    const int m = max(n >> 6, -1022);  // keep the result a normal number
    return __hiloint2double(__double2hiint(v) + (m << 20), __double2loint(v));  // * 2^m
}

// natural log of a finite normal x > 0.  x = 0, subnormals, inf and NaN give finite garbage:
// callers that can see those values patch the result (pow_nonneg).
__device__ __forceinline__ double log_tab(double x, const double *__restrict__ mt) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const int j = (hi >> 14) & 63;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 tab = *reinterpret_cast<const double2 *>(mt + 64 + 2 * j);  // {1/c_j, ln c_j}
    const double r = fma(m, tab.x, -1.0);
    // Estrin on q(r) = -1/2 + r/3 - r^2/4 + r^3/5 - r^4/6 + r^5/7 - r^6/8
    const double r2 = r * r;
    const double b0 = fma(r, DREX_LOG_C[5], DREX_LOG_C[6]);  // -1/2 + r/3
    const double b1 = fma(r, DREX_LOG_C[3], DREX_LOG_C[4]);  // -1/4 + r/5
    const double b2 = fma(r, DREX_LOG_C[1], DREX_LOG_C[2]);  // -1/6 + r/7
    const double r4 = r2 * r2;
    const double b3 = fma(r2, DREX_LOG_C[0], b2);            // ... - r^2/8
    double q = fma(r2, b1, b0);
    q = fma(r4, b3, q);
    const double ed = (double)e;
    double res = fma(ed, DREX_LN2[3], tab.y);  // e * ln2_hi is exact (42 bits x 11 bits)
    res += fma(r2, q, r);
    return fma(ed, DREX_LN2[4], res);
}

// x^y for x >= 0 and a model exponent y (uniform over the launch), C pow() conventions at
// 0; NaN and inf propagate.  exp(y log x) is accurate to ~|y log x| * 2^-52 relative, far
// inside the 1e-10 bound.
__device__ __forceinline__ double pow_nonneg(double x, double y, const double *__restrict__ mt) {
    if (y == 0.0) return 1.0;
    double v = exp_tab(y * log_tab(x, mt), mt);
    v += x - x;  // NaN (and inf) in, NaN out: log_tab/exp_tab drop them
    const bool tiny = (unsigned)__double2hiint(x) < 0x00100000u;  // +0 and subnormals
    return tiny ? ((y > 0.0) ? 0.0 : INFINITY) : v;
}

// a / b to ~1 ulp for normal b without the range checks and slow path of the library
// division (~22 instructions): hardware reciprocal approximation, two Newton steps, one
// residual correction.  b = 0 (and subnormals, flushed) give inf or NaN like a / 0.
__device__ __forceinline__ double div_fast(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(r, fma(-b, r, 1.0), r);
    r = fma(r, fma(-b, r, 1.0), r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// sqrt of x >= 0 to ~1 ulp without the special-case branches of the library routine.
__device__ __forceinline__ double sqrt_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    g = fma(fma(-g, g, x), h, g);  // residual correction
    // x = 0 (rsqrt = inf -> NaN) and subnormals: sqrt is 0 to every purpose here
    return ((unsigned)__double2hiint(x) < 0x00100000u) ? 0.0 : g;
}

}  // namespace drex
