// drex_grain.cuh -- per-grain D-Rex physics in FP64 registers (sm_100a).
//
// One thread evaluates one grain: rotation rate of the crystal axes and the strain
// energy that drives grain-boundary migration.  Replaces the reference's
// core._get_rotation_and_strain and its helpers (src/pydrex/core.py:477-760), restated
// for registers:
//   * the four slip invariants share three matrix-vector products D.a_r;
//   * the 4-element argsort is a stable insertion network on (key, value) tuples, so no
//     dynamically indexed local arrays exist;
//   * the Schmid tensor is two outer products; the 81-term permutation-symbol sum is a
//     cross product;
//   * |gamma_k * gamma0|^(p/n) is factored as |r_k|^p * |gamma0|^(p/n) because
//     |gamma_k| = |r_k|^n by construction (core.py:565-566), which leaves ONE general
//     power per grain; the default exponents (p = 1.5, n = 3.5) reduce the others to
//     square roots.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace drex {

// Per-(phase, fabric) constants, built on the host (drex_capi.cu: make_fabric_consts).
struct FabricConsts {
    double crss[4];      // core.get_crss (core.py:315-342); +inf for inactive systems
    double inv_crss[4];  // 1/crss (0 for inf)
    double rho_pref[3];  // (1/crss_i)^(n-p), i = 0..2 (core.py:675-676)
    double n_minus_1;    // deformation_exponent - 1
    double p;            // stress_exponent
    double p_over_n;
    double lambda;       // nucleation_efficiency
    int phase;           // 0 olivine, 1 enstatite
    int default_exponents;  // p == 1.5 && n == 3.5: powers via sqrt
};

// x^y for x >= 0 with C pow() conventions at 0; exp(y log x) is accurate to
// ~|y log x| * 2^-53 relative, far inside the 1e-10 parity bound.
__device__ __forceinline__ double pow_nonneg(double x, double y) {
    if (x == 0.0) return (y == 0.0) ? 1.0 : (y > 0.0 ? 0.0 : INFINITY);
    return exp(y * log(x));
}

struct GrainOut {
    double da[9];  // orientation_change (core.py:601-642), dimensionless time
    double energy; // strain_energy (core.py:645-684)
};

struct GrainProbe {  // optional intermediates for drex_grain_probe_f64
    double inv[4], rates[4], G[9], gamma0;
};

#define DREX_SWAP_IF(cond, x, y) \
    do {                         \
        double t_ = (cond) ? (y) : (x); \
        (y) = (cond) ? (x) : (y);       \
        (x) = t_;                       \
    } while (0)

// a: orientation (already clipped by the caller where the reference clips),
// D: symmetric strain rate (D/smax), L: velocity gradient (L/smax).
template <bool PROBE>
__device__ __forceinline__ void grain_rhs(const double (&a)[9], const double (&D)[9],
                                          const double (&L)[9], const FabricConsts &fc,
                                          GrainOut &out, GrainProbe *probe) {
    // D.a_r for the three crystal axes (rows of a)
    double Da0[3], Da1[3], Da2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        Da0[i] = D[3 * i] * a[0] + D[3 * i + 1] * a[1] + D[3 * i + 2] * a[2];
        Da1[i] = D[3 * i] * a[3] + D[3 * i + 1] * a[4] + D[3 * i + 2] * a[5];
        Da2[i] = D[3 * i] * a[6] + D[3 * i + 1] * a[7] + D[3 * i + 2] * a[8];
    }
    // slip invariants I_s = l_s . D . n_s (core.py:587-598):
    // (010)[100], (001)[100], (010)[001], (100)[001]
    const double I0 = a[0] * Da1[0] + a[1] * Da1[1] + a[2] * Da1[2];
    const double I1 = a[0] * Da2[0] + a[1] * Da2[1] + a[2] * Da2[2];
    const double I2 = a[6] * Da1[0] + a[7] * Da1[1] + a[8] * Da1[2];
    const double I3 = a[6] * Da0[0] + a[7] * Da0[1] + a[8] * Da0[2];
    if (PROBE) {
        probe->inv[0] = I0; probe->inv[1] = I1; probe->inv[2] = I2; probe->inv[3] = I3;
    }

    double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;  // relative slip rates by system
    double w0 = 0.0, w1 = 0.0, w2 = 0.0;            // |gamma_i|^(p/n) for i = 0..2
    const bool all_zero = (I0 == 0.0) & (I1 == 0.0) & (I2 == 0.0) & (I3 == 0.0);

    if (fc.phase == 0) {
        // stable insertion argsort of |I/crss| carrying (I, 1/crss, crss, system id)
        double k0 = fabs(I0 * fc.inv_crss[0]), k1 = fabs(I1 * fc.inv_crss[1]);
        double k2 = fabs(I2 * fc.inv_crss[2]), k3 = fabs(I3 * fc.inv_crss[3]);
        double v0 = I0, v1 = I1, v2 = I2, v3 = I3;
        double r0 = fc.inv_crss[0], r1 = fc.inv_crss[1], r2 = fc.inv_crss[2], r3 = fc.inv_crss[3];
        double c0 = fc.crss[0], c1 = fc.crss[1], c2 = fc.crss[2], c3 = fc.crss[3];
        double s0 = 0.0, s1 = 1.0, s2 = 2.0, s3 = 3.0;
#define DREX_CSWAP(i, j)                      \
    do {                                      \
        const bool sw_ = k##i > k##j;         \
        DREX_SWAP_IF(sw_, k##i, k##j);        \
        DREX_SWAP_IF(sw_, v##i, v##j);        \
        DREX_SWAP_IF(sw_, r##i, r##j);        \
        DREX_SWAP_IF(sw_, c##i, c##j);        \
        DREX_SWAP_IF(sw_, s##i, s##j);        \
    } while (0)
        // insertion sort as adjacent compare-exchanges (strict >, hence stable)
        DREX_CSWAP(0, 1);
        DREX_CSWAP(1, 2);
        DREX_CSWAP(0, 1);
        DREX_CSWAP(2, 3);
        DREX_CSWAP(1, 2);
        DREX_CSWAP(0, 1);
#undef DREX_CSWAP
        // sorted: 0 = inactive, 1 = min, 2 = int, 3 = max (core.py:556-567)
        const double pre = c3 / v3;
        const double q_min = pre * v1 * r1;
        const double q_int = pre * v2 * r2;
        double rate_min, rate_int, wp_min, wp_int;
        const double am = fabs(q_min), ai = fabs(q_int);
        if (fc.default_exponents) {
            const double sm = sqrt(am), si = sqrt(ai);
            rate_min = q_min * (am * am * sm);  // q |q|^2.5
            rate_int = q_int * (ai * ai * si);
            wp_min = am * sm;                   // |q|^1.5
            wp_int = ai * si;
        } else {
            rate_min = q_min * pow_nonneg(am, fc.n_minus_1);
            rate_int = q_int * pow_nonneg(ai, fc.n_minus_1);
            wp_min = pow_nonneg(am, fc.p);
            wp_int = pow_nonneg(ai, fc.p);
        }
        // scatter back to natural slip-system order
#define DREX_PICK(id, field_min, field_int, field_max) \
    ((s1 == (id)) ? (field_min) : ((s2 == (id)) ? (field_int) : ((s3 == (id)) ? (field_max) : 0.0)))
        g0 = DREX_PICK(0.0, rate_min, rate_int, 1.0);
        g1 = DREX_PICK(1.0, rate_min, rate_int, 1.0);
        g2 = DREX_PICK(2.0, rate_min, rate_int, 1.0);
        g3 = DREX_PICK(3.0, rate_min, rate_int, 1.0);
        w0 = DREX_PICK(0.0, wp_min, wp_int, 1.0);
        w1 = DREX_PICK(1.0, wp_min, wp_int, 1.0);
        w2 = DREX_PICK(2.0, wp_min, wp_int, 1.0);
#undef DREX_PICK
    } else {
        // enstatite: only (100)[001] slips (core.py:731-739)
        g3 = (fabs(I3) > 1e-15) ? 1.0 : 0.0;
    }

    // Schmid tensor G = 2 sum_s gamma_s l_s (x) n_s = 2 [a0 (x) u + a2 (x) v],
    // u = g0 a1 + g1 a2, v = g2 a1 + g3 a0 (core.py:492-501)
    double G[9];
    {
        const double u0 = g0 * a[3] + g1 * a[6], u1 = g0 * a[4] + g1 * a[7], u2 = g0 * a[5] + g1 * a[8];
        const double t0 = g2 * a[3] + g3 * a[0], t1 = g2 * a[4] + g3 * a[1], t2 = g2 * a[5] + g3 * a[2];
        const double u[3] = {u0, u1, u2}, v[3] = {t0, t1, t2};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) G[3 * i + j] = 2.0 * (a[i] * u[j] + a[6 + i] * v[j]);
    }
    // slip rate on the softest system (core.py:513-538)
    double num = 0.0, den = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int k = (j + 1) % 3;
        const double ga = G[3 * j + k] - G[3 * k + j];
        num -= (L[3 * j + k] - L[3 * k + j]) * ga;
        den -= ga * ga;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            num += 2.0 * G[3 * j + l] * L[3 * j + l];
            den += 2.0 * G[3 * j + l] * G[3 * j + l];
        }
    }
    const double gamma0 = (den > -1e-15 && den < 1e-15) ? 0.0 : num / den;

    // spin vector and da_p = omega x a_p (core.py:616-642)
    double om[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int r = (j + 1) % 3, s = (j + 2) % 3;
        om[j] = 0.5 * ((L[3 * s + r] - L[3 * r + s]) - (G[3 * s + r] - G[3 * r + s]) * gamma0);
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        out.da[3 * p + 0] = om[1] * a[3 * p + 2] - om[2] * a[3 * p + 1];
        out.da[3 * p + 1] = om[2] * a[3 * p + 0] - om[0] * a[3 * p + 2];
        out.da[3 * p + 2] = om[0] * a[3 * p + 1] - om[1] * a[3 * p + 0];
    }

    // strain energy over slip systems 0..2 in natural order (core.py:674-684)
    double energy = 0.0;
    if (fc.phase == 0) {
        const double gp = pow_nonneg(fabs(gamma0), fc.p_over_n);
        const double rho0 = fc.rho_pref[0] * (w0 * gp);
        const double rho1 = fc.rho_pref[1] * (w1 * gp);
        const double rho2 = fc.rho_pref[2] * (w2 * gp);
        energy = rho0 * exp(-fc.lambda * rho0 * rho0) + rho1 * exp(-fc.lambda * rho1 * rho1) +
                 rho2 * exp(-fc.lambda * rho2 * rho2);
    }
    if (all_zero) {  // core.py:721-722
#pragma unroll
        for (int i = 0; i < 9; ++i) out.da[i] = 0.0;
        energy = 0.0;
    }
    out.energy = energy;
    if (PROBE) {
        probe->rates[0] = all_zero ? 0.0 : g0;
        probe->rates[1] = all_zero ? 0.0 : g1;
        probe->rates[2] = all_zero ? 0.0 : g2;
        probe->rates[3] = all_zero ? 0.0 : g3;
#pragma unroll
        for (int i = 0; i < 9; ++i) probe->G[i] = all_zero ? 0.0 : G[i];
        probe->gamma0 = all_zero ? 0.0 : gamma0;
    }
}

}  // namespace drex
