// drex_grain.cuh -- per-grain D-Rex physics in FP64 registers (sm_100a).
//
// One thread evaluates one grain: rotation rate of the crystal axes and the strain
// energy that drives grain-boundary migration.  Replaces the reference's
// core._get_rotation_and_strain and its helpers (src/pydrex/core.py:477-760), restated
// for registers:
//   * the four slip invariants share three matrix-vector products D.a_r;
//   * every supported fabric has exactly one slip system with infinite CRSS
//     (core.py:324-342).  Its sort key |I/inf| is 0, so in the reference's 4-way argsort
//     it (or another system whose invariant is exactly 0) takes the "inactive" slot and
//     in every case ends with slip rate 0.  The argsort therefore reduces to ranking the
//     THREE finite systems -- three comparisons and no data movement -- with identical
//     results for all finite inputs (and NaN where the reference yields NaN);
//   * the Schmid tensor is two outer products; the 81-term permutation-symbol sum is a
//     cross product;
//   * |gamma_k * gamma0|^(p/n) is factored as |r_k|^p * |gamma0|^(p/n) because
//     |gamma_k| = |r_k|^n by construction (core.py:565-566), which leaves ONE general
//     power per grain; the default exponents (p = 1.5, n = 3.5) reduce the others to
//     square roots.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "drex_math.cuh"

namespace drex {

// Per-(phase, fabric) constants, built on the host (drex_capi.cu: make_fabric_consts).
struct FabricConsts {
    double crss[4];      // core.get_crss (core.py:315-342); +inf for the inactive system
    double inv_crss[4];  // 1/crss (0 for inf)
    double rho_pref[3];  // (1/crss_i)^(n-p), i = 0..2 (core.py:675-676)
    double n_minus_1;    // deformation_exponent - 1
    double p;            // stress_exponent
    double p_over_n;
    double lambda;       // nucleation_efficiency
    int phase;           // 0 olivine, 1 enstatite
    int default_exponents;  // p == 1.5 && n == 3.5: powers via sqrt
    int inf_index;       // olivine: index of the infinite-CRSS system (2 for C-type, else 3)
    int pad_;
};

struct GrainOut {
    double da[9];  // orientation_change (core.py:601-642), dimensionless time
    double energy; // strain_energy (core.py:645-684)
};

struct GrainProbe {  // optional intermediates for drex_grain_probe_f64
    double inv[4], rates[4], G[9], gamma0;
};

// a: orientation (already clipped by the caller where the reference clips),
// D: symmetric strain rate (D/smax), L: velocity gradient (L/smax),
// mt: the exp/log tables in shared memory (drex_math.cuh: load_math_tables).
// MODE selects what is decided at compile time (the integrator's hot loop is instantiated
// once per variant so that only that variant's code sits in the instruction cache):
//   0 runtime dispatch on fc, 1 olivine with the default exponents and the fourth slip system
//   inactive (A, B, D, E type), 2 any other olivine (generic exponents or C type), 3 enstatite.
constexpr int GRAIN_RUNTIME = 0, GRAIN_OLIVINE_DEFAULT = 1, GRAIN_OLIVINE_GENERIC = 2,
              GRAIN_ENSTATITE = 3;

template <bool PROBE, int MODE = GRAIN_RUNTIME>
__device__ __forceinline__ void grain_rhs(const double (&a)[9], const double (&D)[9],
                                          const double (&L)[9], const FabricConsts &fc,
                                          const double *__restrict__ mt, GrainOut &out,
                                          GrainProbe *probe, double half_scale = 0.5) {
    const bool olivine = (MODE == GRAIN_RUNTIME) ? (fc.phase == 0) : (MODE != GRAIN_ENSTATITE);
    // GRAIN_OLIVINE_DEFAULT also fixes the infinite-CRSS system to the fourth one (every fabric
    // but C-type), which keeps the second copy of the selection code out of that hot loop
    const bool inf3 = (MODE == GRAIN_OLIVINE_DEFAULT) ? true : (fc.inf_index == 3);  // warp-uniform
    // D.a_r for the crystal axes (rows of a)
    double Da1[3], Da2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        Da1[i] = D[3 * i] * a[3] + D[3 * i + 1] * a[4] + D[3 * i + 2] * a[5];
        Da2[i] = D[3 * i] * a[6] + D[3 * i + 1] * a[7] + D[3 * i + 2] * a[8];
    }
    // slip invariants I_s = l_s . D . n_s (core.py:587-598):
    // (010)[100], (001)[100], (010)[001], (100)[001]
    const double I0 = a[0] * Da1[0] + a[1] * Da1[1] + a[2] * Da1[2];
    const double I1 = a[0] * Da2[0] + a[1] * Da2[1] + a[2] * Da2[2];
    const double I2 = a[6] * Da1[0] + a[7] * Da1[1] + a[8] * Da1[2];
    // I3 = a2.D.a0 equals I1 = a0.D.a2 up to rounding (D is symmetric).  Olivine fabrics
    // whose fourth system has infinite CRSS never use its value (only the all-zero test
    // below looks at it), so the 12 operations are skipped there.
    double I3 = I1;
    if (PROBE || !olivine || !inf3) {
        double Da0[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) Da0[i] = D[3 * i] * a[0] + D[3 * i + 1] * a[1] + D[3 * i + 2] * a[2];
        I3 = a[6] * Da0[0] + a[7] * Da0[1] + a[8] * Da0[2];
    }
    if (PROBE) {
        probe->inv[0] = I0; probe->inv[1] = I1; probe->inv[2] = I2; probe->inv[3] = I3;
    }

    double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;  // 2 x relative slip rates by system
    double w0 = 0.0, w1 = 0.0, w2 = 0.0;            // |gamma_i|^(p/n) for i = 0..2
    bool all_zero = (I0 == 0.0) & (I1 == 0.0) & (I2 == 0.0);
    if (all_zero && !(PROBE || !olivine || !inf3)) {
        // the shortcut I3 = I1 above assumed a symmetric D; core.derivatives takes any matrix
        // from its caller, and the reference tests the true fourth invariant (core.py:721).
        // Rare, so the twelve operations are only spent here.
        double Da0[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) Da0[i] = D[3 * i] * a[0] + D[3 * i + 1] * a[1] + D[3 * i + 2] * a[2];
        I3 = a[6] * Da0[0] + a[7] * Da0[1] + a[8] * Da0[2];
    }
    all_zero = all_zero & (I3 == 0.0);
    const bool default_exponents =
        (MODE == GRAIN_RUNTIME) ? (fc.default_exponents != 0) : (MODE == GRAIN_OLIVINE_DEFAULT);

    if (olivine) {
        // the three finite-CRSS systems (x, y, z); z is natural index 2 or 3 (CTA-uniform)
        const double vz = inf3 ? I2 : I3;
        const double rz = inf3 ? fc.inv_crss[2] : fc.inv_crss[3];
        const double cz = inf3 ? fc.crss[2] : fc.crss[3];
        const double ux = I0 * fc.inv_crss[0], uy = I1 * fc.inv_crss[1], uz = vz * rz;  // I/crss
        const double kx = fabs(ux), ky = fabs(uy), kz = fabs(uz);
        // stable ranks (ties keep index order, like the reference's insertion argsort)
        const bool xy = kx <= ky, xz = kx <= kz, yz = ky <= kz;
        const bool x_max = !xy && !xz;  // rank 2
        const bool y_max = xy && !yz;
        const bool z_max = xz && yz;
        const double I_max = z_max ? vz : (y_max ? I1 : I0);
        const double c_max = z_max ? cz : (y_max ? fc.crss[1] : fc.crss[0]);
        const double pre = div_fast(c_max, I_max);  // core.py:560
        const double qx = pre * ux, qy = pre * uy, qz = pre * uz;
        const double ax = fabs(qx), ay = fabs(qy), az = fabs(qz);
        double rx, ry, rz_, px, py, pz;  // q|q|^(n-1) and |q|^p
        if (default_exponents) {
            const double sx = sqrt_fast(ax), sy = sqrt_fast(ay), sz = sqrt_fast(az);
            rx = qx * (ax * ax * sx); ry = qy * (ay * ay * sy); rz_ = qz * (az * az * sz);
            px = ax * sx; py = ay * sy; pz = az * sz;
        } else {
            rx = qx * pow_nonneg(ax, fc.n_minus_1, mt);
            ry = qy * pow_nonneg(ay, fc.n_minus_1, mt);
            rz_ = qz * pow_nonneg(az, fc.n_minus_1, mt);
            px = pow_nonneg(ax, fc.p, mt); py = pow_nonneg(ay, fc.p, mt); pz = pow_nonneg(az, fc.p, mt);
        }
        // the softest system slips at rate exactly 1 (core.py:567); g carries the factor 2
        // of the Schmid tensor (exact scaling, G below is bit-identical to 2 (a0 u + a2 v))
        g0 = x_max ? 2.0 : rx + rx;
        g1 = y_max ? 2.0 : ry + ry;
        const double gz = z_max ? 2.0 : rz_ + rz_;
        w0 = x_max ? 1.0 : px;
        w1 = y_max ? 1.0 : py;
        const double wz = z_max ? 1.0 : pz;
        g2 = inf3 ? gz : 0.0;
        g3 = inf3 ? 0.0 : gz;
        w2 = inf3 ? wz : 0.0;  // C-type: natural index 2 is the infinite-CRSS system
    } else {
        // enstatite: only (100)[001] slips (core.py:731-739)
        g3 = (fabs(I3) > 1e-15) ? 2.0 : 0.0;
    }

    // Schmid tensor G = 2 sum_s gamma_s l_s (x) n_s = 2 [a0 (x) u + a2 (x) v],
    // u = g0 a1 + g1 a2, v = g2 a1 + g3 a0 (core.py:492-501)
    double G[9];
    {
        const double u[3] = {g0 * a[3] + g1 * a[6], g0 * a[4] + g1 * a[7], g0 * a[5] + g1 * a[8]};
        const double v[3] = {g2 * a[3] + g3 * a[0], g2 * a[4] + g3 * a[1], g2 * a[5] + g3 * a[2]};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) G[3 * i + j] = a[i] * u[j] + a[6 + i] * v[j];
    }
    // slip rate on the softest system (core.py:513-538)
    // three independent partial sums per row instead of one 12-term chain: the FP64 pipe has
    // a long dependent-issue latency and only four warps per scheduler to hide it
    double num, den;
    double ga[3], wl[3];  // antisymmetric parts G_jk - G_kj and L_jk - L_kj, k = j + 1 (mod 3)
    {
        double nr[3], dr[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int k = (j + 1) % 3;
            ga[j] = G[3 * j + k] - G[3 * k + j];
            wl[j] = L[3 * j + k] - L[3 * k + j];
            double n2 = G[3 * j] * L[3 * j], d2 = G[3 * j] * G[3 * j];
            n2 = fma(G[3 * j + 1], L[3 * j + 1], n2);
            d2 = fma(G[3 * j + 1], G[3 * j + 1], d2);
            n2 = fma(G[3 * j + 2], L[3 * j + 2], n2);
            d2 = fma(G[3 * j + 2], G[3 * j + 2], d2);
            nr[j] = fma(2.0, n2, -wl[j] * ga[j]);
            dr[j] = fma(2.0, d2, -ga[j] * ga[j]);
        }
        num = (nr[0] + nr[1]) + nr[2];
        den = (dr[0] + dr[1]) + dr[2];
    }
    const double gamma0 = (den > -1e-15 && den < 1e-15) ? 0.0 : div_fast(num, den);

    // spin vector and da_p = omega x a_p (core.py:616-642):
    // omega_j = ((L_sr - L_rs) - (G_sr - G_rs) gamma0) / 2 with r = j + 1, s = j + 2 (mod 3),
    // i.e. minus the antisymmetric pair stored at index r; the caller's time scale rides
    // along in half_scale (= scale / 2)
    double om[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int r = (j + 1) % 3;
        om[j] = half_scale * fma(ga[r], gamma0, -wl[r]);
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        out.da[3 * p + 0] = om[1] * a[3 * p + 2] - om[2] * a[3 * p + 1];
        out.da[3 * p + 1] = om[2] * a[3 * p + 0] - om[0] * a[3 * p + 2];
        out.da[3 * p + 2] = om[0] * a[3 * p + 1] - om[1] * a[3 * p + 0];
    }

    // strain energy over slip systems 0..2 in natural order (core.py:674-684)
    double energy = 0.0;
    if (olivine) {
        const double gp = pow_nonneg(fabs(gamma0), fc.p_over_n, mt);
        const double rho0 = fc.rho_pref[0] * (w0 * gp);
        const double rho1 = fc.rho_pref[1] * (w1 * gp);
        energy = rho0 * exp_tab(-fc.lambda * rho0 * rho0, mt) + rho1 * exp_tab(-fc.lambda * rho1 * rho1, mt);
        if (inf3) {  // otherwise rho_2 = (1/inf)^(n-p) * ... = 0 contributes 0
            const double rho2 = fc.rho_pref[2] * (w2 * gp);
            energy += rho2 * exp_tab(-fc.lambda * rho2 * rho2, mt);
        }
    }
    if (all_zero) {  // core.py:721-722
#pragma unroll
        for (int i = 0; i < 9; ++i) out.da[i] = 0.0;
        energy = 0.0;
    }
    out.energy = energy;
    if (PROBE) {
        probe->rates[0] = all_zero ? 0.0 : 0.5 * g0;  // g carries the Schmid factor 2
        probe->rates[1] = all_zero ? 0.0 : 0.5 * g1;  // g carries the Schmid factor 2
        probe->rates[2] = all_zero ? 0.0 : 0.5 * g2;  // g carries the Schmid factor 2
        probe->rates[3] = all_zero ? 0.0 : 0.5 * g3;  // g carries the Schmid factor 2
#pragma unroll
        for (int i = 0; i < 9; ++i) probe->G[i] = all_zero ? 0.0 : G[i];
        probe->gamma0 = all_zero ? 0.0 : gamma0;
    }
}

}  // namespace drex
