// drex_derivatives.cuh -- batched core.derivatives (src/pydrex/core.py:347-474) and the
// layout / probe kernels.
//
// Data layout: orientations SoA [S][9][N] so that the 9 per-grain loads of a warp are 9
// fully coalesced 256 B transactions; fractions [S][N].
//
// Two launches per call:
//   K1 derivatives_grain_kernel  grid (chunks, S): one thread per grain, reads 72 B,
//      writes 72 B of rotation rates + 8 B strain energy (workspace);
//   K2 derivatives_gbm_kernel    one CTA per system: Ebar = sum_g f_g E_g by a fixed-order
//      block reduction (deterministic), then df_g = vol_frac M f_g (Ebar - E_g)
//      (core.py:431-436); the second sweep over f/E hits L1/L2.
// Algorithmic traffic 160 B per grain-evaluation; actual 176 B
// (the energy round trip through the workspace).
#pragma once
#include "drex_common.cuh"
#include "drex_grain.cuh"

namespace drex {

struct FabricTable {
    FabricConsts fab[6];  // indexed by MineralFabric (core.py:50-67)
};

constexpr int DERIV_BLOCK = 128;

// AoS [S][N][9] <-> SoA [S][9][N].  A block moves 256 grains through shared memory so that BOTH
// sides are coalesced: the AoS side as one contiguous run of 256 x 72 B, the SoA side as nine
// rows of 2 KB (a thread-per-grain copy reads or writes the AoS side at a 72 B stride).  The
// system index rides in grid.x (no 65535 limit), the 256-grain tile in grid.y.
constexpr int LAYOUT_TILE = 256;

__global__ void __launch_bounds__(LAYOUT_TILE) pack_soa_kernel(const double *__restrict__ aos,
                                                               double *__restrict__ soa, int64_t S,
                                                               int64_t N) {
    __shared__ double tile[LAYOUT_TILE * 9];
    const int64_t s = blockIdx.x, g0 = (int64_t)blockIdx.y * LAYOUT_TILE;
    const int n = (int)((N - g0 < LAYOUT_TILE) ? (N - g0) : LAYOUT_TILE);
    const double *src = aos + (s * N + g0) * 9;
    for (int i = threadIdx.x; i < n * 9; i += LAYOUT_TILE) tile[i] = src[i];
    __syncthreads();
    if ((int)threadIdx.x < n) {
        double *dst = soa + s * 9 * N + g0 + threadIdx.x;
#pragma unroll
        for (int c = 0; c < 9; ++c) dst[c * N] = tile[threadIdx.x * 9 + c];
    }
}

__global__ void __launch_bounds__(LAYOUT_TILE) unpack_soa_kernel(const double *__restrict__ soa,
                                                                 double *__restrict__ aos, int64_t S,
                                                                 int64_t N) {
    __shared__ double tile[LAYOUT_TILE * 9];
    const int64_t s = blockIdx.x, g0 = (int64_t)blockIdx.y * LAYOUT_TILE;
    const int n = (int)((N - g0 < LAYOUT_TILE) ? (N - g0) : LAYOUT_TILE);
    if ((int)threadIdx.x < n) {
        const double *src = soa + s * 9 * N + g0 + threadIdx.x;
#pragma unroll
        for (int c = 0; c < 9; ++c) tile[threadIdx.x * 9 + c] = src[c * N];
    }
    __syncthreads();
    double *dst = aos + (s * N + g0) * 9;
    for (int i = threadIdx.x; i < n * 9; i += LAYOUT_TILE) dst[i] = tile[i];
}

struct DerivArgs {
    const double *o, *f, *D, *L, *spin, *vol_frac;
    const int32_t *regime, *phase, *fabric;
    double *d_o, *d_f, *energy;
    double gbm_mobility;
    int64_t S, N;
};

// A block walks DERIV_TILES consecutive 128-grain tiles of one system: the math tables are loaded
// once, and the nine loads of the next tile are issued before the current tile is computed, so
// that the memory latency of a tile hides behind the arithmetic of the one before it (one tile
// per block spent 36 % of its stall samples waiting for those loads).
constexpr int DERIV_TILES = 4;
// (eight CTAs per SM, 64 registers with a little spilling: the kernel is bound by memory latency,
// 32 resident warps beat 16 with everything in registers by 6 %)
__global__ void __launch_bounds__(DERIV_BLOCK, 8)
derivatives_grain_kernel(const DerivArgs A, const __grid_constant__ FabricTable T) {
    __shared__ __align__(16) double mt[MATH_TABLE_DOUBLES];
    const int64_t s = blockIdx.x;  // systems in grid.x: no 65535 limit
    const int64_t N = A.N;
    int64_t g = (int64_t)blockIdx.y * (DERIV_BLOCK * DERIV_TILES) + threadIdx.x;
    const int regime = A.regime[s];
    const double *o = A.o + s * 9 * N;
    double *dst = A.d_o + s * 9 * N;
    if (regime == 0 || regime == 7 || regime == 1) {
        // identity "derivative" (core.py:386-391,467-472); regime 1: every grain gets the spin
        // argument (core.py:392-405)
        for (int t = 0; t < DERIV_TILES && g < N; ++t, g += DERIV_BLOCK) {
#pragma unroll
            for (int c = 0; c < 9; ++c)
                dst[c * N + g] = (regime == 1) ? A.spin[s * 9 + c] : ((c % 4 == 0) ? 1.0 : 0.0);
            A.energy[s * N + g] = 0.0;
        }
        return;
    }
    double a[9], D[9], L[9];
    if (g < N) {
#pragma unroll
        for (int c = 0; c < 9; ++c) a[c] = o[c * N + g];
    }
    load_math_tables(mt);
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        D[c] = A.D[s * 9 + c];
        L[c] = A.L[s * 9 + c];
    }
    const FabricConsts &fc = T.fab[A.fabric[s]];
    const double scale = (regime == 6) ? 0.3 : 1.0;  // core.py:459
    __syncthreads();
#pragma unroll 1
    for (int t = 0; t < DERIV_TILES && g < N; ++t, g += DERIV_BLOCK) {
        double an[9];
        const int64_t gn = g + DERIV_BLOCK;
        if (t + 1 < DERIV_TILES && gn < N) {
#pragma unroll
            for (int c = 0; c < 9; ++c) an[c] = o[c * N + gn];
        }
        GrainOut out;
        grain_rhs<false>(a, D, L, fc, mt, out, nullptr);
#pragma unroll
        for (int c = 0; c < 9; ++c) dst[c * N + g] = (regime == 6) ? scale * out.da[c] : out.da[c];
        A.energy[s * N + g] = out.energy;
#pragma unroll
        for (int c = 0; c < 9; ++c) a[c] = an[c];
    }
}

constexpr int GBM_BLOCK = 512;

__global__ void __launch_bounds__(GBM_BLOCK) derivatives_gbm_kernel(const DerivArgs A) {
    __shared__ double red[64];
    const int64_t s = blockIdx.x;
    const int regime = A.regime[s];
    const double *f = A.f + s * A.N;
    const double *E = A.energy + s * A.N;
    double *df = A.d_f + s * A.N;
    if (regime != 4 && regime != 6) {
        for (int64_t g = threadIdx.x; g < A.N; g += GBM_BLOCK) df[g] = 0.0;
        return;
    }
    double acc = 0.0, dummy = 0.0;
    for (int64_t g = threadIdx.x; g < A.N; g += GBM_BLOCK) acc += f[g] * E[g];
    block_sum2<GBM_BLOCK>(acc, dummy, red);
    const double mean = acc;
    const double pref = A.vol_frac[s] * A.gbm_mobility;
    const double scale = (regime == 6) ? 0.3 : 1.0;  // core.py:464
    for (int64_t g = threadIdx.x; g < A.N; g += GBM_BLOCK) {
        const double resid = (regime == 6) ? scale * (mean - E[g]) : (mean - E[g]);
        df[g] = pref * f[g] * resid;
    }
}

struct ProbeArgs {
    const double *o_aos, *D, *L;
    double *inv, *rates, *G, *gamma0, *energy, *da;
    int64_t n;
};

__global__ void __launch_bounds__(128)
grain_probe_kernel(const ProbeArgs A, const __grid_constant__ FabricConsts fc) {
    __shared__ __align__(16) double mt[MATH_TABLE_DOUBLES];
    load_math_tables(mt);
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A.n) return;
    double a[9], D[9], L[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        a[c] = A.o_aos[g * 9 + c];
        D[c] = A.D[c];
        L[c] = A.L[c];
    }
    GrainOut out;
    GrainProbe pr;
    grain_rhs<true>(a, D, L, fc, mt, out, &pr);
    if (A.inv)
        for (int i = 0; i < 4; ++i) A.inv[g * 4 + i] = pr.inv[i];
    if (A.rates)
        for (int i = 0; i < 4; ++i) A.rates[g * 4 + i] = pr.rates[i];
    if (A.G)
        for (int i = 0; i < 9; ++i) A.G[g * 9 + i] = pr.G[i];
    if (A.gamma0) A.gamma0[g] = pr.gamma0;
    if (A.energy) A.energy[g] = out.energy;
    if (A.da)
        for (int i = 0; i < 9; ++i) A.da[g * 9 + i] = out.da[i];
}

// The reference's private helpers on arbitrary inputs (drex_grain_helper_f64); one thread,
// written term by term as the formulas read (core.py:477-684) -- not a hot path.
struct HelperArgs {
    int op;
    const double *in0, *in1, *in2;
    const int32_t *idx;
    double s0, s1, s2, s3;
    double *out;
};

__global__ void grain_helper_kernel(const HelperArgs A) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // slip systems as (direction row, normal row): (010)[100], (001)[100], (010)[001], (100)[001]
    const int dir[4] = {0, 0, 2, 2}, nrm[4] = {1, 2, 1, 0};
    switch (A.op) {
    case 0: {  // I_s = sum_ij D_ij l_i n_j
        const double *D = A.in0, *a = A.in1;
        for (int s = 0; s < 4; ++s) {
            double t = 0.0;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) t += D[3 * i + j] * a[3 * dir[s] + i] * a[3 * nrm[s] + j];
            A.out[s] = t;
        }
    } break;
    case 1: {
        const double *I = A.in0, *crss = A.in1;
        const int i_inac = A.idx[0], i_min = A.idx[1], i_int = A.idx[2], i_max = A.idx[3];
        const double pre = crss[i_max] / I[i_max];
        const double r_min = pre * I[i_min] / crss[i_min];
        const double r_int = pre * I[i_int] / crss[i_int];
        A.out[i_inac] = 0.0;
        A.out[i_min] = r_min * pow(fabs(r_min), A.s0 - 1.0);
        A.out[i_int] = r_int * pow(fabs(r_int), A.s0 - 1.0);
        A.out[i_max] = 1.0;
    } break;
    case 2: {
        const double *a = A.in0, *rates = A.in1;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double t = 0.0;
                for (int s = 0; s < 4; ++s) t += rates[s] * a[3 * dir[s] + i] * a[3 * nrm[s] + j];
                A.out[3 * i + j] = 2.0 * t;
            }
    } break;
    case 3: {
        const double *G = A.in0, *L = A.in1;
        double num = 0.0, den = 0.0;
        for (int j = 0; j < 3; ++j) {
            const int k = (j + 1) % 3;
            num -= (L[3 * j + k] - L[3 * k + j]) * (G[3 * j + k] - G[3 * k + j]);
            den -= (G[3 * j + k] - G[3 * k + j]) * (G[3 * j + k] - G[3 * k + j]);
            for (int l = 0; l < 3; ++l) {
                num += 2.0 * G[3 * j + l] * L[3 * j + l];
                den += 2.0 * G[3 * j + l] * G[3 * j + l];
            }
        }
        A.out[0] = (den > -1e-15 && den < 1e-15) ? 0.0 : num / den;
    } break;
    case 4: {
        const double *a = A.in0, *L = A.in1, *G = A.in2;
        double w[3];
        for (int j = 0; j < 3; ++j) {
            const int r = (j + 1) % 3, s = (j + 2) % 3;
            w[j] = 0.5 * ((L[3 * s + r] - L[3 * r + s]) - (G[3 * s + r] - G[3 * r + s]) * A.s0);
        }
        for (int p = 0; p < 3; ++p) {
            A.out[3 * p + 0] = w[1] * a[3 * p + 2] - w[2] * a[3 * p + 1];
            A.out[3 * p + 1] = w[2] * a[3 * p + 0] - w[0] * a[3 * p + 2];
            A.out[3 * p + 2] = w[0] * a[3 * p + 1] - w[1] * a[3 * p + 0];
        }
    } break;
    case 5: {
        const double *crss = A.in0, *rates = A.in1;
        double E = 0.0;
        for (int i = 0; i < 3; ++i) {
            const double rho = pow(1.0 / crss[i], A.s2 - A.s1) * pow(fabs(rates[i] * A.s0), A.s1 / A.s2);
            E += rho * exp(-A.s3 * rho * rho);
        }
        A.out[0] = E;
    } break;
    default: break;
    }
}

// ---- Fluidity particle attribute layout -------------------------------------------------
// One row per particle, 10 N + 22 doubles (examples/fluidity/corner2d/corner2d.flml:98-175):
//   [ orientations N x 3 x 3 (AoS) | fractions N | strain | position 3 | L_prev 9 | F 9 ]
// unpack: row -> SoA state + the auxiliary fields; pack: the inverse, with the strain
// advanced by |dt| * max|eig sym(L_new)| (utils.strain_increment, utils.py:143-156).
__global__ void __launch_bounds__(256)
fluidity_unpack_kernel(const double *__restrict__ cpo, int64_t N, double *__restrict__ o_soa,
                       double *__restrict__ f, double *__restrict__ F, double *__restrict__ L_prev,
                       double *__restrict__ strain) {
    const int64_t p = blockIdx.x, row = 10 * N + 22;  // particles in grid.x: no 65535 limit
    const int64_t g = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const double *src = cpo + p * row;
    if (g < N) {
#pragma unroll
        for (int c = 0; c < 9; ++c) o_soa[p * 9 * N + c * N + g] = src[9 * g + c];
        f[p * N + g] = src[9 * N + g];
    }
    if (blockIdx.y == 0 && threadIdx.x < 9) {
        L_prev[p * 9 + threadIdx.x] = src[10 * N + 4 + threadIdx.x];
        F[p * 9 + threadIdx.x] = src[10 * N + 13 + threadIdx.x];
        if (threadIdx.x == 0) strain[p] = src[10 * N];
    }
}

__global__ void __launch_bounds__(256)
fluidity_pack_kernel(const double *__restrict__ o_soa, const double *__restrict__ f,
                     const double *__restrict__ F, const double *__restrict__ strain,
                     const double *__restrict__ position, const double *__restrict__ L_new, double dt,
                     int64_t N, double *__restrict__ cpo) {
    const int64_t p = blockIdx.x, row = 10 * N + 22;  // particles in grid.x: no 65535 limit
    const int64_t g = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    double *dst = cpo + p * row;
    if (g < N) {
#pragma unroll
        for (int c = 0; c < 9; ++c) dst[9 * g + c] = o_soa[p * 9 * N + c * N + g];
        dst[9 * N + g] = f[p * N + g];
    }
    if (blockIdx.y == 0 && threadIdx.x < 9) {
        dst[10 * N + 4 + threadIdx.x] = L_new[p * 9 + threadIdx.x];
        dst[10 * N + 13 + threadIdx.x] = F[p * 9 + threadIdx.x];
        if (threadIdx.x < 3) dst[10 * N + 1 + threadIdx.x] = position[p * 3 + threadIdx.x];
        if (threadIdx.x == 0) dst[10 * N] = strain[p] + fabs(dt) * strain_rate_max(L_new + p * 9);
    }
}

}  // namespace drex
