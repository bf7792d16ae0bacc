// drex_diagnostics.cuh -- post-step texture reductions on device (sm_100a).
//
//  voigt_average_kernel   minerals.voigt_averages (src/pydrex/minerals.py:688-740):
//      per grain the rotated stiffness is P S P^T with P the 6x6 matrix of products of
//      direction cosines (equivalent to tensors.rotate + elastic_tensor_to_voigt,
//      tensors.py:187-273, for tensors with the minor symmetries voigt_to_elastic_tensor
//      builds), ~500 FP64 FMA instead of the reference's 3^8-term loop; one CTA per system,
//      21 accumulators per thread, fixed-order block reduction.  At 80 B and ~1000 flop per
//      grain the kernel is FP64-bound, not HBM-bound (measured 59 % of the DFMA peak).
//  mindex_*               diagnostics.misorientation_index (diagnostics.py:332-367) with
//      stats.misorientation_hist (stats.py:83-127) and geometry.misorientation_angles
//      (geometry.py:74-118): all pairs x all symmetry-operator pairs, tiled on chip; the
//      [C(N,2), 7, 4] float32 pair arrays of the reference (7.7 GB at N = 5000) never
//      exist.  Arithmetic mirrors the reference: float32 transformed quaternions, float32
//      products and sums without contraction, min over operator pairs == acos of the
//      largest |dot| (acos is monotone).
#pragma once
#include "drex_common.cuh"

namespace drex {

// ------------------------------------------------------------------ Voigt average

constexpr int VOIGT_BLOCK = 128;  // three CTAs of 168 registers per SM: 12 warps instead of 8 (-6 %)

// canonical (p, q) of Voigt index I: inverse of tensors.py:174-175
__host__ __device__ constexpr int voigt_p(int I) { return I < 3 ? I : (I == 3 ? 1 : 0); }
__host__ __device__ constexpr int voigt_q(int I) { return I < 3 ? I : (I == 5 ? 1 : 2); }

__global__ void __launch_bounds__(VOIGT_BLOCK, 3)
voigt_average_kernel(const double *__restrict__ o_soa, const double *__restrict__ f,
                     const double *__restrict__ stiffness, const double *__restrict__ phase_fraction,
                     int64_t N, double *__restrict__ Cbar, int accumulate) {
    __shared__ double Ssym[36];
    __shared__ double red[21][VOIGT_BLOCK / 32];
    const int64_t s = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid < 36) {
        const int i = tid / 6, j = tid % 6;
        // (M + M^T)/2 commutes with the congruence P . P^T (tensors.py:203)
        Ssym[tid] = 0.5 * (stiffness[s * 36 + 6 * i + j] + stiffness[s * 36 + 6 * j + i]);
    }
    __syncthreads();
    double acc[21];
#pragma unroll
    for (int i = 0; i < 21; ++i) acc[i] = 0.0;
    const double *o = o_soa + s * 9 * N;
    for (int64_t g = tid; g < N; g += VOIGT_BLOCK) {
        double a[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) a[c] = o[c * N + g];
        const double w = f[s * N + g];
        // rotation R = a^T (minerals.py:735): R[i][k] = a[k][i]
        double Pm[6][6];
#pragma unroll
        for (int I = 0; I < 6; ++I) {
            const int p = voigt_p(I), q = voigt_q(I);
#pragma unroll
            for (int A = 0; A < 6; ++A) {
                const int k = voigt_p(A), l = voigt_q(A);
                const double rpk = a[3 * k + p], rql = a[3 * l + q];
                Pm[I][A] = (A < 3) ? rpk * rql : rpk * rql + a[3 * l + p] * a[3 * k + q];
            }
        }
        int idx = 0;
#pragma unroll
        for (int I = 0; I < 6; ++I) {
            double Q[6];  // row I of P S: only one row is live at a time (register pressure)
#pragma unroll
            for (int B = 0; B < 6; ++B) {
                double t = 0.0;
#pragma unroll
                for (int A = 0; A < 6; ++A) t += Pm[I][A] * Ssym[6 * A + B];
                Q[B] = t;
            }
#pragma unroll
            for (int J = I; J < 6; ++J) {
                double t = 0.0;
#pragma unroll
                for (int B = 0; B < 6; ++B) t += Q[B] * Pm[J][B];
                acc[idx++] += w * t;
            }
        }
    }
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < 21; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) red[i][warp] = v;
    }
    __syncthreads();
    if (tid < 21) {
        double t = 0.0;
        for (int w = 0; w < VOIGT_BLOCK / 32; ++w) t += red[tid][w];
        t *= phase_fraction[s];
        // unpack upper-triangular index
        int I = 0, rem = tid;
        while (rem >= 6 - I) {
            rem -= 6 - I;
            ++I;
        }
        const int J = I + rem;
        double *out = Cbar + s * 36;
        if (accumulate) {
            out[6 * I + J] += t;
            if (I != J) out[6 * J + I] += t;
        } else {
            out[6 * I + J] = t;
            if (I != J) out[6 * J + I] = t;
        }
    }
}

// ------------------------------------------------------------------ elasticity decomposition

// Symmetric 3x3 eigen-decomposition, eigenvalues ascending, eigenvectors in the COLUMNS of V
// (what scipy.linalg.eigh returns, up to the sign of each column).
__device__ inline void eigh3_ascending(const double *A, double *w, double *V) {
    jacobi_eig3(A, w, V);
    for (int i = 0; i < 2; ++i)  // selection sort of 3 values, swapping columns
        for (int j = i + 1; j < 3; ++j)
            if (w[j] < w[i]) {
                const double t = w[i]; w[i] = w[j]; w[j] = t;
                for (int r = 0; r < 3; ++r) {
                    const double u = V[3 * r + i]; V[3 * r + i] = V[3 * r + j]; V[3 * r + j] = u;
                }
            }
}

// voigt(rotate(C, a^T)) for a symmetric Voigt matrix S, all 36 entries (tensors.py:187-273)
__device__ inline void rotate_voigt_full(const double *a, const double *S, double *out) {
    double Pm[6][6], Q[6][6];
    for (int I = 0; I < 6; ++I) {
        const int p = voigt_p(I), q = voigt_q(I);
        for (int A = 0; A < 6; ++A) {
            const int k = voigt_p(A), l = voigt_q(A);
            const double rpk = a[3 * k + p], rql = a[3 * l + q];
            Pm[I][A] = (A < 3) ? rpk * rql : rpk * rql + a[3 * l + p] * a[3 * k + q];
        }
    }
    for (int I = 0; I < 6; ++I)
        for (int B = 0; B < 6; ++B) {
            double t = 0.0;
            for (int A = 0; A < 6; ++A) t += Pm[I][A] * S[6 * A + B];
            Q[I][B] = t;
        }
    for (int I = 0; I < 6; ++I)
        for (int J = 0; J < 6; ++J) {
            double t = 0.0;
            for (int B = 0; B < 6; ++B) t += Q[I][B] * Pm[J][B];
            out[6 * I + J] = t;
        }
}

// tensors.voigt_matrix_to_vector (tensors.py:206-219)
__device__ inline void voigt_to_vector21(const double *m, double *v) {
    const double r2 = 1.4142135623730951;
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        v[i] = m[6 * i + i];
        v[i + 3] = r2 * m[6 * j + k];
        v[i + 6] = 2.0 * m[6 * (i + 3) + i + 3];
        v[i + 9] = 2.0 * m[6 * i + i + 3];
        v[i + 12] = 2.0 * m[6 * k + i + 3];
        v[i + 15] = 2.0 * m[6 * j + i + 3];
        v[i + 18] = 2.0 * r2 * m[6 * (j + 3) + k + 3];
    }
}

__device__ inline double norm_n(const double *v, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i] * v[i];
    return sqrt(s);
}

// diagnostics.elasticity_components (diagnostics.py:38-182; Browaeys & Chevrot 2004), one
// thread per Voigt matrix.  out[m][0..10] = bulk modulus, shear modulus, % anisotropy,
// % hexagonal, % tetragonal, % orthorhombic, % monoclinic, % triclinic, hexagonal axis (3).
// Follows the reference including its quirks: the "nearest eigenvector" search only accepts
// angles below 10 degrees (the initial value of `angle` is 10 while smallest_angle returns
// degrees, diagnostics.py:112) and the signed index multiplies the eigenvector
// (diagnostics.py:123-131).
__global__ void __launch_bounds__(64)
elasticity_components_kernel(const double *__restrict__ voigt, int64_t P, double *__restrict__ out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= P) return;
    double C[36];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j)  // tensors.upper_tri_to_symmetric (tensors.py:143-164)
            C[6 * i + j] = (j >= i) ? voigt[m * 36 + 6 * i + j] : voigt[m * 36 + 6 * j + i];
    // tensors.voigt_decompose (tensors.py:39-73)
    double d[9], v[9];
    for (int i = 0; i < 3; ++i) d[3 * i + i] = C[i] + C[6 + i] + C[12 + i];
    d[1] = d[3] = C[5] + C[11] + C[17];
    d[2] = d[6] = C[4] + C[10] + C[16];
    d[5] = d[7] = C[3] + C[9] + C[15];
    v[0] = C[0] + C[28] + C[35];
    v[4] = C[7] + C[21] + C[35];
    v[8] = C[14] + C[21] + C[28];
    v[1] = v[3] = C[5] + C[11] + C[22];
    v[2] = v[6] = C[4] + C[16] + C[23];
    v[5] = v[7] = C[9] + C[15] + C[29];
    const double K = (d[0] + d[4] + d[8]) / 9.0;
    const double G = ((v[0] + v[4] + v[8]) - 3.0 * K) / 10.0;
    double *o = out + m * 11;
    o[0] = K;
    o[1] = G;
    double iso[21], X[21], tmp[21];
    for (int i = 0; i < 21; ++i) iso[i] = 0.0;
    for (int i = 0; i < 3; ++i) {
        iso[i] = K + 4.0 * G / 3.0;
        iso[i + 3] = 1.4142135623730951 * (K - 2.0 * G / 3.0);
        iso[i + 6] = 2.0 * G;
    }
    voigt_to_vector21(C, X);
    const double normX = norm_n(X, 21);
    for (int i = 0; i < 21; ++i) tmp[i] = X[i] - iso[i];
    o[2] = norm_n(tmp, 21) / normX * 100.0;
    for (int i = 3; i < 11; ++i) o[i] = NAN;  // the reference leaves these unset if no permutation improves

    double wd[3], Vd[9], wv[3], Vv[9], sccs[9];
    eigh3_ascending(d, wd, Vd);
    eigh3_ascending(v, wv, Vv);
    for (int i = 0; i < 3; ++i) {
        double index_vij = 0.0, angle = 10.0;
        for (int j = 0; j < 3; ++j) {
            double dot = 0.0, nd = 0.0, nv = 0.0;
            for (int r = 0; r < 3; ++r) {
                dot += Vd[3 * r + i] * Vv[3 * r + j];
                nd += Vd[3 * r + i] * Vd[3 * r + i];
                nv += Vv[3 * r + j] * Vv[3 * r + j];
            }
            double ang = acos(fmin(1.0, fmax(-1.0, dot / (sqrt(nd) * sqrt(nv))))) * 57.29577951308232;
            if (ang > 90.0) ang = 180.0 - ang;  // diagnostics.smallest_angle (diagnostics.py:386-430)
            if (ang < angle) {
                angle = ang;
                index_vij = (dot != 0.0) ? ((dot > 0.0) ? (double)j : -(double)j) : (double)j;
            }
        }
        const int jj = (int)fabs(index_vij);
        double vec[3], nrm = 0.0;
        for (int r = 0; r < 3; ++r) {
            vec[r] = (Vd[3 * r + i] + index_vij * Vv[3 * r + jj]) / 2.0;
            nrm += vec[r] * vec[r];
        }
        nrm = sqrt(nrm);
        for (int r = 0; r < 3; ++r) sccs[3 * r + i] = vec[r] / nrm;
    }
    double distance = normX;
    for (int i = 0; i < 3; ++i) {
        double perm[9], rot[36], Y[21], mono[21], ortho[21], tetr[21], hexv[21];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 3; ++j) perm[3 * r + j] = sccs[3 * r + (i + j) % 3];
        rotate_voigt_full(perm, C, rot);
        voigt_to_vector21(rot, Y);
        for (int k = 0; k < 21; ++k) mono[k] = Y[k];  // tensors.mono_project (tensors.py:76-89)
        mono[9] = mono[10] = mono[12] = mono[13] = mono[15] = mono[16] = mono[18] = mono[19] = 0.0;
        for (int k = 0; k < 21; ++k) ortho[k] = (k < 9) ? mono[k] : 0.0;  // ortho_project
        for (int k = 0; k < 21; ++k) tetr[k] = ortho[k];                  // tetr_project
        tetr[0] = tetr[1] = 0.5 * (ortho[0] + ortho[1]);
        tetr[3] = tetr[4] = 0.5 * (ortho[3] + ortho[4]);
        tetr[6] = tetr[7] = 0.5 * (ortho[6] + ortho[7]);
        for (int k = 0; k < 21; ++k) hexv[k] = 0.0;                       // hex_project
        const double r2 = 1.4142135623730951;
        hexv[0] = hexv[1] = 3.0 / 8.0 * (tetr[0] + tetr[1]) + tetr[5] / 4.0 / r2 + tetr[8] / 4.0;
        hexv[2] = tetr[2];
        hexv[3] = hexv[4] = (tetr[3] + tetr[4]) / 2.0;
        hexv[5] = (tetr[0] + tetr[1]) / 4.0 / r2 + 3.0 / 4.0 * tetr[5] - tetr[8] / 2.0 / r2;
        hexv[6] = hexv[7] = (tetr[6] + tetr[7]) / 2.0;
        hexv[8] = (tetr[0] + tetr[1]) / 4.0 - tetr[5] / 2.0 / r2 + tetr[8] / 2.0;
        for (int k = 0; k < 21; ++k) tmp[k] = Y[k] - hexv[k];
        const double delta = norm_n(tmp, 21);
        if (delta < distance) {
            distance = delta;
            const double pct = 100.0 / normX;
            for (int k = 0; k < 21; ++k) tmp[k] = Y[k] - mono[k];
            o[7] = norm_n(tmp, 21) * pct;  // triclinic
            for (int k = 0; k < 21; ++k) tmp[k] = mono[k] - ortho[k];
            o[6] = norm_n(tmp, 21) * pct;  // monoclinic
            for (int k = 0; k < 21; ++k) tmp[k] = ortho[k] - tetr[k];
            o[5] = norm_n(tmp, 21) * pct;  // orthorhombic
            for (int k = 0; k < 21; ++k) tmp[k] = tetr[k] - hexv[k];
            o[4] = norm_n(tmp, 21) * pct;  // tetragonal
            for (int k = 0; k < 21; ++k) tmp[k] = hexv[k] - iso[k];
            o[3] = norm_n(tmp, 21) * pct;  // hexagonal
            for (int r = 0; r < 3; ++r) o[8 + r] = perm[3 * r + 2];  // last SCCS axis
        }
    }
}

// ------------------------------------------------------------------ axis statistics

// Scatter (inertia) matrix of one crystallographic axis over a texture
// (stats._scatter_matrix, stats.py:67-80), its eigen-decomposition, and from it
// diagnostics.symmetry_pgr (diagnostics.py:233-270) and diagnostics.bingham_average
// (diagnostics.py:185-213).  One CTA per texture.  out[s][0..8] = P, G, R, mean axis (3, the
// eigenvector of the largest eigenvalue, unit length, sign arbitrary), eigenvalues descending.
__global__ void __launch_bounds__(256)
axis_scatter_kernel(const double *__restrict__ o_soa, int64_t N, int row, double *__restrict__ out) {
    __shared__ double red[6][8];
    const int64_t s = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *o = o_soa + s * 9 * N + (int64_t)(3 * row) * N;
    double acc[6] = {0, 0, 0, 0, 0, 0};  // xx, yy, zz, xy, xz, yz
    for (int64_t g = tid; g < N; g += 256) {
        const double x = o[g], y = o[N + g], z = o[2 * N + g];
        acc[0] += x * x; acc[1] += y * y; acc[2] += z * z;
        acc[3] += x * y; acc[4] += x * z; acc[5] += y * z;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) red[i][warp] = v;
    }
    __syncthreads();
    if (tid == 0) {
        double t[6];
        for (int i = 0; i < 6; ++i) {
            double v = 0.0;
            for (int w = 0; w < 8; ++w) v += red[i][w];
            t[i] = v;
        }
        const double A[9] = {t[0], t[3], t[4], t[3], t[1], t[5], t[4], t[5], t[2]};
        double w[3], V[9];
        eigh3_ascending(A, w, V);
        const double sum = w[0] + w[1] + w[2];
        double *o9 = out + s * 9;
        o9[0] = (w[2] - w[1]) / sum;
        o9[1] = 2.0 * (w[1] - w[0]) / sum;
        o9[2] = 3.0 * w[0] / sum;
        const double nrm = sqrt(V[2] * V[2] + V[5] * V[5] + V[8] * V[8]);
        o9[3] = V[2] / nrm; o9[4] = V[5] / nrm; o9[5] = V[8] / nrm;
        o9[6] = w[2]; o9[7] = w[1]; o9[8] = w[0];
    }
}

// ------------------------------------------------------------------ M-index

constexpr int MAX_SYMOPS = 16;
constexpr int MINDEX_TILE = 128;  // grains per tile side == threads per CTA
constexpr int MAX_THETA = 180;

struct SymOps {
    double m[MAX_SYMOPS][16];  // 4x4 matrices acting on scalar-last quaternions
    int n;
};

// scipy Rotation.from_matrix(...).as_quat() (third-party; scipy 1.18.1
// spatial/transform/_rotation_xp.py:51-157): nearest rotation for non-orthogonal input,
// then the largest-of-diagonal-and-trace branch and normalisation.
__device__ inline void matrix_to_quat(const double *M_in, double *q) {
    double X[9];
    for (int i = 0; i < 9; ++i) X[i] = M_in[i];
    for (int it = 0; it < 40; ++it) {  // X <- (X + X^-T)/2 converges to the polar factor
        const double c0[3] = {X[4] * X[8] - X[5] * X[7], X[5] * X[6] - X[3] * X[8], X[3] * X[7] - X[4] * X[6]};
        const double c1[3] = {X[7] * X[2] - X[8] * X[1], X[8] * X[0] - X[6] * X[2], X[6] * X[1] - X[7] * X[0]};
        const double c2[3] = {X[1] * X[5] - X[2] * X[4], X[2] * X[3] - X[0] * X[5], X[0] * X[4] - X[1] * X[3]};
        const double det = X[0] * c0[0] + X[1] * c0[1] + X[2] * c0[2];
        const double inv_t[9] = {c0[0] / det, c0[1] / det, c0[2] / det, c1[0] / det, c1[1] / det,
                                 c1[2] / det, c2[0] / det, c2[1] / det, c2[2] / det};
        double delta = 0.0;
        for (int i = 0; i < 9; ++i) {
            const double nx = 0.5 * (X[i] + inv_t[i]);
            delta = fmax(delta, fabs(nx - X[i]));
            X[i] = nx;
        }
        if (delta < 1e-16) break;
    }
    const double tr = X[0] + X[4] + X[8];
    int choice = 0;
    double best = X[0];
    if (X[4] > best) { best = X[4]; choice = 1; }
    if (X[8] > best) { best = X[8]; choice = 2; }
    if (tr > best) { choice = 3; }
    if (choice == 3) {
        q[0] = X[7] - X[5];
        q[1] = X[2] - X[6];
        q[2] = X[3] - X[1];
        q[3] = 1.0 + tr;
    } else {
        const int i = choice, j = (i + 1) % 3, k = (j + 1) % 3;
        q[i] = 1.0 - tr + 2.0 * X[4 * i];
        q[j] = X[3 * j + i] + X[3 * i + j];
        q[k] = X[3 * k + i] + X[3 * i + k];
        q[3] = X[3 * k + j] - X[3 * j + k];
    }
    const double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= nrm;
}

// K1: transformed float32 quaternions, layout [S][N][n_store] float4
__global__ void __launch_bounds__(128)
mindex_quats_kernel(const double *__restrict__ o_soa, int64_t N, float4 *__restrict__ quats,
                    int n_store, const __grid_constant__ SymOps ops) {
    const int64_t s = blockIdx.x;  // textures in grid.x: no 65535 limit
    const int64_t g = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (g >= N) return;
    double a[9], q[4];
    for (int c = 0; c < 9; ++c) a[c] = o_soa[s * 9 * N + c * N + g];
    matrix_to_quat(a, q);
    float4 *dst = quats + (s * N + g) * n_store;  // the sign-flip ops are not materialised
    for (int k = 0; k < n_store; ++k) {
        double r[4];
        for (int i = 0; i < 4; ++i) {
            double t = 0.0;
            for (int j = 0; j < 4; ++j) t += ops.m[k][4 * i + j] * q[j];  // stats.py:119-123
            r[i] = t;
        }
        dst[k] = make_float4((float)r[0], (float)r[1], (float)r[2], (float)r[3]);  // stats.py:105-112
    }
}

// Bin index of a pair from its largest |dot| WITHOUT evaluating acos per pair.
// The reference's chain is float32: a = acosf(d); h = a * rad2deg (float32); angle = 2 h
// (float64); bin = floor(angle) (np.histogram with 1-degree bins).  Every step is monotone
// in d, so for each integer b there is a float32 threshold thr[b] with angle >= b  <=>
// d <= thr[b].  The host computes the thresholds with ITS libm acosf (the function numpy
// and numba call), which makes the binning identical to the reference's, not merely
// close; the device does a 8-step binary search over at most 181 thresholds.
// thr[1..theta_max] as above; thr[theta_max + 1]: angle > theta_max (dropped, stats.py:127).
__device__ __forceinline__ int mindex_bin(float d, const float *thr, int theta_max) {
    // k = number of b in [1, theta_max + 1] with d <= thr[b]; thr is non-increasing in b
    int lo = 0, hi = theta_max + 1;  // invariant: d <= thr[b] for b <= lo, d > thr[b] for b > hi
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (d <= thr[mid]) lo = mid; else hi = mid - 1;
    }
    if (lo > theta_max) return -1;             // outside the histogram range
    return lo == theta_max ? theta_max - 1 : lo;  // right edge belongs to the last bin
}

// decode (ti, tj), ti <= tj, from the linear upper-triangular tile index
__device__ __forceinline__ void mindex_tile(int linear, int n_tiles, int &ti, int &tj) {
    ti = 0;
    int rem = linear;
    while (rem >= n_tiles - ti) {
        rem -= n_tiles - ti;
        ++ti;
    }
    tj = ti + rem;
}

// float32 product-then-sum exactly as numpy/numba evaluate np.sum(q1 * q2, axis=1) on
// float32 arrays (geometry.py:104-112): rounded products, left-to-right rounded sums.
__device__ __forceinline__ float dot4_ref(const float4 a, const float4 b) {
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z)),
                     __fmul_rn(a.w, b.w));
}

// The four dots (R a).b for R in {I, diag(1,-1,-1,1), diag(1,-1,1,-1), diag(1,1,-1,-1)}: sign
// flips are exact in float32, so the four sums share their products and first partial sums
// and are bit-identical to the reference's separately transformed quaternions.
__device__ __forceinline__ float absmax_sign4(const float4 a, const float4 b, float best) {
    const float p0 = __fmul_rn(a.x, b.x), p1 = __fmul_rn(a.y, b.y);
    const float p2 = __fmul_rn(a.z, b.z), p3 = __fmul_rn(a.w, b.w);
    const float sp = __fadd_rn(p0, p1), sm = __fsub_rn(p0, p1);
    const float dI = __fadd_rn(__fadd_rn(sp, p2), p3);
    const float d1 = __fadd_rn(__fsub_rn(sm, p2), p3);
    const float d2 = __fsub_rn(__fadd_rn(sm, p2), p3);
    const float d3 = __fsub_rn(__fsub_rn(sp, p2), p3);
    return fmaxf(fmaxf(fmaxf(best, fabsf(dI)), fmaxf(fabsf(d1), fabsf(d2))), fabsf(d3));
}

// K2 (generic): all NOPS x NOPS operator pairs.  grid.x = tile pairs (ti <= tj), grid.y = system.
template <int NOPS>
__global__ void __launch_bounds__(MINDEX_TILE)
mindex_pairs_kernel(const float4 *__restrict__ quats, int64_t N, int n_tiles, int theta_max,
                    const float *__restrict__ thresholds, unsigned long long *__restrict__ counts) {
    __shared__ float4 qj[MINDEX_TILE * NOPS];
    __shared__ unsigned int hist[MAX_THETA];
    __shared__ float thr[MAX_THETA + 2];
    const int64_t s = blockIdx.x;  // textures in grid.x, tile pairs in grid.y
    int ti, tj;
    mindex_tile(blockIdx.y, n_tiles, ti, tj);
    const int tid = threadIdx.x;
    for (int b = tid; b < theta_max; b += MINDEX_TILE) hist[b] = 0u;
    for (int b = tid; b < theta_max + 2; b += MINDEX_TILE) thr[b] = thresholds[b];
    const float4 *base = quats + s * N * NOPS;
    const int64_t j0 = (int64_t)tj * MINDEX_TILE;
    const int nj = (int)min((int64_t)MINDEX_TILE, N - j0);
    for (int e = tid; e < nj * NOPS; e += MINDEX_TILE) qj[e] = base[j0 * NOPS + e];
    const int64_t i = (int64_t)ti * MINDEX_TILE + tid;
    float4 qi[NOPS];
    if (i < N) {
#pragma unroll
        for (int k = 0; k < NOPS; ++k) qi[k] = base[i * NOPS + k];
    }
    __syncthreads();
    if (i < N) {
        const int jstart = (ti == tj) ? tid + 1 : 0;  // pairs (i, j) with i < j only (stats.py:113-115)
        for (int jj = jstart; jj < nj; ++jj) {
            float best = 0.0f;  // max |dot| over operator pairs == min angle (geometry.py:102-118)
#pragma unroll
            for (int a = 0; a < NOPS; ++a)
#pragma unroll
                for (int b = 0; b < NOPS; ++b) best = fmaxf(best, fabsf(dot4_ref(qi[a], qj[jj * NOPS + b])));
            const int bin = mindex_bin(fminf(best, 1.0f), thr, theta_max);  // clip (geometry.py:108)
            if (bin >= 0) atomicAdd(&hist[bin], 1u);
        }
    }
    __syncthreads();
    for (int b = tid; b < theta_max; b += MINDEX_TILE)
        if (hist[b]) atomicAdd(&counts[s * theta_max + b], (unsigned long long)hist[b]);
}

// ---- packed float32 arithmetic (sm_100: FMUL2 / FADD2 / FFMA2 on 64-bit register pairs) ------
// Two pairs of grains per instruction.  Every lane of a packed operation is an IEEE
// round-to-nearest float32 operation, so results are bit-identical to the scalar chain -- as
// long as no product is contracted into the sum that follows it.  ptxas (12.9) contracts
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with the explicit rounding modifiers, so the
// products are formed as fma(a, b, -0.0f) with the -0.0f pair coming in as a kernel argument:
// the value is exactly a * b (x + -0 = x for every x, signed zeros included), costs the same
// one instruction, and an FFMA2 cannot be folded into the next add.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b, f32x2_t neg_zero) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(neg_zero));
    return r;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
struct Quat2 {
    f32x2_t x, y, z, w;  // one component of two grains (or one grain's, duplicated) per member
};
__device__ __forceinline__ void absmax2(f32x2_t d, float &b0, float &b1) {
    float lo, hi;
    unpack2(d, lo, hi);
    b0 = fmaxf(b0, fabsf(lo));
    b1 = fmaxf(b1, fabsf(hi));
}
// absmax_sign4 for two grains at once
__device__ __forceinline__ void absmax_sign4_x2(const Quat2 &a, const Quat2 &b, f32x2_t nz, float &b0, float &b1) {
    const f32x2_t p0 = mul2(a.x, b.x, nz), p1 = mul2(a.y, b.y, nz);
    const f32x2_t p2 = mul2(a.z, b.z, nz), p3 = mul2(a.w, b.w, nz);
    const f32x2_t sp = add2(p0, p1), sm = sub2(p0, p1);
    absmax2(add2(add2(sp, p2), p3), b0, b1);
    absmax2(add2(sub2(sm, p2), p3), b0, b1);
    absmax2(sub2(add2(sm, p2), p3), b0, b1);
    absmax2(sub2(sub2(sp, p2), p3), b0, b1);
}
// dot4_ref for two grains at once
__device__ __forceinline__ void absdot_x2(const Quat2 &a, const Quat2 &b, f32x2_t nz, float &b0, float &b1) {
    const f32x2_t p0 = mul2(a.x, b.x, nz), p1 = mul2(a.y, b.y, nz);
    const f32x2_t p2 = mul2(a.z, b.z, nz), p3 = mul2(a.w, b.w, nz);
    absmax2(add2(add2(add2(p0, p1), p2), p3), b0, b1);
}

// The bin of mindex_bin through a table: cell c = floor(d * MINDEX_LUT) starts the count at the
// count of the cell's upper edge (a lower bound: the count does not increase with d) and walks
// up the thresholds from there: exact, and 0-1 steps except for the smallest angles.  The host
// builds the table next to the thresholds (drex_mindex_f32); a CTA copies both to shared memory.
constexpr int MINDEX_LUT = 2048;
struct MindexBins {
    float thr[MAX_THETA + 4];              // thr[theta_max + 2] = -1 ends the walk
    unsigned short lut[MINDEX_LUT + 8];    // lut[c] = count((c + 1) / MINDEX_LUT), c = 0..MINDEX_LUT
};
static_assert(sizeof(MindexBins) % 16 == 0, "copied as uint4");
__host__ __device__ inline int mindex_count(float d, const float *thr, int theta_max) {
    int lo = 0, hi = theta_max + 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (d <= thr[mid]) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ void mindex_load_bins(MindexBins &dst, const MindexBins *__restrict__ src, int tid,
                                                 int nthreads) {
    const uint4 *a = reinterpret_cast<const uint4 *>(src);
    uint4 *b = reinterpret_cast<uint4 *>(&dst);
    for (int i = tid; i < (int)(sizeof(MindexBins) / 16); i += nthreads) b[i] = a[i];
}
__device__ __forceinline__ int mindex_bin_lut(float d, const MindexBins &bins, int theta_max) {
    int k = bins.lut[(int)(d * (float)MINDEX_LUT)];  // d in [0, 1]: the product is exact
    while (d <= bins.thr[k + 1]) ++k;
    if (k > theta_max) return -1;
    return k == theta_max ? theta_max - 1 : k;
}

// K2 (monoclinic / orthorhombic): the 7 operations are identity, three "rotations" (stored
// per grain as slots 1..3) and three diagonal sign matrices (geometry.py:133-150).  The
// sign matrices with the identity form a group of exact float32 sign flips, so the 49
// operator pairs collapse to 4 + 12 + 12 shared-product dots and 9 plain ones: 161 float
// operations per pair instead of 343, with bit-identical maxima.  Slots per grain: 4.
// A thread holds grain i (every component duplicated into a register pair) and walks the j
// tile two grains at a time with packed arithmetic; the tile sits in shared memory as
// [j / 2][slot][component][j % 2], 128 bytes per pair of grains, read as broadcast LDS.128.
// The kernel is bound by instruction issue: ~115 instructions per pair this way, ~205 scalar.
__global__ void __launch_bounds__(MINDEX_TILE)
mindex_pairs_reflect_kernel(const float4 *__restrict__ quats, int64_t N, int n_tiles, int theta_max,
                            const MindexBins *__restrict__ bins_global,
                            unsigned long long *__restrict__ counts, const f32x2_t neg_zero) {
    __shared__ __align__(16) float qj[MINDEX_TILE * 16];
    __shared__ unsigned int hist[MAX_THETA];
    __shared__ __align__(16) MindexBins bins;
    const int64_t s = blockIdx.x;  // textures in grid.x, tile pairs in grid.y
    int ti, tj;
    mindex_tile(blockIdx.y, n_tiles, ti, tj);
    const int tid = threadIdx.x;
    for (int b = tid; b < theta_max; b += MINDEX_TILE) hist[b] = 0u;
    mindex_load_bins(bins, bins_global, tid, MINDEX_TILE);
    const float4 *base = quats + s * N * 4;
    const int64_t j0 = (int64_t)tj * MINDEX_TILE;
    const int nj = (int)min((int64_t)MINDEX_TILE, N - j0);
    for (int e = tid; e < MINDEX_TILE * 4; e += MINDEX_TILE) {
        const int j = e >> 2, slot = e & 3;
        const float4 v = (j < nj) ? base[j0 * 4 + e] : make_float4(0.f, 0.f, 0.f, 0.f);
        float *dst = qj + (j >> 1) * 32 + slot * 8 + (j & 1);
        dst[0] = v.x;
        dst[2] = v.y;
        dst[4] = v.z;
        dst[6] = v.w;
    }
    const int64_t i = (int64_t)ti * MINDEX_TILE + tid;
    Quat2 qi[4];
    {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 v = (i < N) ? base[i * 4 + k] : z;
            qi[k].x = pack2(v.x, v.x);
            qi[k].y = pack2(v.y, v.y);
            qi[k].z = pack2(v.z, v.z);
            qi[k].w = pack2(v.w, v.w);
        }
    }
    __syncthreads();
    if (i < N) {
        const int jstart = (ti == tj) ? tid + 1 : 0;
        const ulonglong2 *tile = reinterpret_cast<const ulonglong2 *>(qj);
        for (int jp = jstart >> 1; 2 * jp < nj; ++jp) {
            Quat2 qb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const ulonglong2 xy = tile[jp * 8 + 2 * k], zw = tile[jp * 8 + 2 * k + 1];
                qb[k].x = xy.x;
                qb[k].y = xy.y;
                qb[k].z = zw.x;
                qb[k].w = zw.y;
            }
            float b0 = 0.0f, b1 = 0.0f;
            absmax_sign4_x2(qi[0], qb[0], neg_zero, b0, b1);  // {I, R} x {I, R}
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                absmax_sign4_x2(qi[0], qb[k], neg_zero, b0, b1);  // {I, R} q1 . rotations of q2
                absmax_sign4_x2(qi[k], qb[0], neg_zero, b0, b1);  // rotations of q1 . {I, R} q2
            }
#pragma unroll
            for (int a = 1; a < 4; ++a)
#pragma unroll
                for (int b = 1; b < 4; ++b) absdot_x2(qi[a], qb[b], neg_zero, b0, b1);
            const int j = 2 * jp;
            if (j >= jstart) {
                const int bin = mindex_bin_lut(fminf(b0, 1.0f), bins, theta_max);
                if (bin >= 0) atomicAdd(&hist[bin], 1u);
            }
            if (j + 1 < nj) {  // (j + 1 >= jstart always: the walk starts at the pair of jstart)
                const int bin = mindex_bin_lut(fminf(b1, 1.0f), bins, theta_max);
                if (bin >= 0) atomicAdd(&hist[bin], 1u);
            }
        }
    }
    __syncthreads();
    for (int b = tid; b < theta_max; b += MINDEX_TILE)
        if (hist[b]) atomicAdd(&counts[s * theta_max + b], (unsigned long long)hist[b]);
}

// K2 (monoclinic / orthorhombic), the product kernel: only the products that can matter.
// The three stored "rotations" are the reference's quat_product (utils.py:365-371, no cross term)
// with the quaternions of the half turns about z, y, x, whose scalar part is cos(pi/2) =
// 6.1e-17 in double: slot 1 is (6e-17 x, 6e-17 y, w + 6e-17 z, -z + 6e-17 w) and likewise
// slot 2 = (~0, w', ~0, -y'), slot 3 = (w', ~0, ~0, -x'): two components of order 1 ("live")
// and two of order 1e-17.  In a dot ((p0 + p1) + p2) + p3 the products with a 1e-17 factor
// vanish in the rounding of the first live product they meet (they are below half an ulp of
// anything larger than 2e-9), so
//   {I, R} q1 . slot k of q2, and slot k of q1 . {I, R} q2:  |pa + pb| and |pb - pa| over the two
//                                                            live components (the four sign
//                                                            variants are these two, twice);
//   slot a . slot a:  pa + pb over the two live components;
//   slot a . slot b, a != b:  only the scalar parts are live in both: |w_a| |w_b|;
// and the {I, R} q1 . {I, R} q2 block as in the kernel above: 25 products and 25 sums per pair
// instead of 64 and 97.  The maxima are those of the full evaluation unless the WINNING dot
// (always >= cos(60 deg)) contains a live product below 2e-9 and sits within 1e-16 of a
// float32 rounding boundary -- once in ~1e17 pairs; mindex_pairs_reflect_kernel (all 49 dots,
// DREX_MINDEX_FULL=1) is the cross-check and tests/test_gpu_parity.py compares the histograms
// of the two kernels for equality.  Tile in shared memory: [j / 2][10 components][j % 2].
__global__ void __launch_bounds__(MINDEX_TILE)
mindex_pairs_reflect_live_kernel(const float4 *__restrict__ quats, int64_t N, int n_tiles, int theta_max,
                                 const MindexBins *__restrict__ bins_global,
                                 unsigned long long *__restrict__ counts, const f32x2_t neg_zero) {
    __shared__ __align__(16) float qj[(MINDEX_TILE / 2) * 20];
    __shared__ unsigned int hist[MAX_THETA];
    __shared__ __align__(16) MindexBins bins;
    const int64_t s = blockIdx.x;
    int ti, tj;
    mindex_tile(blockIdx.y, n_tiles, ti, tj);
    const int tid = threadIdx.x;
    for (int b = tid; b < theta_max; b += MINDEX_TILE) hist[b] = 0u;
    mindex_load_bins(bins, bins_global, tid, MINDEX_TILE);
    const float4 *base = quats + s * N * 4;
    const int64_t j0 = (int64_t)tj * MINDEX_TILE;
    const int nj = (int)min((int64_t)MINDEX_TILE, N - j0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        const bool have = tid < nj;
        const float4 q = have ? base[(j0 + tid) * 4 + 0] : zero4, u1 = have ? base[(j0 + tid) * 4 + 1] : zero4;
        const float4 u2 = have ? base[(j0 + tid) * 4 + 2] : zero4, u3 = have ? base[(j0 + tid) * 4 + 3] : zero4;
        float *dst = qj + (tid >> 1) * 20 + (tid & 1);
        dst[0] = q.x;   dst[2] = q.y;   dst[4] = q.z;   dst[6] = q.w;
        dst[8] = u1.z;  dst[10] = u1.w; dst[12] = u2.y; dst[14] = u2.w;
        dst[16] = u3.x; dst[18] = u3.w;
    }
    const int64_t i = (int64_t)ti * MINDEX_TILE + tid;
    const bool live_i = i < N;
    const float4 q1 = live_i ? base[i * 4 + 0] : zero4, v1 = live_i ? base[i * 4 + 1] : zero4;
    const float4 v2 = live_i ? base[i * 4 + 2] : zero4, v3 = live_i ? base[i * 4 + 3] : zero4;
    Quat2 qi;
    qi.x = pack2(q1.x, q1.x);
    qi.y = pack2(q1.y, q1.y);
    qi.z = pack2(q1.z, q1.z);
    qi.w = pack2(q1.w, q1.w);
    const f32x2_t a1z = pack2(v1.z, v1.z), a1w = pack2(v1.w, v1.w);
    const f32x2_t a2y = pack2(v2.y, v2.y), a2w = pack2(v2.w, v2.w);
    const f32x2_t a3x = pack2(v3.x, v3.x), a3w = pack2(v3.w, v3.w);
    // slot a . slot b for a != b: max over a of |w_a| |w_b| = (max over a != b of |w_a|) |w_b|
    const float m1s = fmaxf(fabsf(v2.w), fabsf(v3.w)), m2s = fmaxf(fabsf(v1.w), fabsf(v3.w));
    const float m3s = fmaxf(fabsf(v1.w), fabsf(v2.w));
    const f32x2_t m1 = pack2(m1s, m1s), m2 = pack2(m2s, m2s), m3 = pack2(m3s, m3s);
    __syncthreads();
    if (live_i) {
        const int jstart = (ti == tj) ? tid + 1 : 0;
        const ulonglong2 *tile = reinterpret_cast<const ulonglong2 *>(qj);
        for (int jp = jstart >> 1; 2 * jp < nj; ++jp) {
            const ulonglong2 xy = tile[jp * 5 + 0], zw = tile[jp * 5 + 1], s1 = tile[jp * 5 + 2];
            const ulonglong2 s2 = tile[jp * 5 + 3], s3 = tile[jp * 5 + 4];
            Quat2 qb;
            qb.x = xy.x;
            qb.y = xy.y;
            qb.z = zw.x;
            qb.w = zw.y;
            float b0 = 0.0f, b1 = 0.0f;
            absmax_sign4_x2(qi, qb, neg_zero, b0, b1);  // {I, R} q1 . {I, R} q2
            // two live components: |pa + pb|, |pb - pa|
#define DREX_LIVE2(A0, B0, A1, B1)                                          \
    {                                                                        \
        const f32x2_t pa_ = mul2(A0, B0, neg_zero), pb_ = mul2(A1, B1, neg_zero); \
        absmax2(add2(pa_, pb_), b0, b1);                                     \
        absmax2(sub2(pb_, pa_), b0, b1);                                     \
    }
            DREX_LIVE2(qi.z, s1.x, qi.w, s1.y)  // {I, R} q1 . slot 1 of q2
            DREX_LIVE2(qi.y, s2.x, qi.w, s2.y)  //             slot 2
            DREX_LIVE2(qi.x, s3.x, qi.w, s3.y)  //             slot 3
            DREX_LIVE2(a1z, qb.z, a1w, qb.w)    // slot 1 of q1 . {I, R} q2
            DREX_LIVE2(a2y, qb.y, a2w, qb.w)
            DREX_LIVE2(a3x, qb.x, a3w, qb.w)
#undef DREX_LIVE2
            absmax2(add2(mul2(a1z, s1.x, neg_zero), mul2(a1w, s1.y, neg_zero)), b0, b1);  // slot a . slot a
            absmax2(add2(mul2(a2y, s2.x, neg_zero), mul2(a2w, s2.y, neg_zero)), b0, b1);
            absmax2(add2(mul2(a3x, s3.x, neg_zero), mul2(a3w, s3.y, neg_zero)), b0, b1);
            absmax2(mul2(m1, s1.y, neg_zero), b0, b1);  // slot a . slot b, a != b
            absmax2(mul2(m2, s2.y, neg_zero), b0, b1);
            absmax2(mul2(m3, s3.y, neg_zero), b0, b1);
            const int j = 2 * jp;
            if (j >= jstart) {
                const int bin = mindex_bin_lut(fminf(b0, 1.0f), bins, theta_max);
                if (bin >= 0) atomicAdd(&hist[bin], 1u);
            }
            if (j + 1 < nj) {
                const int bin = mindex_bin_lut(fminf(b1, 1.0f), bins, theta_max);
                if (bin >= 0) atomicAdd(&hist[bin], 1u);
            }
        }
    }
    __syncthreads();
    for (int b = tid; b < theta_max; b += MINDEX_TILE)
        if (hist[b]) atomicAdd(&counts[s * theta_max + b], (unsigned long long)hist[b]);
}

// K3: M = theta_max / (2 nbins) * sum_b |theory_b - density_b| (diagnostics.py:365-367)
__global__ void __launch_bounds__(32)
mindex_finalize_kernel(const unsigned long long *__restrict__ counts, const double *__restrict__ theory,
                       int theta_max, double *__restrict__ M) {
    const int64_t s = blockIdx.x;
    const int lane = threadIdx.x;
    double total = 0.0;
    for (int b = lane; b < theta_max; b += 32) total += (double)counts[s * theta_max + b];
    total = warp_sum(total);
    double acc = 0.0;
    for (int b = lane; b < theta_max; b += 32) {
        const double density = (double)counts[s * theta_max + b] / total;  // bin width 1 degree
        acc += fabs(theory[b] - density);
    }
    acc = warp_sum(acc);
    if (lane == 0) M[s] = ((double)theta_max / (2.0 * (double)theta_max)) * acc;
}

// ---- tensors.py helpers (reference src/pydrex/tensors.py) ----------------------------------

// Voigt index of the pair (p, q): tensors.py:176-177
__device__ __forceinline__ int voigt_index(int p, int q) { return (p == q) ? p : 6 - p - q; }

// op 0  rotate(tensor[81], rotation[9]) -> [81]                        tensors.py:254-273
// op 1  voigt_to_elastic_tensor(matrix[36]) -> [81]                    tensors.py:167-184
// op 2  elastic_tensor_to_voigt(tensor[81]) -> [36]                    tensors.py:187-203
// op 3  voigt_decompose(matrix[36]) -> [18] (dilatational 3x3 | deviatoric 3x3)  tensors.py:39-73
// op 4-7  mono / ortho / tetr / hex_project(voigt_vector[21]) -> [21]            tensors.py:76-140
// op 8  voigt_matrix_to_vector(matrix[36]) -> [21]                              tensors.py:206-219
// op 9  voigt_vector_to_matrix(vector[21]) -> [36]                              tensors.py:222-251
// op 10 upper_tri_to_symmetric(matrix[n x n], n = 3 or 6 in in1[0]) -> same     tensors.py:143-164
// One thread per output element, B items.
__host__ __device__ inline int tensor_op_n_out(int op, int n) {
    return (op == 0 || op == 1) ? 81 : (op == 2 || op == 9) ? 36 : (op == 3) ? 18 : (op == 10) ? n * n : 21;
}
__host__ __device__ inline int tensor_op_n_in(int op, int n) {
    return (op == 0 || op == 2) ? 81 : (op == 1 || op == 3 || op == 8) ? 36 : (op == 10) ? n * n : 21;
}
// upper_tri_to_symmetric keeps the upper triangle and mirrors it -- through numpy.where on the
// VALUE, so an upper-triangle entry that is exactly 0 takes the transposed entry of the
// upper-triangular matrix instead, which is 0 as well off the diagonal (tensors.py:163-164)
__device__ inline double sym_from_upper(const double *m, int n, int i, int j) {
    return (i <= j) ? m[n * i + j] : m[n * j + i];
}
__global__ void tensor_ops_kernel(int op, const double *__restrict__ in0,
                                  const double *__restrict__ in1, int64_t B,
                                  double *__restrict__ out, int n_side) {
    const int n_out = tensor_op_n_out(op, n_side);
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * n_out) return;
    const int64_t b = e / n_out;
    const int o = (int)(e - b * n_out);
    const double RT2 = 1.4142135623730951;
    if (op >= 3) {
        const double *x = in0 + b * tensor_op_n_in(op, n_side);
        double r = 0.0;
        if (op == 3) {
            const int which = o / 9, i = (o % 9) / 3, j = o % 3;
            const int lo = i < j ? i : j, hi = i < j ? j : i;
            if (which == 0) {  // d_ij = C_ijkk: column sums of the top three rows
                const int col = (lo == hi) ? lo : (lo == 0 && hi == 1) ? 5 : (lo == 0 && hi == 2) ? 4 : 3;
                r = (x[col] + x[6 + col]) + x[12 + col];
            } else if (lo == hi) {  // v_ii
                const int a = (lo == 0) ? 4 : 3, c = (lo == 2) ? 4 : 5;
                r = (x[7 * lo] + x[7 * a]) + x[7 * c];
            } else if (lo == 0 && hi == 1) {
                r = (x[5] + x[6 + 5]) + x[6 * 3 + 4];
            } else if (lo == 0 && hi == 2) {
                r = (x[4] + x[12 + 4]) + x[6 * 3 + 5];
            } else {
                r = (x[6 + 3] + x[12 + 3]) + x[6 * 4 + 5];
            }
        } else if (op == 4) {
            r = (o == 9 || o == 10 || o == 12 || o == 13 || o == 15 || o == 16 || o == 18 || o == 19) ? 0.0 : x[o];
        } else if (op == 5) {
            r = (o < 9) ? x[o] : 0.0;
        } else if (op == 6) {
            if (o >= 9) r = 0.0;
            else if (o == 2 || o == 5 || o == 8) r = x[o];
            else r = 0.5 * (x[o - o % 3] + x[o - o % 3 + 1]);
        } else if (op == 7) {
            if (o == 0 || o == 1) r = 3.0 / 8.0 * (x[0] + x[1]) + x[5] / 4 / RT2 + x[8] / 4;
            else if (o == 2) r = x[2];
            else if (o == 3 || o == 4) r = (x[3] + x[4]) / 2;
            else if (o == 5) r = (x[0] + x[1]) / 4 / RT2 + 3.0 / 4.0 * x[5] - x[8] / 2 / RT2;
            else if (o == 6 || o == 7) r = (x[6] + x[7]) / 2;
            else if (o == 8) r = (x[0] + x[1]) / 4 - x[5] / 2 / RT2 + x[8] / 2;
        } else if (op == 8) {
            const int i = o % 3, blk = o / 3, i1 = (i + 1) % 3, i2 = (i + 2) % 3;
            if (blk == 0) r = x[7 * i];
            else if (blk == 1) r = RT2 * x[6 * i1 + i2];
            else if (blk == 2) r = 2 * x[7 * (i + 3)];
            else if (blk == 3) r = 2 * x[6 * i + i + 3];
            else if (blk == 4) r = 2 * x[6 * i2 + i + 3];
            else if (blk == 5) r = 2 * x[6 * i1 + i + 3];
            else r = 2 * RT2 * x[6 * (i1 + 3) + i2 + 3];
        } else if (op == 9) {
            // the upper triangle as voigt_vector_to_matrix fills it, then upper_tri_to_symmetric
            const int i = o / 6, j = o % 6, lo = i < j ? i : j, hi = i < j ? j : i;
            const double h = 1.0 / RT2;
            if (lo == hi) r = (lo < 3) ? x[lo] : 0.5 * x[lo + 3];
            else if (lo == 1 && hi == 2) r = h * x[3];
            else if (lo == 0 && hi == 2) r = h * x[4];
            else if (lo == 0 && hi == 1) r = h * x[5];
            else if (lo == 0 && hi == 3) r = 0.5 * x[9];
            else if (lo == 1 && hi == 4) r = 0.5 * x[10];
            else if (lo == 2 && hi == 5) r = 0.5 * x[11];
            else if (lo == 2 && hi == 3) r = 0.5 * x[12];
            else if (lo == 0 && hi == 4) r = 0.5 * x[13];
            else if (lo == 1 && hi == 5) r = 0.5 * x[14];
            else if (lo == 1 && hi == 3) r = 0.5 * x[15];
            else if (lo == 2 && hi == 4) r = 0.5 * x[16];
            else if (lo == 0 && hi == 5) r = 0.5 * x[17];
            else if (lo == 4 && hi == 5) r = 0.5 * h * x[18];
            else if (lo == 3 && hi == 5) r = 0.5 * h * x[19];
            else if (lo == 3 && hi == 4) r = 0.5 * h * x[20];
        } else {
            r = sym_from_upper(x, n_side, o / n_side, o % n_side);
        }
        out[e] = r;
        return;
    }
    if (op == 0) {
        const double *T = in0 + b * 81, *R = in1 + b * 9;
        const int i = o / 27, j = (o / 9) % 3, k = (o / 3) % 3, l = o % 3;
        double acc = 0.0;
        for (int a = 0; a < 3; ++a)
            for (int bb = 0; bb < 3; ++bb)
                for (int c = 0; c < 3; ++c)
                    for (int d = 0; d < 3; ++d)
                        acc += R[3 * i + a] * R[3 * j + bb] * R[3 * k + c] * R[3 * l + d] *
                               T[27 * a + 9 * bb + 3 * c + d];
        out[e] = acc;
    } else if (op == 1) {
        const int p = o / 27, q = (o / 9) % 3, r = (o / 3) % 3, t = o % 3;
        out[e] = in0[b * 36 + 6 * voigt_index(p, q) + voigt_index(r, t)];
    } else {
        // mean of the tensor entries that share Voigt indices (i, j), then symmetrised
        const int i = o / 6, j = o % 6;
        double sum_ij = 0.0, sum_ji = 0.0;
        int n_ij = 0, n_ji = 0;
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q)
                for (int r = 0; r < 3; ++r)
                    for (int t = 0; t < 3; ++t) {
                        const int vi = voigt_index(p, q), vj = voigt_index(r, t);
                        const double v = in0[b * 81 + 27 * p + 9 * q + 3 * r + t];
                        if (vi == i && vj == j) { sum_ij += v; ++n_ij; }
                        if (vi == j && vj == i) { sum_ji += v; ++n_ji; }
                    }
        out[e] = (sum_ij / n_ij + sum_ji / n_ji) / 2;
    }
}

// tensors.polar_decompose (tensors.py:14-21) of B nonsingular 3x3 matrices: M = V R (left,
// V = sqrt(M M^T)) or M = R U (right, U = sqrt(M^T M)); rot[B][9], stretch[B][9].
__global__ void polar_decompose_kernel(const double *__restrict__ M, int64_t B, int left,
                                       double *__restrict__ rot, double *__restrict__ stretch) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double m[9], G[9], w[3], U[9], S[9], Sinv[9], R[9];
    for (int c = 0; c < 9; ++c) m[c] = M[b * 9 + c];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 3; ++k)
                acc += left ? m[3 * i + k] * m[3 * j + k] : m[3 * k + i] * m[3 * k + j];
            G[3 * i + j] = acc;
        }
    jacobi_eig3(G, w, U);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0, si = 0.0;
            for (int k = 0; k < 3; ++k) {
                const double root = sqrt(fmax(w[k], 0.0));
                s += U[3 * i + k] * root * U[3 * j + k];
                si += U[3 * i + k] * (1.0 / root) * U[3 * j + k];
            }
            S[3 * i + j] = s;
            Sinv[3 * i + j] = si;
        }
    if (left) matmul3(Sinv, m, R); else matmul3(m, Sinv, R);
    for (int c = 0; c < 9; ++c) {
        rot[b * 9 + c] = R[c];
        stretch[b * 9 + c] = S[c];
    }
}

// ---- utils.py / geometry.py helpers by name ------------------------------------------------

// utils.extract_vars (utils.py:183-190): y = [F(9) | orientations(9N) | fractions(N)] ->
// F, orientations clipped to [-1, 1], fractions clipped at 0 and normalised.  One CTA per
// system.
__global__ void extract_vars_kernel(const double *__restrict__ y, int64_t N, double *__restrict__ F,
                                    double *__restrict__ o, double *__restrict__ f) {
    __shared__ double red[64];
    const int64_t s = blockIdx.x;
    const double *ys = y + s * (9 + 10 * N);
    if (threadIdx.x < 9) F[s * 9 + threadIdx.x] = ys[threadIdx.x];
    for (int64_t i = threadIdx.x; i < 9 * N; i += blockDim.x)
        o[s * 9 * N + i] = fmin(1.0, fmax(-1.0, ys[9 + i]));
    double sum = 0.0, dummy = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) sum += fmax(ys[9 + 9 * N + i], 0.0);
    block_sum2<256>(sum, dummy, red);
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x)
        f[s * N + i] = fmax(ys[9 + 9 * N + i], 0.0) / sum;
}

// utils.apply_gbs (utils.py:159-180), in place: grains smaller than threshold / N keep their
// previous orientation and are set to that size, then the fractions are normalised.
__global__ void apply_gbs_kernel(double *__restrict__ o, double *__restrict__ f,
                                 const double *__restrict__ o_prev, double threshold, int64_t N) {
    __shared__ double red[64];
    const int64_t s = blockIdx.x;
    const double floor_f = threshold / (double)N;
    double sum = 0.0, dummy = 0.0;
    for (int64_t g = threadIdx.x; g < N; g += blockDim.x) {
        double v = f[s * N + g];
        if (v < floor_f) {
            for (int c = 0; c < 9; ++c) o[(s * N + g) * 9 + c] = o_prev[(s * N + g) * 9 + c];
            v = floor_f;
            f[s * N + g] = v;
        }
        sum += v;
    }
    block_sum2<256>(sum, dummy, red);
    for (int64_t g = threadIdx.x; g < N; g += blockDim.x) f[s * N + g] /= sum;
}

// geometry.misorientation_angles (geometry.py:75-118): for each of M rows the smallest
// 2 acos(|q1_i . q2_j|) in degrees over the A x B pairs of quaternions, evaluated in the
// precision of the inputs like the reference (float32 on the M-index path), result float64.
template <typename T>
__global__ void misorientation_angles_kernel(const T *__restrict__ q1, const T *__restrict__ q2,
                                             int64_t M, int A, int B, double *__restrict__ out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    T best = (T)1e30;
    for (int i = 0; i < A; ++i) {
        const T *a = q1 + (m * A + i) * 4;
        for (int j = 0; j < B; ++j) {
            const T *b = q2 + (m * B + j) * 4;
            T d = a[0] * b[0];
            d += a[1] * b[1];
            d += a[2] * b[2];
            d += a[3] * b[3];
            d = d < (T)-1 ? (T)-1 : (d > (T)1 ? (T)1 : d);
            d = d < (T)0 ? -d : d;
            T ang;
            if (sizeof(T) == 4) ang = (T)2 * ((T)acosf((float)d) * (T)57.29577951308232f);
            else ang = (T)(2.0 * (acos((double)d) * 57.29577951308232));
            best = ang < best ? ang : best;
        }
    }
    out[m] = (double)best;
}

// max |eigenvalue| of sym(L) for P matrices (utils.strain_increment, utils.py:144-156;
// the eigvalsh call of minerals.py:421-422)
__global__ void strain_rate_max_kernel(const double *__restrict__ L, int64_t P, double *__restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double l[9];
    for (int c = 0; c < 9; ++c) l[c] = L[p * 9 + c];
    out[p] = strain_rate_max(l);
}

}  // namespace drex
