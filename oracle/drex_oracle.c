/*
 * drex_oracle.c -- CPU restatement of the PyDRex CPO hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for pydrex_b200: a plain, scalar, single-threaded C
 * restatement of the arithmetic the reference performs on its hot path.  It is NOT part
 * of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  The product path (pydrex_b200/) never does.
 *
 * Parity status: PINNED.  Every function below is checked in tests/test_oracle.py against
 * golden vectors generated from the unmodified reference (tests/golden/make_golden.py,
 * which imports /root/reference/src/pydrex in the build container).
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference/src/pydrex).  The code is written from the algorithm, not copied:
 * tensor sums are expressed with small vector helpers, the permutation-symbol sum is a
 * cross product, sorting is an explicit stable insertion sort (what numba's argsort does
 * for 4 elements), and the 3x3 eigen/polar problems use cyclic Jacobi instead of LAPACK.
 *
 * All matrices are row-major double[9]; orientations are double[N][9] (AoS, as the
 * reference stores them), rows = crystal a, b, c axes in the global frame.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ small helpers */

static inline double dot3(const double *u, const double *v) {
    return u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
}
static inline void matvec3(const double *M, const double *v, double *out) {
    for (int i = 0; i < 3; ++i) out[i] = dot3(M + 3 * i, v);
}
static inline void cross3(const double *u, const double *v, double *out) {
    out[0] = u[1] * v[2] - u[2] * v[1];
    out[1] = u[2] * v[0] - u[0] * v[2];
    out[2] = u[0] * v[1] - u[1] * v[0];
}

/* Cyclic Jacobi for a symmetric 3x3: A = V diag(w) V^T, V columns = eigenvectors. */
static void jacobi_eig3(const double *A_in, double *w, double *V) {
    double A[9];
    memcpy(A, A_in, sizeof(A));
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
        double diag = fabs(A[0]) + fabs(A[4]) + fabs(A[8]);
        if (off <= 1e-300 || off <= 1e-22 * diag) break;
        for (int p = 0; p < 2; ++p) {
            for (int q = p + 1; q < 3; ++q) {
                double apq = A[3 * p + q];
                if (apq == 0.0) continue;
                double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { /* A <- A J */
                    double akp = A[3 * k + p], akq = A[3 * k + q];
                    A[3 * k + p] = c * akp - s * akq;
                    A[3 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) { /* A <- J^T A */
                    double apk = A[3 * p + k], aqk = A[3 * q + k];
                    A[3 * p + k] = c * apk - s * aqk;
                    A[3 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = V[3 * k + p], vkq = V[3 * k + q];
                    V[3 * k + p] = c * vkp - s * vkq;
                    V[3 * k + q] = s * vkp + c * vkq;
                }
            }
        }
    }
    w[0] = A[0];
    w[1] = A[4];
    w[2] = A[8];
}

/* max |eigenvalue| of the symmetric part of L: minerals.py:421-422 */
ORACLE_API double oracle_strain_rate_max(const double *L) {
    double D[9], w[3], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) D[3 * i + j] = 0.5 * (L[3 * i + j] + L[3 * j + i]);
    jacobi_eig3(D, w, V);
    return fmax(fabs(w[0]), fmax(fabs(w[1]), fabs(w[2])));
}

/* Left stretch V of M = V R (tensors.py:14-19 with left=True, element [1]):
 * U diag(S) U^T = sqrt(M M^T). */
ORACLE_API void oracle_left_stretch(const double *M, double *Vout) {
    double G[9], w[3], U[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) G[3 * i + j] = dot3(M + 3 * i, M + 3 * j);
    jacobi_eig3(G, w, U);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += U[3 * i + k] * sqrt(fmax(w[k], 0.0)) * U[3 * j + k];
            Vout[3 * i + j] = s;
        }
}

/* ------------------------------------------------------------------ core.py */

/* core.py:315-342.  Returns 0, or -1 for an unsupported (phase, fabric). */
ORACLE_API int oracle_get_crss(int phase, int fabric, double *crss) {
    static const double table[6][4] = {
        {1, 2, 3, INFINITY}, {3, 2, 1, INFINITY}, {3, 2, INFINITY, 1},
        {1, 1, 3, INFINITY}, {3, 1, 2, INFINITY}, {INFINITY, INFINITY, INFINITY, 1}};
    if (phase == 0 && fabric >= 0 && fabric <= 4) {
        memcpy(crss, table[fabric], 4 * sizeof(double));
        return 0;
    }
    if (phase == 1 && fabric == 5) {
        memcpy(crss, table[5], 4 * sizeof(double));
        return 0;
    }
    return -1;
}

/* Slip systems in the reference's order, as (direction row, normal row) of the
 * orientation matrix: (010)[100], (001)[100], (010)[001], (100)[001]
 * (core.py:575-598, minerals.py:37-42). */
static const int SLIP_DIR[4] = {0, 0, 2, 2};
static const int SLIP_NRM[4] = {1, 2, 1, 0};

/* core.py:571-598: I_s = l_s . (D n_s) */
ORACLE_API void oracle_slip_invariants(const double *D, const double *a, double *inv) {
    for (int s = 0; s < 4; ++s) {
        double Dn[3];
        matvec3(D, a + 3 * SLIP_NRM[s], Dn);
        inv[s] = dot3(a + 3 * SLIP_DIR[s], Dn);
    }
}

/* Stable insertion argsort of 4 keys (numba's np.argsort on <16 elements is an
 * insertion sort; core.py:724,732). */
static void argsort4(const double *key, int *idx) {
    for (int i = 0; i < 4; ++i) idx[i] = i;
    for (int i = 1; i < 4; ++i) {
        int cur = idx[i], j = i - 1;
        while (j >= 0 && key[idx[j]] > key[cur]) {
            idx[j + 1] = idx[j];
            --j;
        }
        idx[j + 1] = cur;
    }
}

/* core.py:541-568 */
ORACLE_API void oracle_slip_rates_olivine(const double *inv, const int *idx, const double *crss,
                                          double n, double *rates) {
    int i_inac = idx[0], i_min = idx[1], i_int = idx[2], i_max = idx[3];
    double pre = crss[i_max] / inv[i_max];
    double r_min = pre * inv[i_min] / crss[i_min];
    double r_int = pre * inv[i_int] / crss[i_int];
    rates[i_inac] = 0.0;
    rates[i_min] = r_min * pow(fabs(r_min), n - 1.0);
    rates[i_int] = r_int * pow(fabs(r_int), n - 1.0);
    rates[i_max] = 1.0;
}

/* core.py:477-501: G = 2 sum_s rate_s l_s (x) n_s */
ORACLE_API void oracle_deformation_rate(const double *a, const double *rates, double *G) {
    for (int i = 0; i < 9; ++i) G[i] = 0.0;
    for (int s = 0; s < 4; ++s) {
        const double *l = a + 3 * SLIP_DIR[s], *nrm = a + 3 * SLIP_NRM[s];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) G[3 * i + j] += 2.0 * rates[s] * l[i] * nrm[j];
    }
}

/* core.py:504-538 (corrected index k = j+1) */
ORACLE_API double oracle_slip_rate_softest(const double *G, const double *L) {
    double num = 0.0, den = 0.0;
    for (int j = 0; j < 3; ++j) {
        int k = (j + 1) % 3;
        double gasym = G[3 * j + k] - G[3 * k + j];
        num -= (L[3 * j + k] - L[3 * k + j]) * gasym;
        den -= gasym * gasym;
        for (int l = 0; l < 3; ++l) {
            num += 2.0 * G[3 * j + l] * L[3 * j + l];
            den += 2.0 * G[3 * j + l] * G[3 * j + l];
        }
    }
    if (den > -1e-15 && den < 1e-15) return 0.0;
    return num / den;
}

/* core.py:601-642: spin vector, then da_p = a_p x omega for each row p
 * (sum_{rs} eps_{qrs} a_{ps} w_r = (w x a_p)_q ... with the reference's index order
 * this equals -(a_p x w)_q; written out below exactly as the index sum gives). */
ORACLE_API void oracle_orientation_change(const double *a, const double *L, const double *G,
                                          double gamma0, double *da) {
    double w[3];
    for (int j = 0; j < 3; ++j) {
        int r = (j + 1) % 3, s = (j + 2) % 3;
        w[j] = 0.5 * ((L[3 * s + r] - L[3 * r + s]) - (G[3 * s + r] - G[3 * r + s]) * gamma0);
    }
    for (int p = 0; p < 3; ++p) {
        /* da[p][q] = sum_{r,s} eps[q][r][s] a[p][s] w[r] = (w x a_p)[q] */
        cross3(w, a + 3 * p, da + 3 * p);
    }
}

/* core.py:645-684: loops over slip systems 0..2 in natural order (never index 3). */
ORACLE_API double oracle_strain_energy(const double *crss, const double *rates, double gamma0,
                                       double p, double n, double lambda) {
    double E = 0.0;
    for (int i = 0; i < 3; ++i) {
        double rho = pow(1.0 / crss[i], n - p) * pow(fabs(rates[i] * gamma0), p / n);
        E += rho * exp(-lambda * rho * rho);
    }
    return E;
}

/* core.py:687-760.  Also exports the intermediates when the pointers are non-NULL. */
ORACLE_API int oracle_rotation_and_strain(int phase, int fabric, const double *a, const double *D,
                                          const double *L, double p, double n, double lambda,
                                          double *da, double *energy, double *inv_out,
                                          double *rates_out, double *G_out, double *gamma0_out) {
    double crss[4], inv[4], rates[4] = {0, 0, 0, 0}, G[9], key[4];
    int idx[4];
    if (oracle_get_crss(phase, fabric, crss) != 0) return -1;
    oracle_slip_invariants(D, a, inv);
    if (inv_out) memcpy(inv_out, inv, sizeof(inv));
    if (inv[0] == 0 && inv[1] == 0 && inv[2] == 0 && inv[3] == 0) {
        for (int i = 0; i < 9; ++i) da[i] = 0.0;
        *energy = 0.0;
        if (rates_out) memset(rates_out, 0, 4 * sizeof(double));
        if (G_out) memset(G_out, 0, 9 * sizeof(double));
        if (gamma0_out) *gamma0_out = 0.0;
        return 0;
    }
    if (phase == 0) {
        for (int s = 0; s < 4; ++s) key[s] = fabs(inv[s] / crss[s]);
        argsort4(key, idx);
        oracle_slip_rates_olivine(inv, idx, crss, n, rates);
    } else {
        if (fabs(inv[3]) > 1e-15) rates[3] = 1.0;
    }
    oracle_deformation_rate(a, rates, G);
    double gamma0 = oracle_slip_rate_softest(G, L);
    oracle_orientation_change(a, L, G, gamma0, da);
    *energy = oracle_strain_energy(crss, rates, gamma0, p, n, lambda);
    if (rates_out) memcpy(rates_out, rates, sizeof(rates));
    if (G_out) memcpy(G_out, G, sizeof(G));
    if (gamma0_out) *gamma0_out = gamma0;
    return 0;
}

/* core._get_rotation_and_strain (core.py:687-760) for G grains, each with ITS OWN flow:
 * Ls[g] = L / strain_rate_max of grain g (D is formed here as its symmetric part,
 * minerals.py:421,438-439).  Used by oracle/device_model.py, where chunks of grains of one
 * system sit at different times of the interval. */
ORACLE_API int oracle_grain_batch(int phase, int fabric, int64_t G, const double *o,
                                  const double *Ls, double p, double n, double lambda,
                                  double *da, double *energy) {
    for (int64_t g = 0; g < G; ++g) {
        const double *L = Ls + 9 * g;
        double D[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) D[3 * i + j] = 0.5 * (L[3 * i + j] + L[3 * j + i]);
        int rc = oracle_rotation_and_strain(phase, fabric, o + 9 * g, D, L, p, n, lambda,
                                            da + 9 * g, energy + g, NULL, NULL, NULL, NULL);
        if (rc != 0) return -1;
    }
    return 0;
}

/* core.py:347-474.  Returns 0; -1 bad phase/fabric; -2 unsupported regime; -3 bad regime. */
ORACLE_API int oracle_derivatives(int regime, int phase, int fabric, int64_t n_grains,
                                  const double *o, const double *f, const double *D,
                                  const double *L, const double *spin, double p, double n,
                                  double lambda, double M, double vol_frac, double *od,
                                  double *fd) {
    switch (regime) {
    case 0:
    case 7: /* identity matrices, zero volume change: core.py:386-391,467-472 */
        for (int64_t g = 0; g < n_grains; ++g) {
            for (int i = 0; i < 9; ++i) od[9 * g + i] = (i % 4 == 0) ? 1.0 : 0.0;
            fd[g] = 0.0;
        }
        return 0;
    case 1: /* every grain gets the `deformation_gradient_spin` matrix: core.py:392-405 */
        for (int64_t g = 0; g < n_grains; ++g) {
            memcpy(od + 9 * g, spin, 9 * sizeof(double));
            fd[g] = 0.0;
        }
        return 0;
    case 2:
    case 3:
    case 5:
        return -2;
    case 4:
    case 6: {
        double scale = (regime == 6) ? 0.3 : 1.0; /* core.py:459,464 */
        double *E = (double *)malloc(sizeof(double) * (size_t)(n_grains > 0 ? n_grains : 1));
        double mean = 0.0;
        for (int64_t g = 0; g < n_grains; ++g) {
            int rc = oracle_rotation_and_strain(phase, fabric, o + 9 * g, D, L, p, n, lambda,
                                                od + 9 * g, E + g, NULL, NULL, NULL, NULL);
            if (rc != 0) {
                free(E);
                return -1;
            }
            if (regime == 6)
                for (int i = 0; i < 9; ++i) od[9 * g + i] *= 0.3;
            mean += f[g] * E[g]; /* core.py:432 */
        }
        for (int64_t g = 0; g < n_grains; ++g)
            fd[g] = vol_frac * M * f[g] * (scale * (mean - E[g])); /* core.py:434-435,464-465 */
        free(E);
        return 0;
    }
    default:
        return -3;
    }
}

/* ------------------------------------------------------------------ utils.py / minerals.py */

/* utils.py:183-190 on a state vector y = [F(9) | o(9N) | f(N)]; writes clipped copies. */
ORACLE_API void oracle_extract_vars(const double *y, int64_t N, double *F, double *o, double *f) {
    memcpy(F, y, 9 * sizeof(double));
    for (int64_t i = 0; i < 9 * N; ++i) o[i] = fmin(1.0, fmax(-1.0, y[9 + i]));
    double sum = 0.0;
    for (int64_t g = 0; g < N; ++g) {
        f[g] = fmax(y[9 + 9 * N + g], 0.0);
        sum += f[g];
    }
    for (int64_t g = 0; g < N; ++g) f[g] /= sum;
}

/* utils.py:159-180, in place */
ORACLE_API void oracle_apply_gbs(double *o, double *f, double threshold, const double *o_prev,
                                 int64_t N) {
    double floor_v = threshold / (double)N, sum = 0.0;
    for (int64_t g = 0; g < N; ++g) {
        if (f[g] < floor_v) {
            memcpy(o + 9 * g, o_prev + 9 * g, 9 * sizeof(double));
            f[g] = floor_v;
        }
        sum += f[g];
    }
    for (int64_t g = 0; g < N; ++g) f[g] /= sum;
}

/* The ODE right-hand side, minerals.py:388-453, for a given L = L(t, x(t)).
 * work must hold 10N doubles.  Returns the oracle_derivatives status. */
ORACLE_API int oracle_eval_rhs(const double *L, const double *y, int64_t N, int regime, int phase,
                               int fabric, double p, double n, double lambda, double M,
                               double vol_frac, double *dy, double *work) {
    double F[9], D[9], Ls[9], dF[9], spin[9];
    double *o = work, *f = work + 9 * N;
    double smax = oracle_strain_rate_max(L);
    oracle_extract_vars(y, N, F, o, f);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            D[3 * i + j] = 0.5 * (L[3 * i + j] + L[3 * j + i]) / smax;
            Ls[3 * i + j] = L[3 * i + j] / smax;
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += L[3 * i + k] * F[3 * k + j];
            dF[3 * i + j] = s;
        }
    oracle_left_stretch(dF, spin); /* minerals.py:427-429 */
    int rc = oracle_derivatives(regime, phase, fabric, N, o, f, D, Ls, spin, p, n, lambda, M,
                                vol_frac, dy + 9, dy + 9 + 9 * N);
    if (rc != 0) return rc;
    memcpy(dy, dF, sizeof(dF));
    for (int64_t i = 9; i < 9 + 10 * N; ++i) dy[i] *= smax; /* minerals.py:450-451 */
    return 0;
}

/* ------------------------------------------------------------------ tensors.py / voigt */

/* Voigt index of the pair (p, q): tensors.py:174-175 */
static inline int voigt_index(int p, int q) { return p == q ? p : 6 - p - q; }

/* tensors.py:167-184 */
ORACLE_API void oracle_voigt_to_tensor(const double *S, double *T) {
    for (int p = 0; p < 3; ++p)
        for (int q = 0; q < 3; ++q)
            for (int r = 0; r < 3; ++r)
                for (int s = 0; s < 3; ++s)
                    T[((p * 3 + q) * 3 + r) * 3 + s] = S[6 * voigt_index(p, q) + voigt_index(r, s)];
}

/* tensors.py:187-203 */
ORACLE_API void oracle_tensor_to_voigt(const double *T, double *S) {
    double acc[36] = {0}, cnt[36] = {0};
    for (int p = 0; p < 3; ++p)
        for (int q = 0; q < 3; ++q)
            for (int r = 0; r < 3; ++r)
                for (int s = 0; s < 3; ++s) {
                    int ij = 6 * voigt_index(p, q) + voigt_index(r, s);
                    acc[ij] += T[((p * 3 + q) * 3 + r) * 3 + s];
                    cnt[ij] += 1.0;
                }
    for (int i = 0; i < 36; ++i) acc[i] /= cnt[i];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) S[6 * i + j] = 0.5 * (acc[6 * i + j] + acc[6 * j + i]);
}

/* tensors.py:254-273, done as four successive single-index contractions. */
ORACLE_API void oracle_rotate_tensor(const double *T, const double *R, double *out) {
    double A[81], B[81];
    memcpy(A, T, sizeof(A));
    for (int axis = 0; axis < 4; ++axis) {
        /* contract index `axis` of A with R[., .]: B[..i..] = sum_a R[i][a] A[..a..] */
        int stride = 1;
        for (int k = axis + 1; k < 4; ++k) stride *= 3;
        for (int idx = 0; idx < 81; ++idx) {
            int i = (idx / stride) % 3;
            int base = idx - i * stride;
            double s = 0.0;
            for (int a = 0; a < 3; ++a) s += R[3 * i + a] * A[base + a * stride];
            B[idx] = s;
        }
        memcpy(A, B, sizeof(A));
    }
    memcpy(out, A, sizeof(A));
}

/* minerals.py:729-739 for one snapshot of one phase: Cbar += sum_g voigt(rotate(C, a_g^T))
 * * f_g * phase_fraction.  S is the 6x6 stiffness of the phase; Cbar (36) is accumulated. */
ORACLE_API void oracle_voigt_average_accumulate(const double *o, const double *f, int64_t N,
                                                const double *S, double phase_fraction,
                                                double *Cbar) {
    double T[81], Trot[81], V[36], Rt[9];
    oracle_voigt_to_tensor(S, T);
    for (int64_t g = 0; g < N; ++g) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Rt[3 * i + j] = o[9 * g + 3 * j + i];
        oracle_rotate_tensor(T, Rt, Trot);
        oracle_tensor_to_voigt(Trot, V);
        for (int i = 0; i < 36; ++i) Cbar[i] += V[i] * f[g] * phase_fraction;
    }
}

/* ------------------------------------------------------------------ M-index */

/* Nearest rotation (polar factor U V^T of the SVD) by the Newton iteration
 * X <- (X + X^-T)/2, which is what scipy's Rotation.from_matrix applies to matrices whose
 * Gramian is not within 1e-12 of identity (scipy/spatial/transform/_rotation_xp.py:68-84,
 * scipy 1.18.1; third-party, not under /root/reference). */
static void nearest_rotation(const double *M, double *Rout) {
    double X[9];
    memcpy(X, M, sizeof(X));
    for (int it = 0; it < 40; ++it) {
        double c0[3], c1[3], c2[3], inv_t[9];
        /* cofactor matrix rows = cross products of rows; X^-T = cof(X)/det(X) */
        cross3(X + 3, X + 6, c0);
        cross3(X + 6, X + 0, c1);
        cross3(X + 0, X + 3, c2);
        double det = dot3(X, c0);
        for (int j = 0; j < 3; ++j) {
            inv_t[j] = c0[j] / det;
            inv_t[3 + j] = c1[j] / det;
            inv_t[6 + j] = c2[j] / det;
        }
        double delta = 0.0;
        for (int i = 0; i < 9; ++i) {
            double nx = 0.5 * (X[i] + inv_t[i]);
            delta = fmax(delta, fabs(nx - X[i]));
            X[i] = nx;
        }
        if (delta < 1e-16) break;
    }
    memcpy(Rout, X, sizeof(X));
}

/* scipy Rotation.from_matrix(...).as_quat(), scalar-last (x, y, z, w):
 * _rotation_xp.py:89-157 (largest-of-diagonal-and-trace branch, then normalise). */
ORACLE_API void oracle_matrix_to_quat(const double *M_in, double *q) {
    double M[9];
    nearest_rotation(M_in, M);
    double tr = M[0] + M[4] + M[8];
    double dec[4] = {M[0], M[4], M[8], tr};
    int choice = 0;
    for (int i = 1; i < 4; ++i)
        if (dec[i] > dec[choice]) choice = i;
    if (choice == 3) {
        q[0] = M[7] - M[5];
        q[1] = M[2] - M[6];
        q[2] = M[3] - M[1];
        q[3] = 1.0 + tr;
    } else {
        int i = choice, j = (i + 1) % 3, k = (j + 1) % 3;
        q[i] = 1.0 - tr + 2.0 * M[4 * i];
        q[j] = M[3 * j + i] + M[3 * i + j];
        q[k] = M[3 * k + i] + M[3 * i + k];
        q[3] = M[3 * k + j] - M[3 * j + k];
    }
    double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= nrm;
}

/* The symmetry operations as applied by stats.py:118-123: each op is a 4x4 matrix acting
 * on a scalar-last quaternion.  Proper rotations go through utils.quat_product
 * (utils.py:365-371), which has NO cross term (np.cross(q1[:3], q1[:3]) == 0), so the
 * matrix of q -> quat_product(qs, q) is  [[w_s I3, v_s], [-v_s^T, w_s]].
 * Reflections are the diagonal sign matrices of geometry.py:147-149.
 * `system_a` is the first member of LatticeSystem.value (geometry.py:29-34).
 * Returns the number of ops written (each 16 doubles, row-major). */
ORACLE_API int oracle_symmetry_ops(int system_a, int system_b, double *ops) {
    int n = 0;
#define PUSH_QUAT(x, y, z, w)                                                                  \
    do {                                                                                       \
        double v_[3] = {x, y, z}, w_ = (w);                                                    \
        double *m_ = ops + 16 * n++;                                                           \
        memset(m_, 0, 16 * sizeof(double));                                                    \
        for (int i_ = 0; i_ < 3; ++i_) {                                                       \
            m_[4 * i_ + i_] = w_;                                                              \
            m_[4 * i_ + 3] = v_[i_];                                                           \
            m_[12 + i_] = -v_[i_];                                                             \
        }                                                                                      \
        m_[15] = w_;                                                                           \
    } while (0)
#define PUSH_DIAG(a, b, c, d)                                                                  \
    do {                                                                                       \
        double *m_ = ops + 16 * n++;                                                           \
        memset(m_, 0, 16 * sizeof(double));                                                    \
        m_[0] = a;                                                                             \
        m_[5] = b;                                                                             \
        m_[10] = c;                                                                            \
        m_[15] = d;                                                                            \
    } while (0)
    /* axes in the reference's order: z, y, x (geometry.py:137,155,163,171,178) */
    static const double AX[3][3] = {{0, 0, 1}, {0, 1, 0}, {1, 0, 0}};
    PUSH_QUAT(0, 0, 0, 1);
    if (system_a == 1) { /* triclinic */
    } else if (system_a == 2) { /* monoclinic, orthorhombic: geometry.py:133-150 */
        for (int k = 0; k < 3; ++k) {
            double h = 0.5 * M_PI, s = sin(h), c = cos(h);
            PUSH_QUAT(AX[k][0] * s, AX[k][1] * s, AX[k][2] * s, c);
        }
        PUSH_DIAG(1, -1, -1, 1);
        PUSH_DIAG(1, -1, 1, -1);
        PUSH_DIAG(1, 1, -1, -1);
    } else if (system_a == 3 || system_a == 4) { /* rhombohedral / tetragonal */
        int div = (system_a == 3) ? 3 : 2, imax = (system_a == 3) ? 2 : 3;
        for (int k = 0; k < 3; ++k)
            for (int i = 1; i <= imax; ++i) {
                double h = 0.5 * i * M_PI / div, s = sin(h), c = cos(h);
                PUSH_QUAT(AX[k][0] * s, AX[k][1] * s, AX[k][2] * s, c);
            }
    } else if (system_a == 6) { /* hexagonal: geometry.py:166-181 */
        for (int k = 0; k < 3; ++k)
            for (int i = 1; i <= 2; ++i) {
                double h = 0.5 * i * M_PI / 3, s = sin(h), c = cos(h);
                PUSH_QUAT(AX[k][0] * s, AX[k][1] * s, AX[k][2] * s, c);
            }
        for (int k = 0; k < 3; ++k)
            for (int i = 1; i <= 5; i += 2) {
                double h = 0.5 * i * M_PI / 6, s = sin(h), c = cos(h);
                PUSH_QUAT(AX[k][0] * s, AX[k][1] * s, AX[k][2] * s, c);
            }
    } else {
        return -1;
    }
    (void)system_b;
    return n;
#undef PUSH_QUAT
#undef PUSH_DIAG
}

/* stats.py:203-215 */
ORACLE_API int oracle_max_misorientation(int system_a, int system_b) {
    if ((system_a == 2 && system_b == 4) || system_a == 3) return 120;
    if (system_a == 4 || system_a == 6) return 90;
    if (system_a == 1 || (system_a == 2 && system_b == 2)) return 180;
    return -1;
}

/* stats.py:130-200: expected density of misorientation angles in (low, high) degrees for
 * a random texture (Grimmer 1979). */
ORACLE_API double oracle_misorientations_random(double low, double high, int Ms, int Ns) {
    const double deg = M_PI / 180.0;
    double a = tan(deg * 90.0 / Ms);
    double b = 2.0 * atan(sqrt(1.0 + a * a)) / deg;
    double c = round(2.0 * atan(sqrt(1.0 + 2.0 * a * a)) / deg);
    double counts[2] = {0.0, 0.0};
    double edges[2] = {low, high};
    for (int i = 0; i < 2; ++i) {
        double e = edges[i], d = e * deg;
        if (0 <= e && e <= 180.0 / Ms) {
            counts[i] += (Ns / 180.0) * (1 - cos(d));
        } else if (180.0 / Ms <= e && e <= 180.0 * Ms / Ns) {
            counts[i] += (Ns / 180.0) * a * sin(d);
        } else if (90 <= e && e <= b) {
            counts[i] += (Ms / 90.0) * ((Ms + a) * sin(d) - Ms * (1 - cos(d)));
        } else if (b <= e && e <= c) {
            double nu = pow(tan(deg * e / 2.0), 2.0);
            counts[i] = (Ms / 90.0) *
                        ((Ms + a) * sin(d) - Ms * (1 - cos(d)) +
                         (Ms / 180.0) *
                             ((1 - cos(d)) * (acos((1 - nu * cos(deg * 180.0 / Ms)) / (nu - 1)) / deg +
                                              2.0 * acos(a / (sqrt(nu - a * a) * sqrt(nu - 1))) / deg) -
                              2.0 * sin(d) *
                                  (2.0 * acos(a / sqrt(nu - 1)) / deg +
                                   a * acos(1.0 / sqrt(nu - a * a)) / deg)));
        } else {
            return NAN;
        }
    }
    return (counts[0] + counts[1]) / 2.0;
}

/* stats.py:83-127 + geometry.py:74-118 + diagnostics.py:332-367.
 * o: N orientation matrices.  hist_out (theta_max doubles, density-normalised like
 * np.histogram(..., bins=theta_max, range=(0, theta_max), density=True)) may be NULL.
 * counts_out (theta_max int64) may be NULL.  Returns the M-index.
 * Float32 storage of the transformed quaternions and the float32 dot/acos/rad2deg chain
 * follow stats.py:105-112 and geometry.py:104-116 (arrays are float32 there). */
ORACLE_API double oracle_misorientation_index(const double *o, int64_t N, int system_a,
                                              int system_b, double *hist_out,
                                              int64_t *counts_out) {
    double ops[16 * 16];
    int n_ops = oracle_symmetry_ops(system_a, system_b, ops);
    int tmax = oracle_max_misorientation(system_a, system_b);
    if (n_ops < 0 || tmax < 0) return NAN;
    float *qs = (float *)malloc(sizeof(float) * (size_t)N * n_ops * 4);
    for (int64_t g = 0; g < N; ++g) {
        double q[4];
        oracle_matrix_to_quat(o + 9 * g, q);
        for (int k = 0; k < n_ops; ++k)
            for (int i = 0; i < 4; ++i) {
                double s = 0.0;
                for (int j = 0; j < 4; ++j) s += ops[16 * k + 4 * i + j] * q[j];
                qs[(g * n_ops + k) * 4 + i] = (float)s;
            }
    }
    int64_t *counts = (int64_t *)calloc((size_t)tmax, sizeof(int64_t));
    const float rad2deg_f = (float)(180.0 / M_PI);
    for (int64_t g1 = 0; g1 < N; ++g1)
        for (int64_t g2 = g1 + 1; g2 < N; ++g2) {
            double best = INFINITY;
            for (int i = 0; i < n_ops; ++i)
                for (int j = 0; j < n_ops; ++j) {
                    const float *u = qs + (g1 * n_ops + i) * 4, *v = qs + (g2 * n_ops + j) * 4;
                    volatile float p0 = u[0] * v[0], p1 = u[1] * v[1], p2 = u[2] * v[2],
                                   p3 = u[3] * v[3];
                    volatile float s01 = p0 + p1;
                    volatile float s012 = s01 + p2;
                    float d = s012 + p3;
                    d = fminf(1.0f, fmaxf(-1.0f, d));
                    float half_deg = acosf(fabsf(d)) * rad2deg_f;
                    double ang = 2.0 * (double)half_deg;
                    if (ang < best) best = ang;
                }
            /* np.histogram with `tmax` uniform bins on [0, tmax]; right edge inclusive */
            if (best >= 0.0 && best <= (double)tmax) {
                int bin = (int)best; /* bin width is exactly 1 degree */
                if (bin == tmax) bin = tmax - 1;
                counts[bin] += 1;
            }
        }
    int64_t total = 0;
    for (int b = 0; b < tmax; ++b) total += counts[b];
    double Msum = 0.0;
    for (int b = 0; b < tmax; ++b) {
        double density = total > 0 ? (double)counts[b] / (double)total : NAN; /* width 1 */
        double theory = oracle_misorientations_random((double)b, (double)(b + 1), system_a, system_b);
        Msum += fabs(theory - density);
        if (hist_out) hist_out[b] = density;
        if (counts_out) counts_out[b] = counts[b];
    }
    free(counts);
    free(qs);
    return ((double)tmax / (2.0 * tmax)) * Msum; /* diagnostics.py:365-367 */
}
