/*
 * drex_b200.h -- C ABI of libdrex_b200.so: the B200 (sm_100a) CUDA implementation of the
 * PyDRex crystallographic-preferred-orientation (CPO) hot path.
 *
 * Every entry point replaces one reference interface (paths relative to the reference
 * tree, src/pydrex/...).  The reference is pure Python (numba + SciPy); its "FFI" for this
 * path is a ctypes binding, shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; all array pointers are DEVICE pointers unless the
 *    name ends in _host;
 *  - the caller owns every buffer; the library allocates nothing that outlives a call and
 *    keeps no pointer after returning; scratch comes from a caller-provided workspace
 *    whose size is queried first (drex_integrate_workspace_bytes, ...);
 *  - every function is stream-ordered on `stream` (a cudaStream_t passed as void*) and
 *    returns immediately after enqueueing; 0 = ok, <0 = error (see DREX_E_*), message via
 *    drex_last_error (thread-local);
 *  - a "system" is one (particle, mineral phase) grain ensemble of N grains;
 *  - orientations are stored SoA, component-major: o[system][c][grain], c = 3*row + col of
 *    the 3x3 direction-cosine matrix (rows = crystal a, b, c axes) -- coalesced per-grain
 *    loads; the reference's AoS [N][3][3] layout is converted by drex_pack_soa_f64;
 *  - 3x3 matrices (L, F, D, ...) are row-major double[9].
 */
#ifndef DREX_B200_H
#define DREX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DREX_VERSION_MAJOR 0
#define DREX_VERSION_MINOR 2

/* error codes */
#define DREX_OK 0
#define DREX_E_INVALID (-1)   /* bad argument (ValueError on the Python side) */
#define DREX_E_UNSUPPORTED (-2) /* unsupported regime / phase / fabric (ValueError) */
#define DREX_E_CUDA (-3)      /* CUDA runtime error */
#define DREX_E_WORKSPACE (-4) /* workspace too small */
#define DREX_E_NODEVICE (-5)  /* no sm_100 device */

/* enums mirror core.MineralPhase / MineralFabric / DeformationRegime (core.py:35-118) */
#define DREX_PHASE_OLIVINE 0
#define DREX_PHASE_ENSTATITE 1
#define DREX_REGIME_MIN_VISCOSITY 0
#define DREX_REGIME_MATRIX_DIFFUSION 1
#define DREX_REGIME_MATRIX_DISLOCATION 4
#define DREX_REGIME_FRICTIONAL_YIELDING 6
#define DREX_REGIME_MAX_VISCOSITY 7

/* velocity-gradient providers for the integrator (replace the Python callables
 * get_velocity_gradient(t, get_position(t)) of minerals.py:388-392) */
#define DREX_L_CONSTANT 0  /* K = 1: L(t) = nodes[0] */
#define DREX_L_LINEAR 1    /* K = 2: linear in time between L(t0) and L(t1) */
#define DREX_L_CHEBYSHEV 2 /* K >= 2 Chebyshev-Lobatto nodes on [t0, t1], barycentric */
#define DREX_L_PIECEWISE 3 /* K >= 2 uniform nodes on [t0, t1], piecewise linear */

/* integrator status per system */
#define DREX_STATUS_OK 0
#define DREX_STATUS_MAX_STEPS 1
#define DREX_STATUS_STEP_UNDERFLOW 2
#define DREX_STATUS_NOT_FINITE 3

/* ---- library ---------------------------------------------------------------------- */

int drex_version(int *major, int *minor);
/* copies the calling thread's last error message (NUL-terminated) into buf */
int drex_last_error(char *buf, size_t n);
/* sm count / compute capability of `device`; DREX_E_NODEVICE unless cc >= 10.0 */
int drex_device_info(int device, int *sm_count, int *cc_major, int *cc_minor);

/* ---- layout ----------------------------------------------------------------------- */

/* AoS [S][N][9] (reference layout, minerals.py:221-222) -> SoA [S][9][N] and back */
int drex_pack_soa_f64(const double *aos, double *soa, int64_t S, int64_t N, void *stream);
int drex_unpack_soa_f64(const double *soa, double *aos, int64_t S, int64_t N, void *stream);

/* Fluidity particle-attribute wire format (examples/fluidity/corner2d/corner2d.flml:98-175):
 * one row of 10 N + 22 doubles per particle,
 *   [ orientations N x 3 x 3 | fractions N | strain | position 3 | L_prev 9 | F 9 ].
 * unpack: rows cpo[P][10N+22] -> SoA state, F, previous velocity gradient, strain.
 * pack:   the inverse; writes position and L_new into the row and advances the strain by
 *         |dt| * max|eig sym(L_new)| (utils.strain_increment, utils.py:143-156). */
int drex_fluidity_unpack_f64(const double *cpo, int64_t P, int64_t N, double *o_soa, double *f,
                             double *F, double *L_prev, double *strain, void *stream);
int drex_fluidity_pack_f64(const double *o_soa, const double *f, const double *F,
                           const double *strain, const double *position, const double *L_new,
                           double dt, int64_t P, int64_t N, double *cpo, void *stream);

/* ---- core.derivatives (core.py:347-474) ------------------------------------------- */

/* Scalar parameters of the D-Rex model shared by all systems of a call
 * (params dict of minerals.py:441-445, 470). */
typedef struct drex_params {
    double stress_exponent;       /* p */
    double deformation_exponent;  /* n */
    double nucleation_efficiency; /* lambda* */
    double gbm_mobility;          /* M* */
    double gbs_threshold;         /* chi; used by drex_integrate_f64 only */
} drex_params_t;

size_t drex_derivatives_workspace_bytes(int64_t S, int64_t N);

/* Batched core.derivatives: for each system s of S,
 *   (orientations_diff, fractions_diff) = derivatives(regime[s], phase[s], fabric[s], N,
 *        o[s], f[s], D[s], L[s], spin[s], p, n, lambda, M, vol_frac[s]).
 * o_soa/do_soa: [S][9][N]; f/df: [S][N]; D, L, spin: [S][9] (D, L already divided by the
 * max strain rate as at minerals.py:438-439); regime/phase/fabric: int32 [S];
 * vol_frac: [S].  Unsupported regimes (2, 3, 5) or (phase, fabric) pairs are rejected
 * from the HOST copies regime_host/phase_host/fabric_host (int32 [S]) before launch. */
int drex_derivatives_f64(const double *o_soa, const double *f, const double *D, const double *L,
                         const double *spin, const int32_t *regime, const int32_t *phase,
                         const int32_t *fabric, const double *vol_frac,
                         const int32_t *regime_host, const int32_t *phase_host,
                         const int32_t *fabric_host, const drex_params_t *params, int64_t S,
                         int64_t N, double *do_soa, double *df, void *workspace,
                         size_t workspace_bytes, void *stream);

/* Per-grain intermediates of core._get_rotation_and_strain (core.py:687-760), i.e. what
 * the reference's private helpers _get_slip_invariants (571-598), _get_slip_rates_olivine
 * (541-568), _get_deformation_rate (477-501), _get_slip_rate_softest (504-538) and
 * _get_strain_energy (645-684) return, for G grains of one (phase, fabric).
 * o_aos: [G][9]; D, L: [9]; outputs (any may be NULL): invariants [G][4], slip_rates [G][4],
 * deformation_rate [G][9], slip_rate_softest [G], strain_energy [G],
 * orientation_change [G][9]. */
int drex_grain_probe_f64(const double *o_aos, const double *D, const double *L, int phase,
                         int fabric, const drex_params_t *params, int64_t G, double *invariants,
                         double *slip_rates, double *deformation_rate, double *slip_rate_softest,
                         double *strain_energy, double *orientation_change, void *stream);

/* The reference's private per-grain helpers, callable on arbitrary inputs (the reference
 * tests call them directly, tests/test_core.py:142-164,753-762).  One grain per call; all
 * pointers are device pointers; `out` receives 4, 4, 9, 1, 9 or 1 doubles.
 *   op 0  _get_slip_invariants(strain_rate=in0[9], orientation=in1[9])        core.py:571-598
 *   op 1  _get_slip_rates_olivine(invariants=in0[4], slip_indices=idx[4],
 *                                 crss=in1[4], deformation_exponent=s0)       core.py:541-568
 *   op 2  _get_deformation_rate(orientation=in0[9], slip_rates=in1[4])        core.py:477-501
 *   op 3  _get_slip_rate_softest(deformation_rate=in0[9], velocity_gradient=in1[9])  504-538
 *   op 4  _get_orientation_change(orientation=in0[9], velocity_gradient=in1[9],
 *                                 deformation_rate=in2[9], slip_rate_softest=s0)     601-642
 *   op 5  _get_strain_energy(crss=in0[4], slip_rates=in1[4], slip_rate_softest=s0,
 *                            stress_exponent=s1, deformation_exponent=s2,
 *                            nucleation_efficiency=s3)                        core.py:645-684 */
int drex_grain_helper_f64(int op, const double *in0, const double *in1, const double *in2,
                          const int32_t *idx, double s0, double s1, double s2, double s3,
                          double *out, void *stream);

/* ---- Mineral.update_orientations (minerals.py:312-541) ---------------------------- */

typedef struct drex_tolerances {
    /* error weight of component i: atol_abs + (atol_rel + rtol) * |y_new_i|;
     * reference default (minerals.py:523-524): rtol 1e-6, atol 1e-4 + 1e-6 |y0_i|;
     * weighted MAX norm as in ODEPACK lsoda, taken per chunk of 32 grains (see below) */
    double rtol;
    double atol_abs;        /* orientation components */
    double atol_rel;
    double atol_abs_F;      /* the nine deformation-gradient components (LSODA takes one atol per
                               component, minerals.py:523; the blocks of y may differ) */
    double atol_abs_f;      /* the N volume fractions */
    double first_step_frac; /* first trial step = frac * |t1 - t0| when h_io is NULL or <= 0;
                               0 selects the whole interval */
    double max_step_frac;   /* cap on the step as a fraction of |t1 - t0| (<= 0: none) */
    int32_t max_steps;      /* attempted steps per chunk before DREX_STATUS_MAX_STEPS */
    int32_t method;         /* 0: Bogacki-Shampine 3(2) with FSAL (the only method) */
} drex_tolerances_t;

/* Workspace of drex_integrate_f64.  in_place != 0: the call may be made with o_out == o_in
 * (the library then keeps a copy of the interval-start orientations in the workspace, which
 * grain-boundary sliding needs: S * 9 * N doubles more). */
size_t drex_integrate_workspace_bytes(int64_t S, int64_t N, int device, int in_place);

/* Grains per integration chunk (one warp's unit of work, with its own step sequence): the
 * second dimension of h_io is ceil(N / drex_integrate_chunk_grains()). */
int32_t drex_integrate_chunk_grains(void);

/* Advance S systems from t0 to t1 (per particle), then apply grain-boundary sliding once
 * against the interval-start orientations (minerals.py:464-474, utils.apply_gbs
 * utils.py:159-180) and the final clip/renormalise (utils.extract_vars utils.py:183-190).
 *
 * The ODE of minerals.py:388-453 is solved in its decoupled form: the volume fractions follow
 * a replicator equation, f_g(t) ~ f_g(t0) exp(v_g(t)) with dv_g/dt = -s vf M* E_g, so every
 * grain is integrated on its own (nine direction cosines and v_g) and one normalisation at t1
 * recovers the fractions.  Grains are stepped in chunks of 32 (chunk q = grains 32 q ... 32 q +
 * 31), each chunk with its own adaptive steps under the weighted max norm of its components;
 * the deformation gradient is integrated separately with the same tolerances.
 *
 * sys_particle [S]  index into the per-particle arrays t0/t1/L_nodes
 * sys_phase/sys_fabric/sys_regime [S]  int32; *_host copies are validated before launch
 * sys_vol_frac [S]  phase volume fraction (minerals.py:412-415)
 * F [S][9]          deformation gradient, in/out (each phase integrates it, minerals.py:426)
 * o_in/f_in         state at t0; o_out/f_out state at t1.  f_out may alias f_in; o_out may
 *                   alias o_in if the workspace was sized with in_place != 0 (costs one
 *                   device-to-device copy of the orientations per call)
 * o_prev            orientations that grain-boundary sliding restores (utils.py:170); NULL:
 *                   o_in.  Differs from o_in when a caller splits one reference interval into
 *                   several calls (a deformation regime that changes inside it): the inner
 *                   calls run with gbs_threshold = 0, the last one with o_prev = the
 *                   orientations at the start of the whole interval.  Must not alias o_out.
 * L_kind, K, L_nodes [P][K][9]  velocity-gradient provider (DREX_L_*)
 * h_io [S][ceil(N/32)]  per chunk, in: first trial step (0: use first_step_frac); out: next
 *                   suggested step; may be NULL.  Steps are carried as STRAIN increments (time
 *                   step x max |eig sym L| at the interval end), so that a suggestion made at
 *                   one strain rate is still right when the next interval is faster or slower
 * order [S]         optional (may be NULL) permutation of 0..S-1: the order in which the
 *                   persistent CTAs pull systems from the work queue.  Results do not
 *                   depend on it; putting expensive systems first (longest-processing-
 *                   time-first) shortens the tail of the launch.
 * status [S]        int32 output: the worst DREX_STATUS_* over the system's chunks and its
 *                   deformation gradient
 * n_rhs, n_accept, n_reject [S]  int32 outputs (may be NULL): right-hand-side evaluations,
 *                   accepted and rejected steps SUMMED OVER THE SYSTEM'S CHUNKS
 * grain_evals [S]   int64 output (may be NULL): per-grain right-hand sides actually evaluated,
 *                   i.e. sum over chunks of (evaluations x grains in the chunk)
 */
/* The `order` argument for the NEXT call from this call's counts: systems sorted by measured cost,
 * most expensive first (cost = grain_evals[s] x 4 for olivine, x 1 for enstatite: the relative
 * price of their right-hand sides), ties by index.  One small launch (S <= 16384; larger S:
 * DREX_E_UNSUPPORTED, sort on the caller's side).  Stream-ordered like everything else, so the
 * array the previous call read can be overwritten in place. */
int drex_integrate_next_order(const int64_t *grain_evals, const int32_t *sys_phase, int64_t S,
                              int32_t *order, void *stream);

int drex_integrate_f64(int64_t S, int64_t P, int64_t N, const int32_t *sys_particle,
                       const int32_t *sys_phase, const int32_t *sys_fabric,
                       const int32_t *sys_regime, const double *sys_vol_frac,
                       const int32_t *sys_phase_host, const int32_t *sys_fabric_host,
                       const int32_t *sys_regime_host, double *F, const double *o_in,
                       const double *f_in, const double *o_prev, double *o_out, double *f_out,
                       const double *t0, const double *t1, int L_kind, int K,
                       const double *L_nodes, const drex_params_t *params,
                       const drex_tolerances_t *tol, double *h_io, const int32_t *order,
                       int32_t *status, int32_t *n_rhs, int32_t *n_accept, int32_t *n_reject,
                       int64_t *grain_evals, void *workspace, size_t workspace_bytes,
                       void *stream);

/* ---- minerals.voigt_averages (minerals.py:688-740) -------------------------------- */

/* Cbar[s] (+)= phase_fraction[s] * sum_g f[s][g] * voigt(rotate(C(stiffness[s]), o[s][g]^T))
 * with tensors.rotate / elastic_tensor_to_voigt semantics (tensors.py:187-273).
 * stiffness: [S][36] 6x6 Voigt matrices (GPa); Cbar: [S][36]; accumulate != 0 adds to Cbar. */
int drex_voigt_average_f64(const double *o_soa, const double *f, const double *stiffness,
                           const double *phase_fraction, int64_t S, int64_t N, double *Cbar,
                           int accumulate, void *stream);

/* diagnostics.elasticity_components (diagnostics.py:38-182; Browaeys & Chevrot 2004) for P
 * Voigt matrices voigt[P][36] (upper triangle is used).  out[P][11] = bulk modulus, shear
 * modulus, % anisotropy, % hexagonal, % tetragonal, % orthorhombic, % monoclinic,
 * % triclinic, hexagonal axis (3; defined up to sign like the reference's LAPACK vectors). */
int drex_elasticity_components_f64(const double *voigt, int64_t P, double *out, void *stream);

/* Scatter-matrix statistics of one crystallographic axis (row 0/1/2 = a/b/c) of S textures:
 * stats._scatter_matrix (stats.py:67-80) -> diagnostics.symmetry_pgr (diagnostics.py:233-270)
 * and diagnostics.bingham_average (diagnostics.py:185-213).  out[S][9] = P, G, R, mean axis
 * (unit eigenvector of the largest eigenvalue, sign arbitrary), eigenvalues descending. */
int drex_axis_scatter_f64(const double *o_soa, int64_t S, int64_t N, int row, double *out,
                          void *stream);

/* ---- utils.py / geometry.py helpers by name ------------------------------------------ */

/* utils.extract_vars (utils.py:183-190) for S state vectors y[S][9 + 10 N]: F [S][9],
 * orientations [S][N][9] clipped to [-1, 1], fractions [S][N] clipped at 0 and normalised. */
int drex_extract_vars_f64(const double *y, int64_t S, int64_t N, double *F, double *o_aos, double *f,
                          void *stream);
/* utils.apply_gbs (utils.py:159-180), in place on o_aos [S][N][9] and f [S][N]: grains with
 * f < gbs_threshold / N take their orientation from o_prev_aos and that size; f is
 * normalised afterwards. */
int drex_apply_gbs_f64(double *o_aos, double *f, const double *o_prev_aos, double gbs_threshold,
                       int64_t S, int64_t N, void *stream);
/* geometry.misorientation_angles (geometry.py:75-118): q1 [M][A][4], q2 [M][B][4] (float32
 * when is_float32, else float64; scalar-last or scalar-first alike) -> angles [M] float64,
 * the smallest 2 acos|q1_i . q2_j| in degrees, evaluated in the input precision. */
int drex_misorientation_angles(const void *q1, const void *q2, int is_float32, int64_t M, int A,
                               int B, double *angles, void *stream);
/* max |eigenvalue| of (L + L^T) / 2 for P matrices L [P][9]: the eigvalsh call of
 * minerals.py:421-422 and of utils.strain_increment (utils.py:144-156). */
int drex_strain_rate_max_f64(const double *L, int64_t P, double *out, void *stream);

/* ---- tensors.py helpers (tensors.py:14-21, 167-203, 254-273) ---------------------- */

/* B items at a time, row-major (in1 is only read by op 0):
 *   op 0  tensors.rotate(tensor [81], rotation [9]) -> [81]                  tensors.py:254-273
 *   op 1  tensors.voigt_to_elastic_tensor(matrix [36]) -> [81]               tensors.py:167-184
 *   op 2  tensors.elastic_tensor_to_voigt(tensor [81]) -> [36]               tensors.py:187-203
 *   op 3  tensors.voigt_decompose(matrix [36]) -> [18]: dilatational stiffness d_ij = C_ijkk,
 *         then deviatoric stiffness v_ij = C_ijkj                            tensors.py:39-73
 *   op 4 / 5 / 6 / 7  tensors.mono_project / ortho_project / tetr_project / hex_project
 *         (voigt_vector [21]) -> [21]                                        tensors.py:76-140
 *   op 8  tensors.voigt_matrix_to_vector(matrix [36]) -> [21]                tensors.py:206-219
 *   op 9  tensors.voigt_vector_to_matrix(vector [21]) -> [36]                tensors.py:222-251
 *   op 10 / 11  tensors.upper_tri_to_symmetric of 3x3 [9] / 6x6 [36] matrices  tensors.py:143-164 */
int drex_tensor_ops_f64(int op, const double *in0, const double *in1, int64_t B, double *out,
                        void *stream);
/* tensors.polar_decompose for B nonsingular 3x3 matrices: left != 0 gives M = V R
 * (rotation, stretch) = (R, V = sqrt(M M^T)), left == 0 gives M = R U with U = sqrt(M^T M). */
int drex_polar_decompose_f64(const double *matrices, int64_t B, int left, double *rotation,
                             double *stretch, void *stream);

/* ---- stats.resample_orientations (stats.py:14-64) ----------------------------------- */

/* Volume-weighted resampling of S textures of N grains: sort by fraction ascending (ties by
 * grain index, i.e. numpy.argsort(kind="stable")), sequential running sum with the last
 * entry forced to 1, numpy.searchsorted(cumfrac, u) per uniform variate.  o_aos: [S][N][3][3]
 * (the reference's layout), f: [S][N], uniforms: [S][n_samples] in [0,1) -- the caller draws
 * them (numpy.random.default_rng(seed).random((S, n_samples)) reproduces the reference's
 * stream).  out_o: [S][n_samples][3][3], out_f: [S][n_samples].  Workspace is only needed
 * above 11000 grains per texture (query first; 0 otherwise). */
size_t drex_resample_workspace_bytes(int64_t S, int64_t N);
int drex_resample_orientations_f64(const double *o_aos, const double *f, const double *uniforms,
                                   int64_t S, int64_t N, int64_t n_samples, double *out_o,
                                   double *out_f, void *workspace, size_t workspace_bytes,
                                   void *stream);

/* The uniform variates of stats.resample_orientations on the device (stats.py:49, 59):
 * out[t][0..n) = numpy.random.default_rng(seeds[t]).random(offsets[t] + n)[offsets[t]:], bit for
 * bit (SeedSequence hashing of the seed's 32-bit words, PCG64, 53-bit doubles).  seeds: [T]
 * non-negative integers below 2^64; offsets: [T] or NULL (all 0): cut ONE seeded stream into
 * per-texture pieces with seeds[t] = seed, offsets[t] = t * n, as the reference's loop over
 * textures draws them.  Device pointers. */
int drex_numpy_uniform_f64(const uint64_t *seeds, const uint64_t *offsets, int64_t T, int64_t n,
                           double *out, void *stream);

/* ---- diagnostics.misorientation_index (diagnostics.py:332-367) -------------------- */

size_t drex_mindex_workspace_bytes(int64_t S, int64_t N, int system_a, int system_b);

/* Closed-form random-texture density per 1-degree bin (stats.misorientations_random,
 * stats.py:130-200), host-side; theory_host: [theta_max].  Returns theta_max or <0. */
int drex_mindex_theory_host(int system_a, int system_b, double *theory_host);
/* stats.misorientations_random(low, high, system) for arbitrary bin edges in degrees;
 * DREX_E_INVALID for bad bounds, DREX_E_UNSUPPORTED where the reference asserts. */
int drex_misorientations_random_host(double low, double high, int system_a, int system_b,
                                     double *density_host);

/* M-index of S textures of N grains each.  (system_a, system_b) = LatticeSystem.value
 * (geometry.py:29-34).  o_soa: [S][9][N] float64.  M: [S] float64.  counts (optional):
 * [S][theta_max] int64 histogram of pair misorientation angles in 1-degree bins.
 * Quaternions are float32 and the symmetry operations are applied exactly as
 * stats.misorientation_hist does (stats.py:105-123), including utils.quat_product's
 * missing cross term (utils.py:365-371). */
int drex_mindex_f32(const double *o_soa, int64_t S, int64_t N, int system_a, int system_b,
                    double *M, int64_t *counts, void *workspace, size_t workspace_bytes,
                    void *stream);

/* ---- steady flows and pathlines (velocity.py:17-97, pathlines.py:10-116) ---------- */

#define DREX_FLOW_SIMPLE_SHEAR 0 /* velocity.simple_shear_2d: i0 = direction, i1 = deformation
                                    plane (axis indices 0..2), a = strain_rate */
#define DREX_FLOW_CELL 1         /* velocity.cell_2d: i0 = horizontal, i1 = vertical,
                                    a = velocity_edge, b = edge_length */
#define DREX_FLOW_CORNER 2       /* velocity.corner_2d: i0 = horizontal, i1 = vertical,
                                    a = plate_speed */
typedef struct drex_flow_field {
    int32_t kind, i0, i1;
    double a, b;
} drex_flow_field_t;

#define DREX_PATH_OK 0
#define DREX_PATH_NOT_FINITE 1 /* the field is NaN on the path (corner origin, outside the cell) */
#define DREX_PATH_MAX_STEPS 2
#define DREX_PATH_EMPTY 3      /* final location outside the box, or no strain budget */
#define DREX_PATH_TIME_LIMIT 4 /* no event before t_limit: the path starts there (usable) */

/* get_velocity(t, x) / get_velocity_gradient(t, x) of the analytic flows at P points
 * x[P][3]; v[P][3] and L[P][9] (row-major 3x3), either may be NULL.  Points where the
 * reference returns NaN or raises (corner origin, outside the cell) give NaN. */
int drex_flow_field_f64(const drex_flow_field_t *field, const double *x, int64_t P, double *v,
                        double *L, void *stream);

/* pathlines.get_pathline for P particles at once: trace each final location backwards in
 * the steady flow until `max_strain` (tensorial strain, utils.strain_increment) has been
 * accumulated, the path leaves the box [min_coords, max_coords] (host arrays of 3), or
 * t_limit (< 0; the reference uses -100 Myr in seconds) is reached.  Dormand-Prince 5(4),
 * error weights atol + rtol |y| on position and strain alike.
 *   timestamps_given = 0: timestamps[j][p] = linspace(t_start[p], 0, n_intervals + 1)[j]
 *   timestamps_given = 1: the caller's ascending stamps (<= 0) are used as they are
 * Outputs are interval-major: timestamps [n_intervals+1][P], positions [n_intervals+1][P][3],
 * strains [n_intervals+1][P] (accumulated since the first stamp; may be NULL), and -- when
 * n_nodes >= 2 -- node_positions [n_intervals][P][n_nodes][3] and L_nodes
 * [n_intervals][P][n_nodes][9] at the Chebyshev-Lobatto nodes of every interval, i.e. block j
 * of L_nodes is the DREX_L_CHEBYSHEV table drex_integrate_f64 takes for the step from
 * timestamps[j] to timestamps[j+1] (either may be NULL).  step_times [P][max_record]
 * (optional) receives the accepted solver times of the first pass, descending, closed by
 * t_start; n_taken [P] (optional) the number of times recorded.  status [P]: DREX_PATH_*. */
int drex_pathlines_f64(const drex_flow_field_t *field, const double *final_locations, int64_t P,
                       const double *min_coords, const double *max_coords, double max_strain,
                       double t_limit, double rtol, double atol, int64_t n_intervals,
                       int32_t n_nodes, int32_t timestamps_given, int32_t max_steps,
                       double *t_start, double *timestamps, double *positions, double *strains,
                       double *node_positions, double *L_nodes, double *step_times,
                       int32_t max_record, int32_t *n_taken, int32_t *status, void *stream);

/* ---- measurement helper ----------------------------------------------------------- */

/* Runs a dependent-DFMA loop that saturates the FP64 pipe and returns the measured FP64
 * rate in FLOP/s (2 per DFMA) in *flops_host -- the denominator for the FP64 roofline,
 * which MEASURED_PEAKS.json does not carry. */
int drex_fp64_peak_probe(int device, double *flops_host);

/* Compute-only ceiling of the per-grain physics: olivine right-hand sides per second when
 * every thread re-evaluates register-resident data at the integrator's occupancy (512
 * threads per SM): what the integrator's stage loop could reach with nothing else in the
 * way.  Measurement aid like drex_fp64_peak_probe. */
int drex_grain_rhs_ceiling_probe(int device, double *evals_per_s_host);

#ifdef __cplusplus
}
#endif
#endif /* DREX_B200_H */
