// drex_capi.cu -- the C ABI of libdrex_b200.so (declared in include/drex_b200.h).
//
// Host-side argument checking and kernel launches only; every numeric result comes from
// the sm_100a kernels in the .cuh files next to this one.  There is no CPU fallback: a
// call without a usable device returns DREX_E_CUDA / DREX_E_NODEVICE.
#include "../../include/drex_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "drex_derivatives.cuh"
#include "drex_diagnostics.cuh"
#include "drex_integrate.cuh"
#include "drex_pathlines.cuh"
#include "drex_resample.cuh"

namespace {

thread_local char g_error[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

#define DREX_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(DREX_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));      \
    } while (0)

// core.get_crss (core.py:315-342)
bool crss_for(int phase, int fabric, double *crss) {
    static const double table[6][4] = {{1, 2, 3, INFINITY}, {3, 2, 1, INFINITY},
                                       {3, 2, INFINITY, 1}, {1, 1, 3, INFINITY},
                                       {3, 1, 2, INFINITY}, {INFINITY, INFINITY, INFINITY, 1}};
    const bool ok = (phase == 0 && fabric >= 0 && fabric <= 4) || (phase == 1 && fabric == 5);
    if (ok) memcpy(crss, table[fabric], 4 * sizeof(double));
    return ok;
}

drex::FabricConsts make_fabric_consts(int fabric, const drex_params_t &p) {
    drex::FabricConsts fc;
    memset(&fc, 0, sizeof(fc));
    const int phase = (fabric == 5) ? 1 : 0;
    double crss[4];
    crss_for(phase, fabric, crss);
    for (int i = 0; i < 4; ++i) {
        fc.crss[i] = crss[i];
        fc.inv_crss[i] = 1.0 / crss[i];
    }
    for (int i = 0; i < 3; ++i)  // core.py:675-676
        fc.rho_pref[i] = std::pow(1.0 / crss[i], p.deformation_exponent - p.stress_exponent);
    fc.n_minus_1 = p.deformation_exponent - 1.0;
    fc.p = p.stress_exponent;
    fc.p_over_n = p.stress_exponent / p.deformation_exponent;
    fc.lambda = p.nucleation_efficiency;
    fc.phase = phase;
    fc.default_exponents = (p.stress_exponent == 1.5 && p.deformation_exponent == 3.5) ? 1 : 0;
    fc.inf_index = 3;
    for (int i = 0; i < 4; ++i)
        if (std::isinf(crss[i]) && phase == 0) fc.inf_index = i;
    return fc;
}

drex::FabricTable make_fabric_table(const drex_params_t &p) {
    drex::FabricTable t;
    for (int fab = 0; fab < 6; ++fab) t.fab[fab] = make_fabric_consts(fab, p);
    return t;
}

int validate_systems(int64_t S, const int32_t *regime, const int32_t *phase, const int32_t *fabric) {
    for (int64_t s = 0; s < S; ++s) {
        const int r = regime[s];
        if (r == 2 || r == 3 || r == 5)  // core.py:406-413,437-438
            return fail(DREX_E_UNSUPPORTED, "this deformation mechanism is not yet supported.");
        if (r < 0 || r > 7)  // core.py:473-474
            return fail(DREX_E_INVALID, "regime must be a valid `DeformationRegime`, not %d", r);
        double crss[4];
        if (!crss_for(phase[s], fabric[s], crss)) {
            if (phase[s] == 0)
                return fail(DREX_E_UNSUPPORTED, "unsupported olivine fabric: %d", fabric[s]);
            if (phase[s] == 1)
                return fail(DREX_E_UNSUPPORTED, "unsupported enstatite fabric: %d", fabric[s]);
            return fail(DREX_E_UNSUPPORTED, "phase must be a valid `MineralPhase`, not %d", phase[s]);
        }
    }
    return DREX_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// The integrator runs one CTA per SM (128 registers x 512 threads fill the register file).
int integrate_slots(int device) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return sms * drex::INT_CTAS_PER_SM;
}

// Dynamic shared memory of the integrator: the per-warp state / prefetch rows and stage flows.
inline size_t integrate_smem_bytes() {
    return (size_t)drex::INT_WARPS * (drex::WARP_DOUBLES * sizeof(double) + 3 * sizeof(drex::ChunkFlow));
}

// Per-CTA scratch: the per-chunk partial sums of the systems in flight.
inline int64_t integrate_chunk_cap(int64_t N) {
    return ((N + drex::INT_CHUNK - 1) / drex::INT_CHUNK + 31) / 32 * 32;
}
// One lane mask per chunk of every system: which grains slide (written by integrate_kernel, read
// by gbs_restore_kernel).
inline size_t integrate_mask_bytes(int64_t S, int64_t N) {
    return align_up((size_t)S * (size_t)((N + 31) / 32) * sizeof(uint32_t), 256);
}
inline size_t integrate_scratch_bytes(int slots, int64_t N) {
    return align_up((size_t)slots * drex::INT_NCTX * (size_t)integrate_chunk_cap(N) * sizeof(double), 256);
}

}  // namespace

extern "C" {

int drex_version(int *major, int *minor) {
    if (major) *major = DREX_VERSION_MAJOR;
    if (minor) *minor = DREX_VERSION_MINOR;
    return DREX_OK;
}

int drex_last_error(char *buf, size_t n) {
    if (!buf || n == 0) return DREX_E_INVALID;
    strncpy(buf, g_error, n - 1);
    buf[n - 1] = '\0';
    return DREX_OK;
}

int drex_device_info(int device, int *sm_count, int *cc_major, int *cc_minor) {
    cudaDeviceProp prop;
    DREX_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (prop.major < 10)
        return fail(DREX_E_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a",
                    device, prop.major, prop.minor);
    return DREX_OK;
}

int drex_pack_soa_f64(const double *aos, double *soa, int64_t S, int64_t N, void *stream) {
    if (S <= 0 || N <= 0) return DREX_OK;
    if (!aos || !soa) return fail(DREX_E_INVALID, "drex_pack_soa_f64: null pointer");
    dim3 grid((unsigned)S, (unsigned)((N + drex::LAYOUT_TILE - 1) / drex::LAYOUT_TILE));
    drex::pack_soa_kernel<<<grid, drex::LAYOUT_TILE, 0, (cudaStream_t)stream>>>(aos, soa, S, N);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_unpack_soa_f64(const double *soa, double *aos, int64_t S, int64_t N, void *stream) {
    if (S <= 0 || N <= 0) return DREX_OK;
    if (!aos || !soa) return fail(DREX_E_INVALID, "drex_unpack_soa_f64: null pointer");
    dim3 grid((unsigned)S, (unsigned)((N + drex::LAYOUT_TILE - 1) / drex::LAYOUT_TILE));
    drex::unpack_soa_kernel<<<grid, drex::LAYOUT_TILE, 0, (cudaStream_t)stream>>>(soa, aos, S, N);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_fluidity_unpack_f64(const double *cpo, int64_t P, int64_t N, double *o_soa, double *f,
                             double *F, double *L_prev, double *strain, void *stream) {
    if (P < 0 || N < 0) return fail(DREX_E_INVALID, "negative size");
    if (P == 0) return DREX_OK;
    if (!cpo || !o_soa || !f || !F || !L_prev || !strain)
        return fail(DREX_E_INVALID, "drex_fluidity_unpack_f64: null pointer");
    dim3 grid((unsigned)P, (unsigned)((N + 255) / 256 > 0 ? (N + 255) / 256 : 1));
    drex::fluidity_unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cpo, N, o_soa, f, F, L_prev, strain);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_fluidity_pack_f64(const double *o_soa, const double *f, const double *F,
                           const double *strain, const double *position, const double *L_new,
                           double dt, int64_t P, int64_t N, double *cpo, void *stream) {
    if (P < 0 || N < 0) return fail(DREX_E_INVALID, "negative size");
    if (P == 0) return DREX_OK;
    if (!cpo || !o_soa || !f || !F || !strain || !position || !L_new)
        return fail(DREX_E_INVALID, "drex_fluidity_pack_f64: null pointer");
    dim3 grid((unsigned)P, (unsigned)((N + 255) / 256 > 0 ? (N + 255) / 256 : 1));
    drex::fluidity_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(o_soa, f, F, strain, position,
                                                                       L_new, dt, N, cpo);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

size_t drex_derivatives_workspace_bytes(int64_t S, int64_t N) {
    return (size_t)(S > 0 ? S : 0) * (size_t)(N > 0 ? N : 0) * sizeof(double);
}

int drex_derivatives_f64(const double *o_soa, const double *f, const double *D, const double *L,
                         const double *spin, const int32_t *regime, const int32_t *phase,
                         const int32_t *fabric, const double *vol_frac,
                         const int32_t *regime_host, const int32_t *phase_host,
                         const int32_t *fabric_host, const drex_params_t *params, int64_t S,
                         int64_t N, double *do_soa, double *df, void *workspace,
                         size_t workspace_bytes, void *stream) {
    if (S < 0 || N < 0) return fail(DREX_E_INVALID, "negative size");
    if (!params || !regime_host || !phase_host || !fabric_host)
        return fail(DREX_E_INVALID, "drex_derivatives_f64: null host argument");
    int rc = validate_systems(S, regime_host, phase_host, fabric_host);
    if (rc != DREX_OK) return rc;
    if (S == 0 || N == 0) return DREX_OK;
    if (!o_soa || !f || !D || !L || !spin || !regime || !phase || !fabric || !vol_frac || !do_soa || !df)
        return fail(DREX_E_INVALID, "drex_derivatives_f64: null device pointer");
    if (workspace_bytes < drex_derivatives_workspace_bytes(S, N) || !workspace)
        return fail(DREX_E_WORKSPACE, "drex_derivatives_f64: workspace too small (%zu < %zu)",
                    workspace_bytes, drex_derivatives_workspace_bytes(S, N));
    drex::DerivArgs A;
    A.o = o_soa; A.f = f; A.D = D; A.L = L; A.spin = spin; A.vol_frac = vol_frac;
    A.regime = regime; A.phase = phase; A.fabric = fabric;
    A.d_o = do_soa; A.d_f = df; A.energy = (double *)workspace;
    A.gbm_mobility = params->gbm_mobility;
    A.S = S; A.N = N;
    const drex::FabricTable T = make_fabric_table(*params);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t per_block = (int64_t)drex::DERIV_BLOCK * drex::DERIV_TILES;
    dim3 grid((unsigned)S, (unsigned)((N + per_block - 1) / per_block));
    drex::derivatives_grain_kernel<<<grid, drex::DERIV_BLOCK, 0, st>>>(A, T);
    DREX_CUDA_CHECK(cudaGetLastError());
    drex::derivatives_gbm_kernel<<<(unsigned)S, drex::GBM_BLOCK, 0, st>>>(A);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_grain_probe_f64(const double *o_aos, const double *D, const double *L, int phase,
                         int fabric, const drex_params_t *params, int64_t G, double *invariants,
                         double *slip_rates, double *deformation_rate, double *slip_rate_softest,
                         double *strain_energy, double *orientation_change, void *stream) {
    if (!params) return fail(DREX_E_INVALID, "drex_grain_probe_f64: null params");
    double crss[4];
    if (!crss_for(phase, fabric, crss))
        return fail(DREX_E_UNSUPPORTED, "unsupported phase/fabric: %d/%d", phase, fabric);
    if (G <= 0) return DREX_OK;
    if (!o_aos || !D || !L) return fail(DREX_E_INVALID, "drex_grain_probe_f64: null pointer");
    drex::ProbeArgs A;
    A.o_aos = o_aos; A.D = D; A.L = L;
    A.inv = invariants; A.rates = slip_rates; A.G = deformation_rate;
    A.gamma0 = slip_rate_softest; A.energy = strain_energy; A.da = orientation_change;
    A.n = G;
    const drex::FabricConsts fc = make_fabric_consts(fabric, *params);
    drex::grain_probe_kernel<<<(unsigned)((G + 127) / 128), 128, 0, (cudaStream_t)stream>>>(A, fc);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_grain_helper_f64(int op, const double *in0, const double *in1, const double *in2,
                          const int32_t *idx, double s0, double s1, double s2, double s3,
                          double *out, void *stream) {
    if (op < 0 || op > 5) return fail(DREX_E_INVALID, "drex_grain_helper_f64: unknown op %d", op);
    if (!in0 || !in1 || !out || (op == 4 && !in2) || (op == 1 && !idx))
        return fail(DREX_E_INVALID, "drex_grain_helper_f64: null pointer");
    drex::HelperArgs A;
    A.op = op; A.in0 = in0; A.in1 = in1; A.in2 = in2; A.idx = idx;
    A.s0 = s0; A.s1 = s1; A.s2 = s2; A.s3 = s3; A.out = out;
    drex::grain_helper_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(A);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int32_t drex_integrate_chunk_grains(void) { return drex::INT_CHUNK; }

size_t drex_integrate_workspace_bytes(int64_t S, int64_t N, int device, int in_place) {
    if (S <= 0 || N <= 0) return 256;
    int slots = integrate_slots(device);
    if (slots < 1) slots = 1;
    if ((int64_t)slots > S) slots = (int)S;
    size_t need = 256 + integrate_scratch_bytes(slots, N) + integrate_mask_bytes(S, N);
    // o_out == o_in: grain-boundary sliding restores interval-start orientations that the
    // integration has already overwritten, so the call keeps a copy of them here
    if (in_place) need += (size_t)S * 9 * (size_t)N * sizeof(double);
    return need;
}

int drex_integrate_next_order(const int64_t *grain_evals, const int32_t *sys_phase, int64_t S,
                              int32_t *order, void *stream) {
    if (S < 0) return fail(DREX_E_INVALID, "negative size");
    if (S == 0) return DREX_OK;
    if (!grain_evals || !sys_phase || !order) return fail(DREX_E_INVALID, "drex_integrate_next_order: null pointer");
    if (S > drex::ORDER_MAX)
        return fail(DREX_E_UNSUPPORTED, "drex_integrate_next_order: more than %d systems", drex::ORDER_MAX);
    const size_t smem = (size_t)S * sizeof(unsigned long long);
    if (smem > 48 * 1024)
        DREX_CUDA_CHECK(cudaFuncSetAttribute(drex::next_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
    drex::next_order_kernel<<<1, drex::ORDER_THREADS, smem, (cudaStream_t)stream>>>(grain_evals, sys_phase, (int)S,
                                                                                    order);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

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
                       void *stream) {
    if (S < 0 || P < 0 || N < 0) return fail(DREX_E_INVALID, "negative size");
    if (!params || !tol || !sys_phase_host || !sys_fabric_host || !sys_regime_host)
        return fail(DREX_E_INVALID, "drex_integrate_f64: null host argument");
    int rc = validate_systems(S, sys_regime_host, sys_phase_host, sys_fabric_host);
    if (rc != DREX_OK) return rc;
    if (L_kind < 0 || L_kind > 3) return fail(DREX_E_INVALID, "unknown L_kind %d", L_kind);
    if (K < 1 || (L_kind != DREX_L_CONSTANT && K < 2) || (L_kind == DREX_L_LINEAR && K != 2))
        return fail(DREX_E_INVALID, "L_kind %d needs a different node count than K = %d", L_kind, K);
    if (!(tol->rtol >= 0.0) || !(tol->atol_abs >= 0.0) || !(tol->atol_rel >= 0.0) ||
        !(tol->atol_abs_F >= 0.0) || !(tol->atol_abs_f >= 0.0) ||
        !(tol->rtol + tol->atol_rel + fmin(tol->atol_abs, fmin(tol->atol_abs_F, tol->atol_abs_f)) > 0.0))
        return fail(DREX_E_INVALID, "tolerances must be non-negative and not all zero");
    if (tol->method != 0) return fail(DREX_E_INVALID, "unknown integration method %d", tol->method);
    if (S == 0 || N == 0) return DREX_OK;
    if (!sys_particle || !sys_phase || !sys_fabric || !sys_regime || !sys_vol_frac || !F || !o_in ||
        !f_in || !o_out || !f_out || !t0 || !t1 || !L_nodes || !status)
        return fail(DREX_E_INVALID, "drex_integrate_f64: null device pointer");
    int device = 0;
    DREX_CUDA_CHECK(cudaGetDevice(&device));
    int slots = integrate_slots(device);
    if (slots < 1) return fail(DREX_E_CUDA, "could not query the device of the integrator kernel");
    if ((int64_t)slots > S) slots = (int)S;
    const size_t state_bytes = (size_t)S * 9 * (size_t)N * sizeof(double);
    const char *in_lo = (const char *)o_in, *out_lo = (const char *)o_out;
    const bool aliased = o_prev == nullptr && in_lo < out_lo + state_bytes && out_lo < in_lo + state_bytes;
    if (o_prev != nullptr) {
        const char *pv = (const char *)o_prev;
        if (pv < out_lo + state_bytes && out_lo < pv + state_bytes)
            return fail(DREX_E_INVALID, "drex_integrate_f64: o_prev must not alias o_out");
    }
    const size_t scratch_bytes = integrate_scratch_bytes(slots, N);
    const size_t mask_bytes = integrate_mask_bytes(S, N);
    const size_t need = 256 + scratch_bytes + mask_bytes + (aliased ? state_bytes : 0);
    if (!workspace || workspace_bytes < need)
        return fail(DREX_E_WORKSPACE, "drex_integrate_f64: workspace too small (%zu < %zu)",
                    workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    drex::IntegrateArgs A;
    A.S = S; A.P = P; A.N = N;
    A.sys_particle = sys_particle; A.sys_phase = sys_phase; A.sys_fabric = sys_fabric;
    A.sys_regime = sys_regime; A.sys_vol_frac = sys_vol_frac;
    A.F = F; A.o_in = o_in; A.f_in = f_in; A.o_out = o_out; A.f_out = f_out;
    A.o_prev = o_prev ? o_prev : o_in;
    A.t0 = t0; A.t1 = t1; A.L_kind = L_kind; A.K = K; A.L_nodes = L_nodes;
    A.gbm_mobility = params->gbm_mobility;
    A.gbs_threshold = params->gbs_threshold;
    A.tol.rtol = tol->rtol; A.tol.atol_abs = tol->atol_abs; A.tol.atol_rel = tol->atol_rel;
    A.tol.atol_abs_F = tol->atol_abs_F; A.tol.atol_abs_f = tol->atol_abs_f;
    A.tol.first_step_frac = tol->first_step_frac; A.tol.max_step_frac = tol->max_step_frac;
    A.tol.max_steps = tol->max_steps > 0 ? tol->max_steps : 100000;
    A.tol.method = 0;
    A.h_io = h_io; A.order = order; A.status = status; A.n_rhs = n_rhs; A.n_accept = n_accept;
    A.n_reject = n_reject; A.grain_evals = grain_evals;
    A.queue = (int *)workspace;
    A.scratch = (double *)((char *)workspace + 256);
    A.chunk_cap = integrate_chunk_cap(N);
    A.slide_mask = (uint32_t *)((char *)workspace + 256 + scratch_bytes);
    DREX_CUDA_CHECK(cudaMemsetAsync(workspace, 0, 256, st));
    if (aliased) {
        double *copy = (double *)((char *)workspace + 256 + scratch_bytes + mask_bytes);
        DREX_CUDA_CHECK(cudaMemcpyAsync(copy, o_in, state_bytes, cudaMemcpyDeviceToDevice, st));
        A.o_in = copy;
        A.o_prev = copy;
    }
    const drex::FabricTable T = make_fabric_table(*params);
    if (const char *dbg = getenv("DREX_DEBUG_HOSTPTR")) {  // development aid, see watchdog_fire
        unsigned long long *ptr = (unsigned long long *)strtoull(dbg, nullptr, 0);
        DREX_CUDA_CHECK(cudaMemcpyToSymbolAsync(drex::g_drex_debug, &ptr, sizeof(ptr), 0,
                                                cudaMemcpyHostToDevice, st));
    }
    const size_t dyn_smem = integrate_smem_bytes();
    DREX_CUDA_CHECK(cudaFuncSetAttribute(drex::integrate_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    drex::integrate_kernel<<<(unsigned)slots, drex::INT_BLOCK, dyn_smem, st>>>(A, T);
    DREX_CUDA_CHECK(cudaGetLastError());
    if (params->gbs_threshold > 0.0) {  // nothing slides below a threshold of zero
        int sms = 0;
        DREX_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        drex::gbs_restore_kernel<<<(unsigned)(8 * sms), drex::RESTORE_BLOCK, 0, st>>>(
            A.slide_mask, A.o_prev, o_out, S, N);
        DREX_CUDA_CHECK(cudaGetLastError());
    }
    return DREX_OK;
}

int drex_voigt_average_f64(const double *o_soa, const double *f, const double *stiffness,
                           const double *phase_fraction, int64_t S, int64_t N, double *Cbar,
                           int accumulate, void *stream) {
    if (S < 0 || N < 0) return fail(DREX_E_INVALID, "negative size");
    if (S == 0) return DREX_OK;
    if (!o_soa || !f || !stiffness || !phase_fraction || !Cbar)
        return fail(DREX_E_INVALID, "drex_voigt_average_f64: null pointer");
    drex::voigt_average_kernel<<<(unsigned)S, drex::VOIGT_BLOCK, 0, (cudaStream_t)stream>>>(
        o_soa, f, stiffness, phase_fraction, N, Cbar, accumulate);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_elasticity_components_f64(const double *voigt, int64_t P, double *out, void *stream) {
    if (P < 0) return fail(DREX_E_INVALID, "negative size");
    if (P == 0) return DREX_OK;
    if (!voigt || !out) return fail(DREX_E_INVALID, "drex_elasticity_components_f64: null pointer");
    drex::elasticity_components_kernel<<<(unsigned)((P + 63) / 64), 64, 0, (cudaStream_t)stream>>>(voigt, P, out);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_axis_scatter_f64(const double *o_soa, int64_t S, int64_t N, int row, double *out,
                          void *stream) {
    if (S < 0 || N < 0) return fail(DREX_E_INVALID, "negative size");
    if (row < 0 || row > 2) return fail(DREX_E_INVALID, "axis must be 'a', 'b', or 'c'");
    if (S == 0) return DREX_OK;
    if (!o_soa || !out) return fail(DREX_E_INVALID, "drex_axis_scatter_f64: null pointer");
    drex::axis_scatter_kernel<<<(unsigned)S, 256, 0, (cudaStream_t)stream>>>(o_soa, N, row, out);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

// ---- utils.py / geometry.py helpers by name ----------------------------------------------------

int drex_extract_vars_f64(const double *y, int64_t S, int64_t N, double *F, double *o_aos, double *f,
                          void *stream) {
    if (S < 0 || N < 0) return fail(DREX_E_INVALID, "negative size");
    if (S == 0) return DREX_OK;
    if (!y || !F || !o_aos || !f) return fail(DREX_E_INVALID, "drex_extract_vars_f64: null pointer");
    drex::extract_vars_kernel<<<(unsigned)S, 256, 0, (cudaStream_t)stream>>>(y, N, F, o_aos, f);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_apply_gbs_f64(double *o_aos, double *f, const double *o_prev_aos, double gbs_threshold,
                       int64_t S, int64_t N, void *stream) {
    if (S < 0 || N < 0) return fail(DREX_E_INVALID, "negative size");
    if (S == 0 || N == 0) return DREX_OK;
    if (!o_aos || !f || !o_prev_aos) return fail(DREX_E_INVALID, "drex_apply_gbs_f64: null pointer");
    drex::apply_gbs_kernel<<<(unsigned)S, 256, 0, (cudaStream_t)stream>>>(o_aos, f, o_prev_aos,
                                                                         gbs_threshold, N);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_misorientation_angles(const void *q1, const void *q2, int is_float32, int64_t M, int A,
                               int B, double *angles, void *stream) {
    if (M < 0 || A <= 0 || B <= 0) return fail(DREX_E_INVALID, "drex_misorientation_angles: bad sizes");
    if (M == 0) return DREX_OK;
    if (!q1 || !q2 || !angles) return fail(DREX_E_INVALID, "drex_misorientation_angles: null pointer");
    const unsigned grid = (unsigned)((M + 127) / 128);
    if (is_float32)
        drex::misorientation_angles_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(
            (const float *)q1, (const float *)q2, M, A, B, angles);
    else
        drex::misorientation_angles_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(
            (const double *)q1, (const double *)q2, M, A, B, angles);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_strain_rate_max_f64(const double *L, int64_t P, double *out, void *stream) {
    if (P < 0) return fail(DREX_E_INVALID, "negative size");
    if (P == 0) return DREX_OK;
    if (!L || !out) return fail(DREX_E_INVALID, "drex_strain_rate_max_f64: null pointer");
    drex::strain_rate_max_kernel<<<(unsigned)((P + 127) / 128), 128, 0, (cudaStream_t)stream>>>(L, P, out);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

// ---- tensors.py helpers ---------------------------------------------------------------------

int drex_tensor_ops_f64(int op, const double *in0, const double *in1, int64_t B, double *out,
                        void *stream) {
    if (op < 0 || op > 11) return fail(DREX_E_INVALID, "drex_tensor_ops_f64: unknown op %d", op);
    if (B < 0) return fail(DREX_E_INVALID, "negative size");
    if (B == 0) return DREX_OK;
    if (!in0 || !out || (op == 0 && !in1)) return fail(DREX_E_INVALID, "drex_tensor_ops_f64: null pointer");
    // ops 10 / 11: upper_tri_to_symmetric of 3x3 / 6x6 matrices
    const int n_side = (op == 11) ? 6 : 3;
    if (op == 11) op = 10;
    const int64_t n = B * drex::tensor_op_n_out(op, n_side);
    drex::tensor_ops_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(op, in0, in1, B, out,
                                                                                            n_side);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_polar_decompose_f64(const double *matrices, int64_t B, int left, double *rotation,
                             double *stretch, void *stream) {
    if (B < 0) return fail(DREX_E_INVALID, "negative size");
    if (B == 0) return DREX_OK;
    if (!matrices || !rotation || !stretch) return fail(DREX_E_INVALID, "drex_polar_decompose_f64: null pointer");
    drex::polar_decompose_kernel<<<(unsigned)((B + 63) / 64), 64, 0, (cudaStream_t)stream>>>(
        matrices, B, left ? 1 : 0, rotation, stretch);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

// ---- stats.resample_orientations ---------------------------------------------------------

size_t drex_resample_workspace_bytes(int64_t S, int64_t N) {
    if (S <= 0 || N <= 0) return 0;
    return N <= drex::RESAMPLE_SMEM_N_MAX ? 0 : (size_t)S * (size_t)drex::resample_ws_stride(N);
}

int drex_resample_orientations_f64(const double *o_aos, const double *f, const double *uniforms,
                                   int64_t S, int64_t N, int64_t n_samples, double *out_o,
                                   double *out_f, void *workspace, size_t workspace_bytes,
                                   void *stream) {
    if (S < 0 || N < 0 || n_samples < 0) return fail(DREX_E_INVALID, "negative size");
    if (S == 0 || n_samples == 0) return DREX_OK;
    if (N == 0) return fail(DREX_E_INVALID, "drex_resample_orientations_f64: no grains to sample");
    if (N > 0x7ffffffe) return fail(DREX_E_INVALID, "drex_resample_orientations_f64: too many grains");
    if (!o_aos || !f || !uniforms || !out_o || !out_f)
        return fail(DREX_E_INVALID, "drex_resample_orientations_f64: null pointer");
    const size_t need = drex_resample_workspace_bytes(S, N);
    if (need && (!workspace || workspace_bytes < need))
        return fail(DREX_E_WORKSPACE, "drex_resample_orientations_f64: workspace too small (%zu < %zu)",
                    workspace_bytes, need);
    const int64_t npad = drex::resample_npad(N);
    const size_t smem = need ? 0 : (size_t)N * 20;
    if (smem > 48 * 1024)
        DREX_CUDA_CHECK(cudaFuncSetAttribute(drex::resample_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    drex::resample_kernel<<<(unsigned)S, drex::RESAMPLE_THREADS, smem, (cudaStream_t)stream>>>(
        o_aos, f, uniforms, N, npad, n_samples, out_o, out_f, (unsigned char *)workspace);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_numpy_uniform_f64(const uint64_t *seeds, const uint64_t *offsets, int64_t T, int64_t n,
                           double *out, void *stream) {
    if (T < 0 || n < 0) return fail(DREX_E_INVALID, "negative size");
    if (T == 0 || n == 0) return DREX_OK;
    if (!seeds || !out) return fail(DREX_E_INVALID, "drex_numpy_uniform_f64: null pointer");
    drex::numpy_uniform_kernel<<<(unsigned)((T + 63) / 64), 64, 0, (cudaStream_t)stream>>>(seeds, offsets, T, n,
                                                                                        out);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

// ---- steady flows and pathlines ------------------------------------------------------------

static int make_flow_field(const drex_flow_field_t *field, drex::FlowField *out) {
    if (!field) return fail(DREX_E_INVALID, "flow field: null pointer");
    if (field->kind < 0 || field->kind > 2)
        return fail(DREX_E_INVALID, "flow field: unknown kind %d", field->kind);
    if (field->i0 < 0 || field->i0 > 2 || field->i1 < 0 || field->i1 > 2 || field->i0 == field->i1)
        return fail(DREX_E_INVALID, "flow field: axes must be two different indices in 0..2");
    if (field->kind == DREX_FLOW_CELL && !(field->b > 0.0))
        return fail(DREX_E_INVALID, "edge length of 2D cell must be positive, not %g", field->b);
    out->kind = field->kind; out->i0 = field->i0; out->i1 = field->i1;
    out->a = field->a; out->b = field->b;
    return DREX_OK;
}

int drex_flow_field_f64(const drex_flow_field_t *field, const double *x, int64_t P, double *v,
                        double *L, void *stream) {
    drex::FlowField fl;
    if (int rc = make_flow_field(field, &fl)) return rc;
    if (P < 0) return fail(DREX_E_INVALID, "negative size");
    if (P == 0) return DREX_OK;
    if (!x) return fail(DREX_E_INVALID, "drex_flow_field_f64: null pointer");
    drex::flow_field_kernel<<<(unsigned)((P + 127) / 128), 128, 0, (cudaStream_t)stream>>>(fl, x, P, v, L);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

int drex_pathlines_f64(const drex_flow_field_t *field, const double *final_locations, int64_t P,
                       const double *min_coords, const double *max_coords, double max_strain,
                       double t_limit, double rtol, double atol, int64_t n_intervals,
                       int32_t n_nodes, int32_t timestamps_given, int32_t max_steps,
                       double *t_start, double *timestamps, double *positions, double *strains,
                       double *node_positions, double *L_nodes, double *step_times,
                       int32_t max_record, int32_t *n_taken, int32_t *status, void *stream) {
    drex::PathlineArgs A;
    if (int rc = make_flow_field(field, &A.field)) return rc;
    if (P < 0 || n_intervals < 0) return fail(DREX_E_INVALID, "negative size");
    if (P == 0) return DREX_OK;
    if (!final_locations || !min_coords || !max_coords || !status)
        return fail(DREX_E_INVALID, "drex_pathlines_f64: null pointer");
    if (!(t_limit < 0.0)) return fail(DREX_E_INVALID, "drex_pathlines_f64: t_limit must be negative");
    if (!(rtol > 0.0) || !(atol >= 0.0)) return fail(DREX_E_INVALID, "drex_pathlines_f64: bad tolerances");
    if (n_nodes == 1 || n_nodes < 0 || n_nodes > 129)
        return fail(DREX_E_INVALID, "drex_pathlines_f64: n_nodes must be 0 or in 2..129");
    if ((positions || n_intervals > 0 || timestamps_given) && !timestamps)
        return fail(DREX_E_INVALID, "drex_pathlines_f64: timestamps buffer required");
    if ((node_positions || L_nodes) && (n_nodes < 2 || !positions))
        return fail(DREX_E_INVALID, "drex_pathlines_f64: node tables need n_nodes >= 2 and positions");
    if (step_times && max_record <= 0)
        return fail(DREX_E_INVALID, "drex_pathlines_f64: max_record must be positive");
    for (int i = 0; i < 3; ++i) {
        if (!(min_coords[i] <= max_coords[i]))
            return fail(DREX_E_INVALID, "drex_pathlines_f64: min_coords must not exceed max_coords");
        A.box.lo[i] = min_coords[i];
        A.box.hi[i] = max_coords[i];
    }
    A.final_locations = final_locations; A.P = P;
    A.max_strain = max_strain; A.t_limit = t_limit; A.rtol = rtol; A.atol = atol;
    A.n_intervals = n_intervals; A.n_nodes = (node_positions || L_nodes) ? n_nodes : 0;
    A.timestamps_given = timestamps_given; A.max_steps = max_steps > 0 ? max_steps : 1000000;
    A.t_start = t_start; A.timestamps = timestamps; A.positions = positions; A.strains = strains;
    A.node_positions = node_positions; A.L_nodes = L_nodes; A.step_times = step_times;
    A.max_record = max_record; A.n_taken = n_taken; A.status = status;
    drex::pathlines_kernel<<<(unsigned)((P + 127) / 128), 128, 0, (cudaStream_t)stream>>>(A);
    DREX_CUDA_CHECK(cudaGetLastError());
    return DREX_OK;
}

// ---- M-index ---------------------------------------------------------------------------

static int max_misorientation(int a, int b) {  // stats.py:203-215
    if ((a == 2 && b == 4) || (a == 3 && b == 6)) return 120;
    if ((a == 4 && b == 8) || (a == 6 && b == 12)) return 90;
    if ((a == 1 && b == 1) || (a == 2 && b == 2)) return 180;
    return -1;
}

// stats.misorientations_random (stats.py:130-200), Grimmer (1979)
static double misorientations_random(double low, double high, int Ms, int Ns, bool *ok) {
    const double pi = 3.14159265358979323846, deg = pi / 180.0;
    const double a = std::tan(deg * 90.0 / Ms);
    const double b = 2.0 * std::atan(std::sqrt(1.0 + a * a)) / deg;
    const double c = std::round(2.0 * std::atan(std::sqrt(1.0 + 2.0 * a * a)) / deg);
    double counts[2] = {0.0, 0.0};
    const double edges[2] = {low, high};
    for (int i = 0; i < 2; ++i) {
        const double e = edges[i], d = e * deg;
        if (0 <= e && e <= 180.0 / Ms) {
            counts[i] += (Ns / 180.0) * (1 - std::cos(d));
        } else if (180.0 / Ms <= e && e <= 180.0 * Ms / Ns) {
            counts[i] += (Ns / 180.0) * a * std::sin(d);
        } else if (90 <= e && e <= b) {
            counts[i] += (Ms / 90.0) * ((Ms + a) * std::sin(d) - Ms * (1 - std::cos(d)));
        } else if (b <= e && e <= c) {
            const double nu = std::pow(std::tan(deg * e / 2.0), 2.0);
            counts[i] =
                (Ms / 90.0) *
                ((Ms + a) * std::sin(d) - Ms * (1 - std::cos(d)) +
                 (Ms / 180.0) *
                     ((1 - std::cos(d)) *
                          (std::acos((1 - nu * std::cos(deg * 180.0 / Ms)) / (nu - 1)) / deg +
                           2.0 * std::acos(a / (std::sqrt(nu - a * a) * std::sqrt(nu - 1))) / deg) -
                      2.0 * std::sin(d) *
                          (2.0 * std::acos(a / std::sqrt(nu - 1)) / deg +
                           a * std::acos(1.0 / std::sqrt(nu - a * a)) / deg)));
        } else {
            *ok = false;  // the reference hits `assert False` here (stats.py:198)
            return NAN;
        }
    }
    return (counts[0] + counts[1]) / 2.0;
}

// symmetry operations as 4x4 matrices on scalar-last quaternions (geometry.py:121-182
// applied as at stats.py:118-123 through utils.quat_product, utils.py:365-371, which has
// no cross term: q -> [w_s v + w v_s, w_s w - v_s.v])
static int build_symops(int a, int b, drex::SymOps *ops) {
    int n = 0;
    auto push_quat = [&](double x, double y, double z, double w) {
        double *m = ops->m[n++];
        memset(m, 0, 16 * sizeof(double));
        const double v[3] = {x, y, z};
        for (int i = 0; i < 3; ++i) {
            m[4 * i + i] = w;
            m[4 * i + 3] = v[i];
            m[12 + i] = -v[i];
        }
        m[15] = w;
    };
    auto push_diag = [&](double d0, double d1, double d2, double d3) {
        double *m = ops->m[n++];
        memset(m, 0, 16 * sizeof(double));
        m[0] = d0; m[5] = d1; m[10] = d2; m[15] = d3;
    };
    const double pi = 3.14159265358979323846;
    static const double AX[3][3] = {{0, 0, 1}, {0, 1, 0}, {1, 0, 0}};  // z, y, x
    auto push_rots = [&](double unit, int i0, int i1, int step) {
        for (int k = 0; k < 3; ++k)
            for (int i = i0; i <= i1; i += step) {
                const double h = 0.5 * i * unit, s = std::sin(h), c = std::cos(h);
                push_quat(AX[k][0] * s, AX[k][1] * s, AX[k][2] * s, c);
            }
    };
    push_quat(0, 0, 0, 1);
    if (a == 1 && b == 1) {
    } else if (a == 2 && (b == 2 || b == 4)) {
        push_rots(pi, 1, 1, 1);
        push_diag(1, -1, -1, 1);
        push_diag(1, -1, 1, -1);
        push_diag(1, 1, -1, -1);
    } else if (a == 3 && b == 6) {
        push_rots(pi / 3, 1, 2, 1);
    } else if (a == 4 && b == 8) {
        push_rots(pi / 2, 1, 3, 1);
    } else if (a == 6 && b == 12) {
        push_rots(pi / 3, 1, 2, 1);
        push_rots(pi / 6, 1, 5, 2);
    } else {
        return -1;
    }
    ops->n = n;
    return n;
}

int drex_mindex_theory_host(int system_a, int system_b, double *theory_host) {
    const int tmax = max_misorientation(system_a, system_b);
    if (tmax < 0) return fail(DREX_E_UNSUPPORTED, "unsupported lattice system: (%d, %d)", system_a, system_b);
    if (!theory_host) return fail(DREX_E_INVALID, "null pointer");
    bool ok = true;
    for (int i = 0; i < tmax; ++i)
        theory_host[i] = misorientations_random((double)i, (double)(i + 1), system_a, system_b, &ok);
    if (!ok)
        return fail(DREX_E_UNSUPPORTED,
                    "random-texture misorientation density undefined for lattice system (%d, %d)",
                    system_a, system_b);
    return tmax;
}

int drex_misorientations_random_host(double low, double high, int system_a, int system_b,
                                     double *density_host) {
    const int tmax = max_misorientation(system_a, system_b);
    if (tmax < 0) return fail(DREX_E_UNSUPPORTED, "unsupported lattice system: (%d, %d)", system_a, system_b);
    if (!density_host) return fail(DREX_E_INVALID, "null pointer");
    if (!(0 <= low && low <= high && high <= tmax))
        return fail(DREX_E_INVALID, "bounds must obey 0 <= low <= high <= %d", tmax);
    bool ok = true;
    *density_host = misorientations_random(low, high, system_a, system_b, &ok);
    if (!ok)
        return fail(DREX_E_UNSUPPORTED,
                    "random-texture misorientation density undefined for lattice system (%d, %d)",
                    system_a, system_b);
    return DREX_OK;
}

// Bin thresholds for mindex_bin (drex_diagnostics.cuh): thr[b], b = 1..tmax, is the largest
// float32 d with floor(2 * (double)(float)(acosf(d) * rad2deg_f)) >= b; thr[tmax + 1] the largest
// d whose angle exceeds tmax.  Uses the host libm acosf, the function numpy/numba call for the
// reference's float32 arrays (geometry.py:104-116).
static void mindex_thresholds(int tmax, float *thr) {
    const float rad2deg_f = (float)(180.0 / 3.14159265358979323846);
    auto angle = [&](float d) { return 2.0 * (double)(float)(acosf(d) * rad2deg_f); };
    thr[0] = 2.0f;  // unused
    for (int b = 1; b <= tmax + 1; ++b) {
        // condition C(d): b <= tmax ? angle(d) >= b : angle(d) > tmax ; monotone: true for small d
        auto cond = [&](float d) { return b <= tmax ? angle(d) >= (double)b : angle(d) > (double)tmax; };
        const double target = (b <= tmax ? b : tmax) * 0.5 * 3.14159265358979323846 / 180.0;
        float d = (float)std::cos(target);
        if (d < 0.0f) d = 0.0f;
        if (d > 1.0f) d = 1.0f;
        if (!cond(0.0f)) {  // unreachable even at d = 0 (angle(0) = 180)
            thr[b] = -1.0f;
            continue;
        }
        while (d > 0.0f && !cond(d)) d = std::nextafterf(d, -1.0f);
        while (d < 1.0f && cond(std::nextafterf(d, 2.0f))) d = std::nextafterf(d, 2.0f);
        thr[b] = d;
    }
}

static size_t mindex_ws_layout(int64_t S, int64_t N, int n_store, int tmax, size_t *off_counts,
                               size_t *off_theory, size_t *off_thr) {
    size_t quats = align_up((size_t)S * (size_t)N * (size_t)n_store * sizeof(float4), 256);
    *off_counts = quats;
    size_t counts = align_up((size_t)S * (size_t)tmax * sizeof(unsigned long long), 256);
    *off_theory = quats + counts;
    size_t theory = align_up((size_t)tmax * sizeof(double), 256);
    *off_thr = quats + counts + theory;
    return quats + counts + theory + align_up(sizeof(drex::MindexBins), 256);
}

// monoclinic / orthorhombic: identity + 3 rotations stored, 3 sign-flip reflections implied
static inline bool reflect7(int a, int b) { return a == 2 && (b == 2 || b == 4); }

size_t drex_mindex_workspace_bytes(int64_t S, int64_t N, int system_a, int system_b) {
    drex::SymOps ops;
    const int n_ops = build_symops(system_a, system_b, &ops);
    const int tmax = max_misorientation(system_a, system_b);
    if (n_ops < 0 || tmax < 0 || S <= 0 || N <= 0) return 256;
    size_t a, b, c;
    return mindex_ws_layout(S, N, reflect7(system_a, system_b) ? 4 : n_ops, tmax, &a, &b, &c);
}

int drex_mindex_f32(const double *o_soa, int64_t S, int64_t N, int system_a, int system_b,
                    double *M, int64_t *counts, void *workspace, size_t workspace_bytes,
                    void *stream) {
    drex::SymOps ops;
    const int n_ops = build_symops(system_a, system_b, &ops);
    const int tmax = max_misorientation(system_a, system_b);
    if (n_ops < 0 || tmax < 0)
        return fail(DREX_E_UNSUPPORTED, "unsupported lattice system: (%d, %d)", system_a, system_b);
    double theory_host[drex::MAX_THETA];
    int rc = drex_mindex_theory_host(system_a, system_b, theory_host);
    if (rc < 0) return rc;
    if (S < 0 || N < 0) return fail(DREX_E_INVALID, "negative size");
    if (S == 0) return DREX_OK;
    if (!o_soa || !M) return fail(DREX_E_INVALID, "drex_mindex_f32: null pointer");
    size_t off_counts, off_theory, off_thr;
    const bool refl = reflect7(system_a, system_b);
    const int n_store = refl ? 4 : n_ops;
    const size_t need = mindex_ws_layout(S, N, n_store, tmax, &off_counts, &off_theory, &off_thr);
    if (!workspace || workspace_bytes < need)
        return fail(DREX_E_WORKSPACE, "drex_mindex_f32: workspace too small (%zu < %zu)",
                    workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    float4 *quats = (float4 *)workspace;
    unsigned long long *cnt = (unsigned long long *)((char *)workspace + off_counts);
    double *theory = (double *)((char *)workspace + off_theory);
    drex::MindexBins *bins = (drex::MindexBins *)((char *)workspace + off_thr);
    const float *thr = bins->thr;  // (device address; first member)
    drex::MindexBins bins_host;
    memset(&bins_host, 0, sizeof(bins_host));
    mindex_thresholds(tmax, bins_host.thr);
    bins_host.thr[tmax + 2] = -1.0f;  // ends the walk up the thresholds
    for (int c = 0; c <= drex::MINDEX_LUT; ++c)
        bins_host.lut[c] = (unsigned short)drex::mindex_count((float)(c + 1) * (1.0f / (float)drex::MINDEX_LUT),
                                                              bins_host.thr, tmax);
    DREX_CUDA_CHECK(cudaMemsetAsync(cnt, 0, (size_t)S * tmax * sizeof(unsigned long long), st));
    DREX_CUDA_CHECK(cudaMemcpyAsync(theory, theory_host, tmax * sizeof(double), cudaMemcpyHostToDevice, st));
    DREX_CUDA_CHECK(cudaMemcpyAsync(bins, &bins_host, sizeof(bins_host), cudaMemcpyHostToDevice, st));
    if (N > 0) {
        dim3 gq((unsigned)S, (unsigned)((N + 127) / 128));
        drex::mindex_quats_kernel<<<gq, 128, 0, st>>>(o_soa, N, quats, n_store, ops);
        DREX_CUDA_CHECK(cudaGetLastError());
        const int n_tiles = (int)((N + drex::MINDEX_TILE - 1) / drex::MINDEX_TILE);
        if ((int64_t)n_tiles * (n_tiles + 1) / 2 > 65535)
            return fail(DREX_E_INVALID, "drex_mindex_f32: textures of more than %d grains are not supported",
                        361 * drex::MINDEX_TILE);
        dim3 gp((unsigned)S, (unsigned)((int64_t)n_tiles * (n_tiles + 1) / 2));
        if (refl) {
            // all 49 dots (cross-check) or only the products that can matter (see the kernels)
            const char *full = getenv("DREX_MINDEX_FULL");
            const unsigned long long neg_zero2 = 0x8000000080000000ull;  // (-0.0f, -0.0f)
            if (full && full[0] == '1')
                drex::mindex_pairs_reflect_kernel<<<gp, drex::MINDEX_TILE, 0, st>>>(quats, N, n_tiles, tmax, bins,
                                                                                    cnt, neg_zero2);
            else
                drex::mindex_pairs_reflect_live_kernel<<<gp, drex::MINDEX_TILE, 0, st>>>(quats, N, n_tiles, tmax,
                                                                                         bins, cnt, neg_zero2);
        } else
        switch (n_ops) {
        case 1: drex::mindex_pairs_kernel<1><<<gp, drex::MINDEX_TILE, 0, st>>>(quats, N, n_tiles, tmax, thr, cnt); break;
        case 7: drex::mindex_pairs_kernel<7><<<gp, drex::MINDEX_TILE, 0, st>>>(quats, N, n_tiles, tmax, thr, cnt); break;
        case 10: drex::mindex_pairs_kernel<10><<<gp, drex::MINDEX_TILE, 0, st>>>(quats, N, n_tiles, tmax, thr, cnt); break;
        case 16: drex::mindex_pairs_kernel<16><<<gp, drex::MINDEX_TILE, 0, st>>>(quats, N, n_tiles, tmax, thr, cnt); break;
        default: return fail(DREX_E_INVALID, "unexpected symmetry operator count %d", n_ops);
        }
        DREX_CUDA_CHECK(cudaGetLastError());
    }
    drex::mindex_finalize_kernel<<<(unsigned)S, 32, 0, st>>>(cnt, theory, tmax, M);
    DREX_CUDA_CHECK(cudaGetLastError());
    if (counts)
        DREX_CUDA_CHECK(cudaMemcpyAsync(counts, cnt, (size_t)S * tmax * sizeof(unsigned long long),
                                        cudaMemcpyDeviceToDevice, st));
    return DREX_OK;
}

// ---- FP64 peak probe -----------------------------------------------------------------------

__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters) {
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999999, c = 1e-12;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int drex_fp64_peak_probe(int device, double *flops_host) {
    if (!flops_host) return fail(DREX_E_INVALID, "null pointer");
    DREX_CUDA_CHECK(cudaSetDevice(device));
    int sms = 0;
    DREX_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms * 8, threads = 256, iters = 1 << 16;
    double *buf = nullptr;
    DREX_CUDA_CHECK(cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(buf, iters);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaFree(buf);
            return fail(DREX_E_CUDA, "fp64 probe failed: %s", cudaGetErrorString(e));
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3);
        if (rep > 0 && flops > best) best = flops;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    *flops_host = best;
    return DREX_OK;
}

// ---- compute-only ceiling of the per-grain physics ---------------------------------------------

// Each thread evaluates the olivine right-hand side `iters` times on data it keeps in
// registers (the orientation drifts slowly so that nothing can be hoisted): what the
// physics code reaches at the integrator's occupancy (512 threads per SM, 128 registers)
// when nothing else is in the way -- no staging, no accumulators, no barriers.
__global__ void __launch_bounds__(512, 1)
grain_rhs_ceiling_kernel(drex::FabricConsts fc, const double *__restrict__ flow, double *out, int iters) {
    __shared__ __align__(16) double math_tab[drex::MATH_TABLE_DOUBLES];
    drex::load_math_tables(math_tab);
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const double th = 0.37 + 1e-3 * (t % 1000), ph = 1.1 + 1e-3 * (t % 777);
    double a[9] = {cos(th), sin(th) * cos(ph), sin(th) * sin(ph), -sin(th), cos(th) * cos(ph),
                   cos(th) * sin(ph), 0.0, -sin(ph), cos(ph)};
    double D[9], L[9];  // run-time values: nothing folds away
    for (int c = 0; c < 9; ++c) {
        D[c] = flow[c];
        L[c] = flow[9 + c];
    }
    double acc = 0.0;
    for (int i = 0; i < iters; ++i) {
        drex::GrainOut o;
        drex::grain_rhs<false, drex::GRAIN_OLIVINE_DEFAULT>(a, D, L, fc, math_tab, o, nullptr, 0.5);
#pragma unroll
        for (int c = 0; c < 9; ++c) a[c] = fma(1e-4, o.da[c], a[c]);
        acc += o.energy;
    }
    out[t] = acc + a[0];
}

int drex_grain_rhs_ceiling_probe(int device, double *evals_per_s_host) {
    if (!evals_per_s_host) return fail(DREX_E_INVALID, "null pointer");
    DREX_CUDA_CHECK(cudaSetDevice(device));
    int sms = 0;
    DREX_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms, threads = 512, iters = 2000;
    drex_params_t prm = {1.5, 3.5, 5.0, 125.0, 0.3};
    const drex::FabricConsts fc = make_fabric_consts(0, prm);
    double *buf = nullptr;
    DREX_CUDA_CHECK(cudaMalloc(&buf, ((size_t)blocks * threads + 18) * sizeof(double)));
    // a general (traceless, normalised) flow: D = sym(L), max |eigenvalue| 1
    const double flow_h[18] = {0.6, 0.4, -0.2, 0.4, -0.9, 0.3, -0.2, 0.3, 0.3,
                               0.6, 0.9, -0.5, -0.1, -0.9, 0.7, 0.1, -0.1, 0.3};
    double *flow = buf + (size_t)blocks * threads;
    DREX_CUDA_CHECK(cudaMemcpy(flow, flow_h, sizeof(flow_h), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        grain_rhs_ceiling_kernel<<<blocks, threads>>>(fc, flow, buf, iters);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
            cudaFree(buf);
            return fail(DREX_E_CUDA, "grain rhs probe failed: %s", cudaGetErrorString(e));
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)iters * blocks * threads / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    *evals_per_s_host = best;
    return DREX_OK;
}

}  // extern "C"
