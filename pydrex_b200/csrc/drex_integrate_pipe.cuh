// drex_integrate_pipe.cuh -- the adaptive integrator as a two-slot, warp-specialised pipeline.
//
// Same method, error control and per-grain physics as drex_integrate.cuh (Bogacki-Shampine
// 3(2) with FSAL replacing the LSODA loop of Mineral.update_orientations,
// src/pydrex/minerals.py:388-541); what changes is who does what on the SM.
//
// In integrate_kernel one CTA owns one system, and about 40 % of its time goes to sections
// that cannot fill the FP64 pipe: the coupled fraction stages with their block reductions,
// the accept/reject decision, the stage flows L(t_i), barrier skew.  Here a CTA owns TWO
// systems ("slots"):
//   * 15 compute warps only ever run the data-parallel grain passes (initial slope, the
//     three-stage phase 1, end-of-interval copy), alternating between the slots, and never
//     meet at a CTA barrier: they wait on a per-slot "command" mbarrier and arrive on a
//     per-slot "done" mbarrier;
//   * one control warp does everything serial or coupled for BOTH slots while the compute
//     warps are busy with the other slot: work queue, L(t) stage flows, the deformation
//     gradient, the coupled fraction stages (core.py:431-436; four sweeps whose inputs
//     stream through a shared-memory ring filled by TMA bulk copies, so the one warp has
//     kilobytes in flight without holding them in registers), the error norm, the step
//     controller and the fraction part of apply_gbs (utils.py:159-190).
// Everything is in a fixed order, so results do not depend on which CTA or slot ran a system.
// No size limit: the fraction stages always go through memory (L2-resident scratch).
#pragma once
#include "drex_integrate.cuh"

namespace drex {

constexpr int PIPE_CW = INT_WARPS - 1;  // compute warps; warp PIPE_CW is the control warp
constexpr int RING_TILE = 256;          // grains per ring tile (8 per lane)
constexpr int RING_NT = 4;              // tiles in flight
constexpr int RING_MAXIN = 5;           // input arrays per sweep
constexpr int RING_DOUBLES = RING_NT * RING_MAXIN * RING_TILE;

// per-slot global scratch, in units of Npad doubles (YK buffers: chunked SoA as in
// drex_integrate.cuh)
constexpr int PS_YKA = 0, PS_YKB = 18, PS_FA = 36, PS_FB = 37, PS_KFA = 38, PS_KFB = 39;
constexpr int PS_E2 = 40, PS_E3 = 41, PS_E4 = 42, PS_T1 = 43, PS_T2 = 44;
constexpr int PS_ARRAYS = 45;

#define PIPE_FENCE()                                   \
    asm volatile("fence.proxy.async;" ::: "memory")
enum { CMD_INIT = 1, CMD_STEP = 2, CMD_FINISH = 3, CMD_EXIT = 4 };

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// normalised fraction before the sliding test (utils.py:188-189, 171): one expression shared
// by the control warp (fractions) and the compute warps (orientations) so both agree on
// which grains slide
__device__ __forceinline__ double gbs_fraction(double f, double s1) { return fmax(f, 0.0) / s1; }

struct PipeSlot {
    StageFlow st[3];  // stage times t + h/2, t + 3h/4, t + h ([2] doubles as "t0" for the initial slope)
    double F[9], kF[9], Fnew[9], kFnew[9];
    double xnode[33];
    double Ltab[33 * 9];
    // command block: written by the control warp before it arrives on cmd_bar, read by the
    // compute warps after they pass it
    double h, s1;
    const double *yk_cur;
    double *yk_nxt;
    const double *f_cur;
    int cmd, sys, regime, fab, f_active;
    float err[INT_WARPS];  // per compute warp: weighted error max of its grains (inf: not finite)
    // control warp's own bookkeeping
    double t, t_begin, t_end, span, h_carry, sum_clip, gbm_scale, errF;
    double *f_nxt, *kf_cur, *kf_nxt;
    int64_t ip;
    int state, n_rhs, n_acc, n_rej, status, last_rejected, const_flow, done;
};

struct PipeShared {
    PipeSlot slot[2];
    uint64_t cmd_bar[2], done_bar[2];
    uint64_t tma_bar[INT_WARPS];
    uint64_t ring_bar[RING_NT];
};

// One sweep of the control warp over NIN per-grain arrays of Npad doubles: tiles of
// RING_TILE grains are fetched by TMA bulk copies RING_NT tiles ahead; body(t, v) gets the
// tile in registers, v[i][j] = array i at grain t * RING_TILE + j * 32 + lane (the stage is
// released for the next copy before the arithmetic starts, and the loads cannot alias the
// body's global stores).  `seq` counts tiles over the whole launch (ring stage =
// seq % RING_NT, mbarrier parity from seq / RING_NT).  Inputs written through the generic
// proxy must have been fenced (fence.proxy.async) by their writers.
constexpr int RING_J = RING_TILE / 32;
template <int NIN, class Body>
__device__ __forceinline__ void ring_sweep(const double *const (&in)[NIN], int64_t Npad, double *ring,
                                           uint64_t *bars, uint32_t &seq, int lane, Body body) {
    static_assert(NIN <= RING_MAXIN, "ring stage too small");
    const int ntiles = (int)((Npad + RING_TILE - 1) / RING_TILE);
    auto issue = [&](int t, uint32_t n) {
        if (lane == 0) {
            const int stg = (int)(n % RING_NT);
            const int64_t g0 = (int64_t)t * RING_TILE;
            const int64_t cnt = (Npad - g0 < RING_TILE) ? (Npad - g0) : RING_TILE;
            const uint32_t bytes = (uint32_t)(cnt * sizeof(double));
            mbar_expect_tx(&bars[stg], bytes * NIN);
#pragma unroll
            for (int i = 0; i < NIN; ++i)
                bulk_g2s(ring + (stg * RING_MAXIN + i) * RING_TILE, in[i] + g0, bytes, &bars[stg]);
        }
    };
    const uint32_t base = seq;
    for (int t = 0; t < ntiles && t < RING_NT; ++t) issue(t, base + t);
    for (int t = 0; t < ntiles; ++t) {
        const uint32_t n = base + (uint32_t)t;
        const int stg = (int)(n % RING_NT);
        mbar_wait(&bars[stg], (n / RING_NT) & 1u);
        const double *tile = ring + stg * RING_MAXIN * RING_TILE + lane;
        double v[NIN][RING_J];
#pragma unroll
        for (int i = 0; i < NIN; ++i)
#pragma unroll
            for (int j = 0; j < RING_J; ++j) v[i][j] = tile[i * RING_TILE + j * 32];
        // The stage goes back to the async proxy before the arithmetic starts.  The bulk copy
        // that refills it writes through the async proxy and is not ordered after these
        // generic-proxy loads by __syncwarp alone (it was observed to overtake them under the
        // compute warps' shared-memory traffic): the proxy fence orders them.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (t + RING_NT < ntiles) issue(t + RING_NT, n + RING_NT);
        body(t, v);
    }
    seq = base + (uint32_t)ntiles;
}

__global__ void __launch_bounds__(INT_BLOCK, 1)
integrate_pipe_kernel(const __grid_constant__ IntegrateArgs A, const __grid_constant__ FabricTable T) {
    extern __shared__ __align__(128) double dyn_smem[];
    __shared__ PipeShared sh;
    __shared__ __align__(16) double math_tab[MATH_TABLE_DOUBLES];
    load_math_tables(math_tab);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
    const int64_t N = A.N;
    const int64_t n_chunks = (N + 31) / 32;
    const int64_t Npad = n_chunks * 32;
    if (tid == 0) {
        for (int x = 0; x < 2; ++x) {
            mbar_init(&sh.cmd_bar[x], 1);
            mbar_init(&sh.done_bar[x], PIPE_CW);
        }
        for (int w = 0; w < INT_WARPS; ++w) mbar_init(&sh.tma_bar[w], 1);
        for (int i = 0; i < RING_NT; ++i) mbar_init(&sh.ring_bar[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();  // the only CTA-wide barrier of the kernel
    const double relw = A.tol.atol_rel + A.tol.rtol, atol_abs = A.tol.atol_abs;
#ifdef DREX_PHASE_TIMERS
    long long idle_clk = 0;
    const long long start_clk = clock64();
#define PIPE_IDLE(stmt)                        \
    {                                          \
        const long long c0_ = clock64();       \
        stmt;                                  \
        idle_clk += clock64() - c0_;           \
    }
    long long h_clk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PIPE_TIME(slot, stmt)                  \
    {                                          \
        const long long c0_ = clock64();       \
        stmt;                                  \
        h_clk[slot] += clock64() - c0_;        \
    }
#else
#define PIPE_IDLE(stmt) stmt;
#define PIPE_TIME(slot, stmt) stmt;
#endif

    if (warp < PIPE_CW) {
        // =============================== compute warps ===================================
        double *stage = dyn_smem + warp * STAGE_DOUBLES;
        uint64_t *mbar = &sh.tma_bar[warp];
        uint32_t mbar_parity = 0, cpar = 0;
        int dead = 0;
        for (int X = 0; dead != 3; X ^= 1) {
            if ((dead >> X) & 1) continue;
            PIPE_IDLE(mbar_wait(&sh.cmd_bar[X], (cpar >> X) & 1u))
            cpar ^= 1u << X;
            PipeSlot &C = sh.slot[X];
            const int cmd = C.cmd;
            if (cmd == CMD_EXIT) {
                dead |= 1 << X;
                continue;
            }
            const int64_t s = C.sys;
            const int regime = C.regime;
            const int fab_u = __shfl_sync(0xffffffffu, C.fab, 0);
            const FabricConsts &fc = T.fab[fab_u];
            const bool f_active = C.f_active != 0;
            double *slot = A.scratch + ((int64_t)blockIdx.x * 2 + X) * A.slot_doubles;
            double *E2 = slot + PS_E2 * Npad, *E3 = slot + PS_E3 * Npad, *E4 = slot + PS_E4 * Npad;
            const double *o_in = A.o_in + s * 9 * N;

            if (cmd == CMD_INIT) {
                // ---- initial slope at t0 (one right-hand side), state into the chunked buffer
                double *yk_cur = const_cast<double *>(C.yk_cur);
                // fractions into the slot's own 16-byte aligned, zero-padded buffer (the control
                // warp streams them with TMA)
                const double *f_in = A.f_in + s * N;
                double *fA = const_cast<double *>(C.f_cur);
                double an[9], fn;
                {
                    const int64_t g0 = (int64_t)warp * 32 + lane;
                    fn = (warp < n_chunks && g0 < N) ? f_in[g0] : 0.0;
                    const bool ok = warp < n_chunks && g0 < N;
#pragma unroll
                    for (int c = 0; c < 9; ++c) an[c] = ok ? o_in[c * N + g0] : 0.0;
                }
                for (int64_t q = warp; q < n_chunks; q += PIPE_CW) {
                    const int64_t g = q * 32 + lane;
                    double a[9], k[9], e;
#pragma unroll
                    for (int c = 0; c < 9; ++c) a[c] = an[c];
                    fA[g] = fn;
                    {
                        const int64_t gn = (q + PIPE_CW) * 32 + lane;
                        const bool ok = (q + PIPE_CW) < n_chunks && gn < N;
#pragma unroll
                        for (int c = 0; c < 9; ++c) an[c] = ok ? o_in[c * N + gn] : 0.0;
                        fn = ok ? f_in[gn] : 0.0;
                    }
                    double *blk = yk_cur + q * CHUNK_DOUBLES + lane;
#pragma unroll
                    for (int c = 0; c < 9; ++c) blk[c * 32] = a[c];  // idle lanes of the tail: zeros
                    clip9(a);
                    stage_rhs<GRAIN_RUNTIME>(a, C.st[2], regime, fc, math_tab, k, e);
#pragma unroll
                    for (int c = 0; c < 9; ++c) blk[(9 + c) * 32] = k[c];
                    if (f_active && g < N) E4[g] = e;
                }
            } else if (cmd == CMD_STEP) {
                // ---- phase 1: every grain through all three stages ------------------------
                const double *yk_cur = C.yk_cur;
                double *yk_nxt = C.yk_nxt;
                const double h = C.h;
                float emax_f = 0.0f, racc = 0.0f;
                auto issue_chunk = [&](int64_t q) {
                    if (lane == 0) {
                        mbar_expect_tx(mbar, (uint32_t)(CHUNK_DOUBLES * sizeof(double)));
                        bulk_g2s(stage, yk_cur + q * CHUNK_DOUBLES, (uint32_t)(CHUNK_DOUBLES * sizeof(double)), mbar);
                    }
                };
                if (warp < n_chunks) issue_chunk(warp);
                auto phase1 = [&](auto mode_tag) {
                    constexpr int MODE = decltype(mode_tag)::value;
                    for (int64_t q = warp; q < n_chunks; q += PIPE_CW) {
                        const int64_t g = q * 32 + lane;
                        const int64_t qn = q + PIPE_CW;
                        const bool live = g < N;
                        double *sa = stage + lane, *sk = stage + 9 * 32 + lane;
                        double *sy = stage + 18 * 32 + lane, *se = stage + 27 * 32 + lane;
                        double k[9], st[9], e = 0.0;
                        mbar_wait(mbar, mbar_parity);
                        mbar_parity ^= 1u;
#pragma unroll
                        for (int c = 0; c < 9; ++c) st[c] = fma(h * BS_A21, sk[c * 32], sa[c * 32]);
                        // rolled stage loop: one copy of the right-hand side in the hot loop
                        // (instruction cache), bookkeeping behind uniform branches
#pragma unroll 1
                        for (int stg = 0; stg < 3; ++stg) {
                            clip9(st);
                            stage_rhs<MODE>(st, C.st[stg], regime, fc, math_tab, k, e);
                            if (f_active && live) {
                                double *Ep = (stg == 0) ? E2 : ((stg == 1) ? E3 : E4);
                                Ep[g] = e;
                            }
                            if (stg == 0) {
#pragma unroll
                                for (int c = 0; c < 9; ++c) {
                                    const double ac = sa[c * 32], kc = sk[c * 32];
                                    sy[c * 32] = fma(h * BS_B2, k[c], fma(h * BS_B1, kc, ac));
                                    se[c * 32] = fma(BS_E2, k[c], BS_E1 * kc);
                                    st[c] = fma(h * BS_A32, k[c], ac);
                                }
                                __syncwarp();  // rows 0-17 are free: fetch the next chunk
                                if (qn < n_chunks) issue_chunk(qn);
                            } else if (stg == 1) {
#pragma unroll
                                for (int c = 0; c < 9; ++c) {
                                    const double yn = fma(h * BS_B3, k[c], sy[c * 32]);
                                    sy[c * 32] = yn;
                                    se[c * 32] = fma(BS_E3, k[c], se[c * 32]);
                                    st[c] = yn;
                                }
                            }
                        }
                        double *blk = yk_nxt + q * CHUNK_DOUBLES + lane;
#pragma unroll
                        for (int c = 0; c < 9; ++c) {
                            const double yn = sy[c * 32];
                            const double err = h * (se[c * 32] + BS_E4 * k[c]);
                            const float r = err_ratio(err, yn, atol_abs, relw);
                            emax_f = fmaxf(emax_f, r);
                            racc += r;  // non-finite k or y_new makes r non-finite
                            blk[c * 32] = yn;
                            blk[(9 + c) * 32] = k[c];
                        }
                    }
                };
                if (regime != 4 && regime != 6) phase1(std::integral_constant<int, STAGE_OTHER>{});
                else if (fc.phase != 0) phase1(std::integral_constant<int, GRAIN_ENSTATITE>{});
                else if (fc.default_exponents) phase1(std::integral_constant<int, GRAIN_OLIVINE_DEFAULT>{});
                else phase1(std::integral_constant<int, GRAIN_OLIVINE_GENERIC>{});
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    emax_f = fmaxf(emax_f, __shfl_xor_sync(0xffffffffu, emax_f, o));
                    racc += __shfl_xor_sync(0xffffffffu, racc, o);
                }
                if (lane == 0) C.err[warp] = (racc < INFINITY) ? emax_f : INFINITY;
            } else {
                // ---- end of interval, orientations: sliding grains get the interval-start
                // orientation back (utils.py:159-180), everything clipped (utils.py:187)
                const double *yk_cur = C.yk_cur, *f_cur = C.f_cur;
                const double s1 = C.s1, floor_v = A.gbs_threshold / (double)N;
                double *o_out = A.o_out + s * 9 * N;
                auto load_f = [&](int64_t q) {
                    const int64_t g = q * 32 + lane;
                    return (q < n_chunks && g < N) ? f_cur[g] : 1.0;
                };
                auto load_o = [&](int64_t q, double fv, double (&dst)[9]) {
                    const int64_t g = q * 32 + lane;
                    const bool ok = q < n_chunks && g < N;
                    const bool sl = gbs_fraction(fv, s1) < floor_v;
                    const double *blk = yk_cur + q * CHUNK_DOUBLES + lane;
#pragma unroll
                    for (int c = 0; c < 9; ++c) dst[c] = ok ? (sl ? o_in[c * N + g] : blk[c * 32]) : 0.0;
                };
                double an[9];
                double f1 = load_f((int64_t)warp + PIPE_CW);
                load_o(warp, load_f(warp), an);
                for (int64_t q = warp; q < n_chunks; q += PIPE_CW) {
                    const int64_t g = q * 32 + lane;
                    double a[9];
#pragma unroll
                    for (int c = 0; c < 9; ++c) a[c] = an[c];
                    const double f2 = load_f(q + 2 * PIPE_CW);
                    load_o(q + PIPE_CW, f1, an);
                    f1 = f2;
                    clip9(a);
                    if (g < N) {
#pragma unroll
                        for (int c = 0; c < 9; ++c) o_out[c * N + g] = a[c];
                    }
                }
            }
            // this warp's global writes are read by TMA (async proxy) later
            PIPE_FENCE();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.done_bar[X]);
        }
#ifdef DREX_PHASE_TIMERS
        if (tid == 0) {
            unsigned long long *out = reinterpret_cast<unsigned long long *>(A.queue) + 8;
            atomicAdd(out + 0, (unsigned long long)idle_clk);
            atomicAdd(out + 1, (unsigned long long)(clock64() - start_clk));
        }
#endif
        return;
    }

    // ================================== control warp =====================================
    double *ring = dyn_smem + PIPE_CW * STAGE_DOUBLES;
    uint32_t ring_seq = 0, dpar = 0;
    const double floor_v = A.gbs_threshold / (double)N;

    auto slot_base = [&](int X) { return A.scratch + ((int64_t)blockIdx.x * 2 + X) * A.slot_doubles; };
    auto post = [&](int X) {  // publish the command block of slot X
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.cmd_bar[X]);
    };

    // stage flows and deformation-gradient stages for the step [t, t + h], then CMD_STEP
    auto issue_step = [&](int X) {
        PipeSlot &C = sh.slot[X];
        const double t = C.t, h = C.h;
        const int regime = C.regime;
        const double *Ltable = (A.K <= 33) ? C.Ltab : A.L_nodes + C.ip * A.K * 9;
        if (!C.const_flow)
            warp_stage_flows(A, Ltable, C.st, 0, 3, t + BS_A21 * h, t + BS_A32 * h, t + h, C.t_begin,
                             C.t_end, C.xnode, lane);
        __syncwarp();
        if (lane == 0) {
            // dF/dt = L(t) F (minerals.py:426), same BS3(2) stages
            double F2[9], F3[9], k2[9], k3[9];
            for (int c = 0; c < 9; ++c) F2[c] = C.F[c] + h * BS_A21 * C.kF[c];
            matmul3(C.st[0].Lraw, F2, k2);
            for (int c = 0; c < 9; ++c) F3[c] = C.F[c] + h * BS_A32 * k2[c];
            matmul3(C.st[1].Lraw, F3, k3);
            for (int c = 0; c < 9; ++c)
                C.Fnew[c] = C.F[c] + h * (BS_B1 * C.kF[c] + BS_B2 * k2[c] + BS_B3 * k3[c]);
            matmul3(C.st[2].Lraw, C.Fnew, C.kFnew);
            double em = 0.0;
            for (int c = 0; c < 9; ++c) {
                const double err = h * (BS_E1 * C.kF[c] + BS_E2 * k2[c] + BS_E3 * k3[c] + BS_E4 * C.kFnew[c]);
                const double sc = atol_abs + relw * fabs(C.Fnew[c]);
                const double r = fabs(err) / sc;
                em = (r != r) ? INFINITY : fmax(em, r);
            }
            C.errF = em;
            if (regime == 1) {  // stretch of (L F) at the stage states (minerals.py:427-429)
                left_stretch(k2, C.st[0].V);
                left_stretch(k3, C.st[1].V);
                left_stretch(C.kFnew, C.st[2].V);
            }
            C.cmd = CMD_STEP;
        }
        post(X);
    };

    // CMD_FINISH for the orientations, then the fraction part of
    // extract_vars -> apply_gbs -> extract_vars (minerals.py:464-474, utils.py:159-190)
    auto issue_finish = [&](int X) {
        PipeSlot &C = sh.slot[X];
        const double s1 = C.sum_clip;
        if (lane == 0) {
            C.s1 = s1;
            C.cmd = CMD_FINISH;
        }
        post(X);
        double *slot = slot_base(X);
        double *T1 = slot + PS_T1 * Npad, *T2 = slot + PS_T2 * Npad;
        double *f_out = A.f_out + (int64_t)C.sys * N;
        double s2 = 0.0, s3 = 0.0;
        {
            const double *const in[1] = {C.f_cur};
            ring_sweep<1>(in, Npad, ring, sh.ring_bar, ring_seq, lane, [&](int t, const double (&v)[1][RING_J]) {
#pragma unroll
                for (int j = 0; j < RING_J; ++j) {
                    const int64_t g = (int64_t)t * RING_TILE + j * 32 + lane;
                    double x = gbs_fraction(v[0][j], s1);
                    x = (x < floor_v) ? floor_v : x;
                    x = (g < N) ? x : 0.0;
                    if (g < Npad) T1[g] = x;
                    s2 += x;
                }
            });
        }
        s2 = warp_sum(s2);
        PIPE_FENCE();
        __syncwarp();
        {
            const double *const in[1] = {T1};
            ring_sweep<1>(in, Npad, ring, sh.ring_bar, ring_seq, lane, [&](int t, const double (&v)[1][RING_J]) {
#pragma unroll
                for (int j = 0; j < RING_J; ++j) {
                    const int64_t g = (int64_t)t * RING_TILE + j * 32 + lane;
                    const double x = (g < Npad) ? fmax(v[0][j] / s2, 0.0) : 0.0;
                    if (g < Npad) T2[g] = x;
                    s3 += x;
                }
            });
        }
        s3 = warp_sum(s3);
        PIPE_FENCE();
        __syncwarp();
        {
            const double *const in[1] = {T2};
            ring_sweep<1>(in, Npad, ring, sh.ring_bar, ring_seq, lane, [&](int t, const double (&v)[1][RING_J]) {
#pragma unroll
                for (int j = 0; j < RING_J; ++j) {
                    const int64_t g = (int64_t)t * RING_TILE + j * 32 + lane;
                    if (g < N) f_out[g] = v[0][j] / s3;
                }
            });
        }
        if (lane == 0) C.state = CMD_FINISH;
        __syncwarp();
    };

    // next command of a slot whose state is consistent at time C.t
    auto advance = [&](int X) {
        PipeSlot &C = sh.slot[X];
        if (!C.done && C.status == 0 && C.n_acc + C.n_rej >= A.tol.max_steps) {
            if (lane == 0) C.status = 1;  // DREX_STATUS_MAX_STEPS
            __syncwarp();
        }
        if (C.done || C.status != 0) {
            PIPE_TIME(3, issue_finish(X))
        } else {
            if (lane == 0) C.state = CMD_STEP;
            PIPE_TIME(4, issue_step(X))
        }
    };

    // pull the next system from the queue into slot X and start its initial slope
    auto setup = [&](int X) {
        PipeSlot &C = sh.slot[X];
        int sys = 0;
        if (lane == 0) {
            const int ticket = atomicAdd(A.queue, 1);
            sys = (ticket < A.S && A.order != nullptr) ? A.order[ticket] : ticket;
        }
        sys = __shfl_sync(0xffffffffu, sys, 0);
        if (sys >= A.S) {
            if (lane == 0) {
                C.cmd = CMD_EXIT;
                C.state = CMD_EXIT;
            }
            post(X);
            return;
        }
        const int64_t s = sys;
        const int64_t ip = A.sys_particle[s];
        const int regime = A.sys_regime[s];
        const int fab = A.sys_fabric[s];
        const FabricConsts &fc = T.fab[fab];
        const double t_begin = A.t0[ip], t_end = A.t1[ip];
        const double span = t_end - t_begin;
        const double gbm_scale = ((regime == 6) ? 0.3 : 1.0) * A.sys_vol_frac[s] * A.gbm_mobility;
        const bool f_active = (regime == 4 || regime == 6) && fc.phase == 0 && gbm_scale != 0.0;
        double *slot = slot_base(X);
        const double *Ltable = (A.K <= 33) ? C.Ltab : A.L_nodes + ip * A.K * 9;
        if (A.L_kind == 2)
            for (int j = lane; j < A.K && j < 33; j += 32)
                C.xnode[j] = 0.5 * (1.0 - cospi((double)j / (double)(A.K - 1)));
        if (A.K <= 33)
            for (int j = lane; j < A.K * 9; j += 32) C.Ltab[j] = A.L_nodes[ip * A.K * 9 + j];
        __syncwarp();
        warp_stage_flows(A, Ltable, C.st, 2, 1, t_begin, t_begin, t_begin, t_begin, t_end, C.xnode, lane);
        double *fA = slot + PS_FA * Npad;  // the INIT pass copies the fractions here
        if (lane == 0) {
            const double *L = C.st[2].Lraw;
            C.const_flow = (A.L_kind == 0 || A.K == 1) ? 1 : 0;
            if (C.const_flow) {
                C.st[0] = C.st[2];
                C.st[1] = C.st[2];
            }
            for (int c = 0; c < 9; ++c) C.F[c] = A.F[s * 9 + c];
            matmul3(L, C.F, C.kF);
            if (regime == 1) left_stretch(C.kF, C.st[2].V);
            double h = (A.h_io != nullptr) ? A.h_io[s] : 0.0;
            if (!(fabs(h) > 0.0)) h = (A.tol.first_step_frac > 0.0 ? A.tol.first_step_frac : 1.0) * span;
            h = copysign(fabs(h), span);
            if (A.tol.max_step_frac > 0.0 && fabs(h) > A.tol.max_step_frac * fabs(span))
                h = A.tol.max_step_frac * span;
            // a step carried over from a longer interval must not leave this one
            if (!(fabs(h) < fabs(span))) h = span;
            C.h = h;
            C.t = t_begin;
            C.t_begin = t_begin;
            C.t_end = t_end;
            C.span = span;
            C.h_carry = 0.0;
            C.sum_clip = 0.0;  // set after the INIT pass
            C.gbm_scale = gbm_scale;
            C.ip = ip;
            C.sys = sys;
            C.regime = regime;
            C.fab = fab;
            C.f_active = f_active ? 1 : 0;
            C.yk_cur = slot + PS_YKA * Npad;
            C.yk_nxt = slot + PS_YKB * Npad;
            C.f_cur = fA;
            C.f_nxt = slot + PS_FB * Npad;
            C.kf_cur = slot + PS_KFA * Npad;
            C.kf_nxt = slot + PS_KFB * Npad;
            C.n_rhs = 1;
            C.n_acc = 0;
            C.n_rej = 0;
            C.status = 0;
            C.last_rejected = 0;
            C.done = (span == 0.0) ? 1 : 0;
            C.cmd = CMD_INIT;
            C.state = CMD_INIT;
        }
        post(X);
    };

    // sum of max(f, 0) (utils.py:188-189) and the initial slope of the fractions once the
    // compute warps have copied the fractions and evaluated the energies at t0
    auto after_init = [&](int X) {
        PipeSlot &C = sh.slot[X];
        double *slot = slot_base(X);
        const double *E4 = slot + PS_E4 * Npad;
        double sumc = 0.0, se = 0.0;
        if (C.f_active) {
            const double *const in[2] = {C.f_cur, E4};
            ring_sweep<2>(in, Npad, ring, sh.ring_bar, ring_seq, lane, [&](int t, const double (&v)[2][RING_J]) {
#pragma unroll
                for (int j = 0; j < RING_J; ++j) {
                    const int64_t g = (int64_t)t * RING_TILE + j * 32 + lane;
                    const double x = (g < N) ? fmax(v[0][j], 0.0) : 0.0;
                    sumc += x;
                    se += (g < N) ? x * v[1][j] : 0.0;
                }
            });
            sumc = warp_sum(sumc);
            se = warp_sum(se);
            // df = vf M* (f / sum f) (mean - E) s (core.py:431-436)
            const double mean = se / sumc, ps = C.st[2].s * C.gbm_scale / sumc;
            double *kf = C.kf_cur;
            ring_sweep<2>(in, Npad, ring, sh.ring_bar, ring_seq, lane, [&](int t, const double (&v)[2][RING_J]) {
#pragma unroll
                for (int j = 0; j < RING_J; ++j) {
                    const int64_t g = (int64_t)t * RING_TILE + j * 32 + lane;
                    const double k1 = (g < N) ? (ps * fmax(v[0][j], 0.0)) * (mean - v[1][j]) : 0.0;
                    if (g < Npad) kf[g] = k1;
                }
            });
            PIPE_FENCE();
        } else {
            const double *const in[1] = {C.f_cur};
            ring_sweep<1>(in, Npad, ring, sh.ring_bar, ring_seq, lane, [&](int t, const double (&v)[1][RING_J]) {
#pragma unroll
                for (int j = 0; j < RING_J; ++j) {
                    const int64_t g = (int64_t)t * RING_TILE + j * 32 + lane;
                    sumc += (g < N) ? fmax(v[0][j], 0.0) : 0.0;
                }
            });
            sumc = warp_sum(sumc);
        }
        if (lane == 0) C.sum_clip = sumc;
        __syncwarp();
        advance(X);
    };

    // coupled fraction stages, error norm, accept / reject, step controller
    auto after_step = [&](int X) {
        PipeSlot &C = sh.slot[X];
        const double h = C.h;
        float em = 0.0f;
        for (int w = 0; w < PIPE_CW; ++w) em = fmaxf(em, C.err[w]);
        double emax = (double)em;
        double sum_clip_new = C.sum_clip;
        if (C.f_active) {
            double *slot = slot_base(X);
            const double *E2 = slot + PS_E2 * Npad, *E3 = slot + PS_E3 * Npad, *E4 = slot + PS_E4 * Npad;
            double *T1 = slot + PS_T1 * Npad, *T2 = slot + PS_T2 * Npad;
            const double *fc_ = C.f_cur, *k1p = C.kf_cur;
            double *fnx = C.f_nxt, *kfn = C.kf_nxt;
            const double gs = C.gbm_scale;
            double sc = 0.0, se = 0.0;
            {  // stage 2 state
                const double *const in[3] = {fc_, k1p, E2};
                ring_sweep<3>(in, Npad, ring, sh.ring_bar, ring_seq, lane, [&](int t, const double (&v)[3][RING_J]) {
#pragma unroll
                    for (int j = 0; j < RING_J; ++j) {
                        const int64_t g = (int64_t)t * RING_TILE + j * 32 + lane;
                        const bool ok = g < N;
                        const double x = ok ? fmax(v[0][j] + (h * BS_A21) * v[1][j], 0.0) : 0.0;
                        sc += x;
                        se += ok ? x * v[2][j] : 0.0;
                    }
                });
            }
            sc = warp_sum(sc);
            se = warp_sum(se);
            const double mean2 = se / sc, ps2 = C.st[0].s * gs / sc;
            sc = 0.0;
            se = 0.0;
            {  // k2, stage 3 state
                const double *const in[4] = {fc_, k1p, E2, E3};
                ring_sweep<4>(in, Npad, ring, sh.ring_bar, ring_seq, lane, [&](int t, const double (&v)[4][RING_J]) {
#pragma unroll
                    for (int j = 0; j < RING_J; ++j) {
                        const int64_t g = (int64_t)t * RING_TILE + j * 32 + lane;
                        const bool ok = g < N;
                        const double f = v[0][j];
                        const double x2 = fmax(f + (h * BS_A21) * v[1][j], 0.0);
                        const double k2 = ok ? (ps2 * x2) * (mean2 - v[2][j]) : 0.0;
                        if (g < Npad) T1[g] = k2;
                        const double x3 = ok ? fmax(f + (h * BS_A32) * k2, 0.0) : 0.0;
                        sc += x3;
                        se += ok ? x3 * v[3][j] : 0.0;
                    }
                });
            }
            sc = warp_sum(sc);
            se = warp_sum(se);
            const double mean3 = se / sc, ps3 = C.st[1].s * gs / sc;
            sc = 0.0;
            se = 0.0;
            PIPE_FENCE();
            __syncwarp();
            {  // k3, new state
                const double *const in[5] = {fc_, k1p, T1, E3, E4};
                ring_sweep<5>(in, Npad, ring, sh.ring_bar, ring_seq, lane, [&](int t, const double (&v)[5][RING_J]) {
#pragma unroll
                    for (int j = 0; j < RING_J; ++j) {
                        const int64_t g = (int64_t)t * RING_TILE + j * 32 + lane;
                        const bool ok = g < N;
                        const double f = v[0][j], k1 = v[1][j], k2 = v[2][j];
                        const double x3 = fmax(f + (h * BS_A32) * k2, 0.0);
                        const double k3 = ok ? (ps3 * x3) * (mean3 - v[3][j]) : 0.0;
                        const double fn = ok ? f + h * (BS_B1 * k1 + BS_B2 * k2 + BS_B3 * k3) : 0.0;
                        if (g < Npad) {
                            fnx[g] = fn;
                            T2[g] = ok ? BS_E1 * k1 + BS_E2 * k2 + BS_E3 * k3 : 0.0;  // error combination
                        }
                        const double xn = fmax(fn, 0.0);
                        sc += xn;
                        se += ok ? xn * v[4][j] : 0.0;
                    }
                });
            }
            sc = warp_sum(sc);
            se = warp_sum(se);
            sum_clip_new = sc;
            const double mean4 = se / sc, ps4 = C.st[2].s * gs / sc;
            PIPE_FENCE();
            __syncwarp();
            float ef = 0.0f, rf = 0.0f;
            {  // k4 (FSAL), error
                const double *const in[3] = {fnx, T2, E4};
                ring_sweep<3>(in, Npad, ring, sh.ring_bar, ring_seq, lane, [&](int t, const double (&v)[3][RING_J]) {
#pragma unroll
                    for (int j = 0; j < RING_J; ++j) {
                        const int64_t g = (int64_t)t * RING_TILE + j * 32 + lane;
                        const bool ok = g < N;
                        const double fn = v[0][j];
                        const double k4 = ok ? (ps4 * fmax(fn, 0.0)) * (mean4 - v[2][j]) : 0.0;
                        if (g < Npad) kfn[g] = k4;
                        if (ok) {
                            const double err = h * (v[1][j] + BS_E4 * k4);
                            const float r = err_ratio(err, fn, atol_abs, relw);
                            ef = fmaxf(ef, r);
                            rf += r;
                        }
                    }
                });
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ef = fmaxf(ef, __shfl_xor_sync(0xffffffffu, ef, o));
                rf += __shfl_xor_sync(0xffffffffu, rf, o);
            }
            emax = (rf < INFINITY) ? fmax(emax, (double)ef) : INFINITY;
            PIPE_FENCE();
            __syncwarp();
        }
        emax = fmax(emax, C.errF);

        // ---- accept / reject (every lane computes the same values; lane 0 records them) ----
        int n_acc = C.n_acc, n_rej = C.n_rej, status = 0, done = 0, last_rejected = C.last_rejected;
        double t_now = C.t, h_carry = C.h_carry, h_next = h;
        const double t_end = C.t_end, span = C.span;
        bool accept = false;
        if (!(emax < INFINITY)) {
            status = 3;  // DREX_STATUS_NOT_FINITE
        } else {
            accept = emax <= 1.0;
            double fac = (emax > 0.0) ? 0.9 * rcbrt(emax) : 5.0;  // err^(-1/3)
            fac = fmin(accept ? (last_rejected ? 1.0 : 5.0) : 1.0, fmax(0.2, fac));
            h_next = h * fac;
            if (accept) {
                ++n_acc;
                last_rejected = 0;
                t_now = t_now + h;
                h_carry = h_next;
                const double remaining = t_end - t_now;
                if (fabs(remaining) <= 1e-13 * fmax(fabs(t_end), fabs(span))) {
                    done = 1;
                } else {
                    if (A.tol.max_step_frac > 0.0 && fabs(h_next) > A.tol.max_step_frac * fabs(span))
                        h_next = A.tol.max_step_frac * span;
                    if (fabs(h_next) >= fabs(remaining) * (1.0 - 1e-10)) h_next = remaining;
                }
            } else {
                ++n_rej;
                last_rejected = 1;
                h_carry = h_next;
                if (fabs(h_next) <= 1e-14 * fmax(fabs(t_now), fabs(span))) status = 2;  // DREX_STATUS_STEP_UNDERFLOW
            }
        }
        __syncwarp();  // all lanes have read the slot
        if (lane == 0) {
            C.n_rhs += 3;
            C.n_acc = n_acc;
            C.n_rej = n_rej;
            C.status = status;
            C.last_rejected = last_rejected;
            C.done = done;
            C.h_carry = h_carry;
            if (status == 0) {
                C.t = t_now;
                C.h = done ? h_carry : h_next;  // when done: suggestion for the next interval
            }
            if (accept) {
                C.sum_clip = sum_clip_new;
                const double *ykt = C.yk_cur;
                C.yk_cur = C.yk_nxt;
                C.yk_nxt = const_cast<double *>(ykt);
                if (C.f_active) {
                    double *fo = const_cast<double *>(C.f_cur);
                    C.f_cur = C.f_nxt;
                    C.f_nxt = fo;
                    double *kft = C.kf_cur;
                    C.kf_cur = C.kf_nxt;
                    C.kf_nxt = kft;
                }
                for (int c = 0; c < 9; ++c) {
                    C.F[c] = C.Fnew[c];
                    C.kF[c] = C.kFnew[c];
                }
            }
        }
        __syncwarp();
        advance(X);
    };

    auto after_finish = [&](int X) {
        PipeSlot &C = sh.slot[X];
        if (lane == 0) {
            const int64_t s = C.sys;
            for (int c = 0; c < 9; ++c) A.F[s * 9 + c] = C.F[c];
            if (A.h_io) A.h_io[s] = C.h;
            A.status[s] = C.status;
            if (A.n_rhs) A.n_rhs[s] = C.n_rhs;
            if (A.n_accept) A.n_accept[s] = C.n_acc;
            if (A.n_reject) A.n_reject[s] = C.n_rej;
        }
        __syncwarp();
        setup(X);
    };

    setup(0);
    setup(1);
    for (int X = 0;; X ^= 1) {
        const int st0 = sh.slot[0].state, st1 = sh.slot[1].state;
        if (st0 == CMD_EXIT && st1 == CMD_EXIT) break;
        const int state = X ? st1 : st0;
        if (state == CMD_EXIT) continue;
        PIPE_IDLE(mbar_wait(&sh.done_bar[X], (dpar >> X) & 1u))
        dpar ^= 1u << X;
        PIPE_FENCE();
        if (state == CMD_INIT) PIPE_TIME(0, after_init(X))
        else if (state == CMD_STEP) PIPE_TIME(1, after_step(X))
        else PIPE_TIME(2, after_finish(X))
    }
#ifdef DREX_PHASE_TIMERS
    if (lane == 0) {
        unsigned long long *out = reinterpret_cast<unsigned long long *>(A.queue) + 8;
        atomicAdd(out + 2, (unsigned long long)idle_clk);
        atomicAdd(out + 3, (unsigned long long)(clock64() - start_clk));
        for (int i = 0; i < 5; ++i) atomicAdd(out + 4 + i, (unsigned long long)h_clk[i]);
    }
#endif
}

}  // namespace drex
