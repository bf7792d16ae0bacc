// drex_integrate_pipe.cuh -- the adaptive integrator as a two-slot, warp-specialised pipeline.
//
// Same method, error control and per-grain physics as drex_integrate.cuh (Bogacki-Shampine
// 3(2) with FSAL replacing the LSODA loop of Mineral.update_orientations,
// src/pydrex/minerals.py:388-541); what changes is who does what on the SM.
//
// In integrate_kernel one CTA owns one system, and about 40 % of its time goes to sections
// that cannot fill the FP64 pipe: the coupled fraction stages with their block reductions,
// the accept/reject decision, the stage flows L(t_i), barrier skew.  Here a CTA owns TWO
// systems ("slots") and no CTA-wide barrier after start-up:
//   * 15 compute warps run the data-parallel passes.  LONG commands are the grain passes
//     (initial slope, the three-stage phase 1, end-of-interval orientation copy); FRAGMENTS
//     are the short per-grain sweeps of the coupled fraction stages (core.py:431-436) and of
//     the fraction part of apply_gbs (utils.py:159-190), each ending in per-warp partial
//     sums.  A warp waits on a per-slot "command" mbarrier and arrives on a per-slot "done"
//     mbarrier; between the 32-grain chunks of a long command on one slot it polls the
//     OTHER slot and runs its pending fragment, so the reductions of one system hide behind
//     the phase 1 of the other;
//   * one control warp owns everything serial for both slots: work queue, L(t) stage flows,
//     the deformation gradient, combining the 15 partial sums in a fixed order, the error
//     norm, the step controller, posting the next command.
// A grain is always handled by the same thread (chunk q belongs to warp q mod 15, lane =
// grain mod 32) in every pass, so per-grain scratch needs no synchronisation beyond program
// order.  Sums are formed in a fixed order: results do not depend on which CTA or slot ran a
// system, nor on how the two slots interleaved.  No size limit on N.
#pragma once
#include "drex_integrate.cuh"

namespace drex {

#ifndef DREX_PIPE_CW
#define DREX_PIPE_CW 15
#endif
constexpr int PIPE_CW = DREX_PIPE_CW;  // compute warps; warp PIPE_CW is the control warp
constexpr int PIPE_BLOCK = (PIPE_CW + 1) * 32;

// per-slot global scratch, in units of Npad doubles (YK buffers: chunked SoA as in
// drex_integrate.cuh)
constexpr int PS_YKA = 0, PS_YKB = 18, PS_FA = 36, PS_FB = 37, PS_KFA = 38, PS_KFB = 39;
constexpr int PS_E2 = 40, PS_E3 = 41, PS_E4 = 42, PS_T1 = 43, PS_T2 = 44;
constexpr int PS_ARRAYS = 45;

enum { CMD_INIT = 1, CMD_STEP = 2, CMD_FINISH = 3, CMD_EXIT = 4, CMD_FRAG = 5 };
// fragments, in the order a system meets them; slot state = ST_FRAG + fragment while one is out
enum { FRAG_K0 = 0, FRAG_K1, FRAG_S1, FRAG_S2, FRAG_S3, FRAG_S4, FRAG_G1, FRAG_G2, FRAG_G3 };
constexpr int ST_FRAG = 16;
#ifndef DREX_FRAG_U
#define DREX_FRAG_U 4
#endif

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// non-blocking, warp-uniform: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity, int lane) {
    uint32_t done = 0;
    if (lane == 0)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    done = __shfl_sync(0xffffffffu, done, 0);
    __syncwarp();
    return done != 0;
}

// normalised fraction before the sliding test (utils.py:188-189, 171): one expression shared
// by the fraction fragments and the orientation pass so both agree on which grains slide
__device__ __forceinline__ double gbs_fraction(double f, double s1) { return fmax(f, 0.0) / s1; }

struct PipeSlot {
    StageFlow st[3];  // stage times t + h/2, t + 3h/4, t + h ([2] doubles as "t0" for the initial slope)
    double F[9], kF[9], Fnew[9], kFnew[9];
    double xnode[33];
    double Ltab[33 * 9];
    // command block: written by the control warp before it arrives on cmd_bar, read by the
    // compute warps after they pass it
    double h, s1, pa, pb;
    const double *yk_cur;
    double *yk_nxt;
    double *f_cur, *f_nxt, *kf_cur, *kf_nxt, *f_out;
    int cmd, frag, sys, regime, fab, f_active;
    // results of the compute warps
    double part[PIPE_CW + 1][2];  // per-warp partial sums of a fragment
    float err[PIPE_CW + 1];       // per-warp weighted error max (inf: not finite)
    // control warp's own bookkeeping
    double t, t_begin, t_end, span, h_carry, sum_clip, sum_clip_new, gbm_scale, errF, emax_o;
    int64_t ip;
    int state, n_rhs, n_acc, n_rej, status, last_rejected, const_flow, done;
};

struct PipeShared {
    PipeSlot slot[2];
    uint64_t cmd_bar[2], done_bar[2];
    uint64_t tma_bar[PIPE_CW + 1];
};

// One fragment over this thread's grains (chunks warp, warp + 15, ...; four chunks per trip so
// that the loads of a trip are in flight together).  Per-warp partial sums go to C.part /
// C.err.  Not inlined: it is called between the chunks of the long commands, whose hot loops
// must stay small.
__device__ __noinline__ void pipe_fragment(PipeSlot &C, double *slot, int64_t N, int warp, int lane,
                                           double atol_abs, double relw) {
    const int64_t n_chunks = (N + 31) / 32, Npad = n_chunks * 32;
    const int frag = C.frag;
    const double pa = C.pa, pb = C.pb, h = C.h;
    const bool f_active = C.f_active != 0;
    double *f = C.f_cur, *k1 = C.kf_cur, *fnx = C.f_nxt, *kfn = C.kf_nxt, *f_out = C.f_out;
    double *E2 = slot + PS_E2 * Npad, *E3 = slot + PS_E3 * Npad, *E4 = slot + PS_E4 * Npad;
    double *T1 = slot + PS_T1 * Npad, *T2 = slot + PS_T2 * Npad;
    double sc = 0.0, se = 0.0;
    float ef = 0.0f, rf = 0.0f;
    constexpr int U = DREX_FRAG_U;
    for (int64_t q0 = warp; q0 < n_chunks; q0 += U * PIPE_CW) {
        int64_t g[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = q0 + u * PIPE_CW;
            g[u] = q * 32 + lane;
            ok[u] = q < n_chunks && g[u] < N;
        }
        double a[U], b[U], c[U], d[U], e[U];
        switch (frag) {
        case FRAG_K0:  // sum of max(f, 0) (utils.py:188-189) and of max(f, 0) E at t0
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = ok[u] ? f[g[u]] : 0.0;
                b[u] = (ok[u] && f_active) ? E4[g[u]] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double x = fmax(a[u], 0.0);
                sc += x;
                se += x * b[u];
            }
            break;
        case FRAG_K1:  // df = vf M* (f / sum f) (mean - E) s (core.py:431-436) at t0
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = ok[u] ? f[g[u]] : 0.0;
                b[u] = ok[u] ? E4[g[u]] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) k1[g[u]] = (pa * fmax(a[u], 0.0)) * (pb - b[u]);
            break;
        case FRAG_S1:  // stage 2 state
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = ok[u] ? f[g[u]] : 0.0;
                b[u] = ok[u] ? k1[g[u]] : 0.0;
                c[u] = ok[u] ? E2[g[u]] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double x = fmax(a[u] + (h * BS_A21) * b[u], 0.0);
                sc += x;
                se += x * c[u];
            }
            break;
        case FRAG_S2:  // k2, stage 3 state
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = ok[u] ? f[g[u]] : 0.0;
                b[u] = ok[u] ? k1[g[u]] : 0.0;
                c[u] = ok[u] ? E2[g[u]] : 0.0;
                d[u] = ok[u] ? E3[g[u]] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double x2 = fmax(a[u] + (h * BS_A21) * b[u], 0.0);
                const double k2 = (pa * x2) * (pb - c[u]);
                if (ok[u]) T1[g[u]] = k2;
                const double x3 = fmax(a[u] + (h * BS_A32) * k2, 0.0);
                sc += x3;
                se += x3 * d[u];
            }
            break;
        case FRAG_S3:  // k3, new state, error combination
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = ok[u] ? f[g[u]] : 0.0;
                b[u] = ok[u] ? k1[g[u]] : 0.0;
                c[u] = ok[u] ? T1[g[u]] : 0.0;
                d[u] = ok[u] ? E3[g[u]] : 0.0;
                e[u] = ok[u] ? E4[g[u]] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double x3 = fmax(a[u] + (h * BS_A32) * c[u], 0.0);
                const double k3 = (pa * x3) * (pb - d[u]);
                const double fn = a[u] + h * (BS_B1 * b[u] + BS_B2 * c[u] + BS_B3 * k3);
                if (ok[u]) {
                    fnx[g[u]] = fn;
                    T2[g[u]] = BS_E1 * b[u] + BS_E2 * c[u] + BS_E3 * k3;
                }
                const double xn = fmax(fn, 0.0);
                sc += xn;
                se += xn * e[u];
            }
            break;
        case FRAG_S4:  // k4 (FSAL) and the error of the fractions
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = ok[u] ? fnx[g[u]] : 0.0;
                b[u] = ok[u] ? T2[g[u]] : 0.0;
                c[u] = ok[u] ? E4[g[u]] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double k4 = (pa * fmax(a[u], 0.0)) * (pb - c[u]);
                if (ok[u]) {
                    kfn[g[u]] = k4;
                    const double err = h * (b[u] + BS_E4 * k4);
                    const float r = err_ratio(err, a[u], atol_abs, relw);
                    ef = fmaxf(ef, r);
                    rf += r;
                }
            }
            break;
        case FRAG_G1:  // extract_vars, then the sliding floor (utils.py:171-176); pa = sum max(f, 0), pb = floor
#pragma unroll
            for (int u = 0; u < U; ++u) a[u] = ok[u] ? f[g[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                double x = gbs_fraction(a[u], pa);
                x = (x < pb) ? pb : x;
                x = ok[u] ? x : 0.0;
                if (ok[u]) T1[g[u]] = x;
                sc += x;
            }
            break;
        case FRAG_G2:  // renormalise (utils.py:177-179), clip again (utils.py:188)
#pragma unroll
            for (int u = 0; u < U; ++u) a[u] = ok[u] ? T1[g[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double x = fmax(a[u] / pa, 0.0);
                if (ok[u]) T2[g[u]] = x;
                sc += x;
            }
            break;
        default:  // FRAG_G3: final normalisation (utils.py:189)
#pragma unroll
            for (int u = 0; u < U; ++u) a[u] = ok[u] ? T2[g[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) f_out[g[u]] = a[u] / pa;
            break;
        }
    }
    sc = warp_sum(sc);
    se = warp_sum(se);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ef = fmaxf(ef, __shfl_xor_sync(0xffffffffu, ef, o));
        rf += __shfl_xor_sync(0xffffffffu, rf, o);
    }
    if (lane == 0) {
        C.part[warp][0] = sc;
        C.part[warp][1] = se;
        C.err[warp] = (rf < INFINITY) ? ef : INFINITY;
    }
}

__global__ void __launch_bounds__(PIPE_BLOCK, 1)
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
        for (int w = 0; w < PIPE_CW; ++w) mbar_init(&sh.tma_bar[w], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();  // the only CTA-wide barrier of the kernel
    const double relw = A.tol.atol_rel + A.tol.rtol, atol_abs = A.tol.atol_abs;
#ifdef DREX_PHASE_TIMERS
    long long idle_clk = 0, frag_clk = 0;
    const long long start_clk = clock64();
#define PIPE_CLK(acc, stmt)                    \
    {                                          \
        const long long c0_ = clock64();       \
        stmt;                                  \
        acc += clock64() - c0_;                \
    }
#else
#define PIPE_CLK(acc, stmt) stmt;
#endif

    if (warp < PIPE_CW) {
        // =============================== compute warps ===================================
        double *stage = dyn_smem + warp * STAGE_DOUBLES;
        uint64_t *mbar = &sh.tma_bar[warp];
        uint32_t mbar_parity = 0, cpar = 0;
        int dead = 0;
        auto slot_of = [&](int X) { return A.scratch + ((int64_t)blockIdx.x * 2 + X) * A.slot_doubles; };
        auto finish_command = [&](int X) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.done_bar[X]);
        };
        // run slot Y's pending command if it is a fragment (called between the chunks of a
        // long command on the other slot, and from the main loop)
        auto nested_fragment = [&](int Y) {
            if ((dead >> Y) & 1) return;
            if (!mbar_test(&sh.cmd_bar[Y], (cpar >> Y) & 1u, lane)) return;
            PipeSlot &CY = sh.slot[Y];
            if (CY.cmd != CMD_FRAG) return;  // a long command: the main loop takes it
            cpar ^= 1u << Y;
            PIPE_CLK(frag_clk, pipe_fragment(CY, slot_of(Y), N, warp, lane, atol_abs, relw))
            finish_command(Y);
        };
        int next = 0;
        while (dead != 3) {
            int X = next;
            if (((dead >> X) & 1) || !mbar_test(&sh.cmd_bar[X], (cpar >> X) & 1u, lane)) {
                X ^= 1;
                if (((dead >> X) & 1) || !mbar_test(&sh.cmd_bar[X], (cpar >> X) & 1u, lane)) {
                    PIPE_CLK(idle_clk, __nanosleep(100))
                    continue;
                }
            }
            PipeSlot &C = sh.slot[X];
            const int cmd = C.cmd;
            if (cmd == CMD_FRAG) {
                nested_fragment(X);
                continue;
            }
            cpar ^= 1u << X;
            next = X ^ 1;
            if (cmd == CMD_EXIT) {
                dead |= 1 << X;
                continue;
            }
            const int64_t s = C.sys;
            const int regime = C.regime;
            const int fab_u = __shfl_sync(0xffffffffu, C.fab, 0);
            const FabricConsts &fc = T.fab[fab_u];
            const bool f_active = C.f_active != 0;
            double *slot = slot_of(X);
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
                    nested_fragment(X ^ 1);
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
                        nested_fragment(X ^ 1);
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
                    nested_fragment(X ^ 1);
                }
            }
            // this warp's chunk writes are read back by its own TMA copies (async proxy) later
            asm volatile("fence.proxy.async;" ::: "memory");
            finish_command(X);
        }
#ifdef DREX_PHASE_TIMERS
        if (tid == 0) {
            unsigned long long *out = reinterpret_cast<unsigned long long *>(A.queue) + 8;
            atomicAdd(out + 0, (unsigned long long)idle_clk);
            atomicAdd(out + 1, (unsigned long long)(clock64() - start_clk));
            atomicAdd(out + 4, (unsigned long long)frag_clk);
        }
#endif
        return;
    }

    // ================================== control warp =====================================
    uint32_t dpar = 0;
    const double floor_v = A.gbs_threshold / (double)N;

    auto slot_base = [&](int X) { return A.scratch + ((int64_t)blockIdx.x * 2 + X) * A.slot_doubles; };
    auto post = [&](int X) {  // publish the command block of slot X
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.cmd_bar[X]);
    };
    auto post_fragment = [&](int X, int frag, double pa, double pb) {
        PipeSlot &C = sh.slot[X];
        __syncwarp();  // every lane has read what it needs from the slot
        if (lane == 0) {
            C.cmd = CMD_FRAG;
            C.frag = frag;
            C.pa = pa;
            C.pb = pb;
            C.state = ST_FRAG + frag;
        }
        post(X);
    };
    // the 15 per-warp partial sums of the fragment that just finished, added in warp order
    auto combine = [&](int X, double &sc, double &se) {
        const PipeSlot &C = sh.slot[X];
        sc = 0.0;
        se = 0.0;
        for (int w = 0; w < PIPE_CW; ++w) {
            sc += C.part[w][0];
            se += C.part[w][1];
        }
    };
    auto combine_err = [&](int X) {
        const PipeSlot &C = sh.slot[X];
        float em = 0.0f;
        for (int w = 0; w < PIPE_CW; ++w) em = fmaxf(em, C.err[w]);
        return (double)em;
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
            C.state = CMD_STEP;
        }
        post(X);
    };

    // next long command of a slot whose state is consistent at time C.t
    auto advance = [&](int X) {
        PipeSlot &C = sh.slot[X];
        const bool out_of_steps = !C.done && C.status == 0 && C.n_acc + C.n_rej >= A.tol.max_steps;
        const bool finish = C.done || C.status != 0 || out_of_steps;
        __syncwarp();
        if (finish) {
            // end of interval: extract_vars -> apply_gbs -> extract_vars (minerals.py:464-474,
            // utils.py:159-190); orientations in the long pass, fractions in fragments G1-G3
            if (lane == 0) {
                if (out_of_steps) C.status = 1;  // DREX_STATUS_MAX_STEPS
                C.s1 = C.sum_clip;
                C.cmd = CMD_FINISH;
                C.state = CMD_FINISH;
            }
            post(X);
        } else {
            issue_step(X);
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
            if (DREX_STRETCH > 1.0) h = plan_step(h, span);
            C.h = h;
            C.t = t_begin;
            C.t_begin = t_begin;
            C.t_end = t_end;
            C.span = span;
            C.h_carry = 0.0;
            C.sum_clip = 0.0;  // set by fragment K0
            C.gbm_scale = gbm_scale;
            C.ip = ip;
            C.sys = sys;
            C.regime = regime;
            C.fab = fab;
            C.f_active = f_active ? 1 : 0;
            C.yk_cur = slot + PS_YKA * Npad;
            C.yk_nxt = slot + PS_YKB * Npad;
            C.f_cur = slot + PS_FA * Npad;  // the INIT pass copies the fractions here
            C.f_nxt = slot + PS_FB * Npad;
            C.kf_cur = slot + PS_KFA * Npad;
            C.kf_nxt = slot + PS_KFB * Npad;
            C.f_out = A.f_out + s * N;
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

    // accept / reject and the step controller, once the weighted error of the whole state is known
    auto decide = [&](int X, double emax, double sum_clip_new) {
        PipeSlot &C = sh.slot[X];
        const double h = C.h;
        emax = fmax(emax, C.errF);
        // every lane computes the same values; lane 0 records them
        int n_acc = C.n_acc, n_rej = C.n_rej, status = 0, done = 0, last_rejected = C.last_rejected;
        double t_now = C.t, h_carry = C.h_carry, h_next = h;
        const double t_end = C.t_end, span = C.span;
        bool accept = false;
        if (!(emax < INFINITY)) {
            status = 3;  // DREX_STATUS_NOT_FINITE
        } else {
            accept = emax <= 1.0;
            double fac = (emax > 0.0) ? DREX_SAFETY * rcbrt(emax) : 5.0;  // err^(-1/3)
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
                    else if (DREX_STRETCH > 1.0) h_next = plan_step(h_next, remaining);
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
                    double *fo = C.f_cur;
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

    // the command of slot X has been executed by all compute warps: what comes next
    auto handle = [&](int X) {
        PipeSlot &C = sh.slot[X];
        const int state = C.state;
        const double gs = C.gbm_scale;
        double sc, se;
        switch (state) {
        case CMD_INIT:
            post_fragment(X, FRAG_K0, 0.0, 0.0);
            break;
        case ST_FRAG + FRAG_K0: {
            combine(X, sc, se);
            const bool f_active = C.f_active != 0;
            const double s2 = C.st[2].s;
            __syncwarp();
            if (lane == 0) C.sum_clip = sc;
            __syncwarp();
            if (f_active) post_fragment(X, FRAG_K1, s2 * gs / sc, se / sc);
            else advance(X);
            break;
        }
        case ST_FRAG + FRAG_K1:
            advance(X);
            break;
        case CMD_STEP: {
            const double em = combine_err(X);
            const bool f_active = C.f_active != 0;
            const double sum_clip = C.sum_clip;
            __syncwarp();
            if (lane == 0) {
                C.n_rhs += 3;
                C.emax_o = em;
            }
            __syncwarp();
            if (f_active) post_fragment(X, FRAG_S1, 0.0, 0.0);
            else decide(X, em, sum_clip);
            break;
        }
        case ST_FRAG + FRAG_S1:
            combine(X, sc, se);
            post_fragment(X, FRAG_S2, C.st[0].s * gs / sc, se / sc);
            break;
        case ST_FRAG + FRAG_S2:
            combine(X, sc, se);
            post_fragment(X, FRAG_S3, C.st[1].s * gs / sc, se / sc);
            break;
        case ST_FRAG + FRAG_S3: {
            combine(X, sc, se);
            const double s2 = C.st[2].s;
            __syncwarp();
            if (lane == 0) C.sum_clip_new = sc;
            post_fragment(X, FRAG_S4, s2 * gs / sc, se / sc);
            break;
        }
        case ST_FRAG + FRAG_S4: {
            const double ef = combine_err(X);  // inf when a fraction error was not finite
            decide(X, fmax(C.emax_o, ef), C.sum_clip_new);
            break;
        }
        case CMD_FINISH:
            post_fragment(X, FRAG_G1, C.s1, floor_v);
            break;
        case ST_FRAG + FRAG_G1:
            combine(X, sc, se);
            post_fragment(X, FRAG_G2, sc, 0.0);
            break;
        case ST_FRAG + FRAG_G2:
            combine(X, sc, se);
            post_fragment(X, FRAG_G3, sc, 0.0);
            break;
        default:  // ST_FRAG + FRAG_G3: the system is finished
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
            break;
        }
    };

    setup(0);
    setup(1);
    for (;;) {
        const int st0 = sh.slot[0].state, st1 = sh.slot[1].state;
        if (st0 == CMD_EXIT && st1 == CMD_EXIT) break;
        bool any = false;
#pragma unroll 1
        for (int X = 0; X < 2; ++X) {
            if ((X ? st1 : st0) == CMD_EXIT) continue;
            if (!mbar_test(&sh.done_bar[X], (dpar >> X) & 1u, lane)) continue;
            dpar ^= 1u << X;
            any = true;
            handle(X);
        }
        if (!any) PIPE_CLK(idle_clk, __nanosleep(40))
    }
#ifdef DREX_PHASE_TIMERS
    if (lane == 0) {
        unsigned long long *out = reinterpret_cast<unsigned long long *>(A.queue) + 8;
        atomicAdd(out + 2, (unsigned long long)idle_clk);
        atomicAdd(out + 3, (unsigned long long)(clock64() - start_clk));
    }
#endif
}

}  // namespace drex
