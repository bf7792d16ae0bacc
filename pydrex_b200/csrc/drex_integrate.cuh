// drex_integrate.cuh -- batched on-device adaptive integrator for grain ensembles (sm_100a).
//
// Replaces the scipy LSODA loop of Mineral.update_orientations
// (src/pydrex/minerals.py:388-541): state y = [F(9) | orientations(9N) | fractions(N)],
// error control with the reference's weights (minerals.py:523-524) in the weighted max
// norm ODEPACK lsoda uses, grain-boundary sliding once at the end of the interval
// (minerals.py:464-474).
//
// Design (B200-first, not a translation of LSODA; the tests carry a numpy model of it):
//  * THE ODE DECOUPLES.  The rotation rate and strain energy E_g of a grain depend on that
//    grain and L(t) only (core.py:687-760).  The volume fractions obey the replicator equation
//        df_g/dt = s vf M* fhat_g (Ebar - E_g),  fhat = clip(f)/sum clip(f),  Ebar = sum fhat E
//    (core.py:431-436, minerals.py:450-451), whose exact solution is
//        fhat_g(t) = fhat_g(t0) e^{v_g(t)} / sum_j fhat_j(t0) e^{v_j(t)},  dv_g/dt = -s vf M* E_g.
//    So a grain carries TEN unknowns (nine direction cosines and v_g) that see no other grain
//    until the end of the interval, where one normalisation recovers the fractions.  No
//    reduction, no barrier and no memory round trip is left inside a step.
//  * A WARP OWNS A CHUNK OF 32 GRAINS FROM t0 TO t1: the state lives in the warp's shared-memory
//    rows and registers for the whole interval and every chunk walks its OWN Bogacki-Shampine
//    3(2) steps (FSAL), controlled by the weighted max norm over its grains.  Per outer interval
//    a grain costs 80 B of HBM reads and 80 B of writes -- whatever the number of steps.  (The
//    first version of this kernel stepped all N grains of a system together: four block
//    reductions and a 144 B/grain round trip per step, 41 % of its time outside the arithmetic.)
//  * persistent CTAs (one per SM, 16 warps) pull SYSTEMS from an atomic queue (longest first);
//    the 16 warps pull the system's chunks from a shared-memory ticket.  Up to three systems are
//    in flight per CTA: one warp prepares the next system (flow tables, the deformation gradient)
//    while the others still integrate chunks of the current one, and the warp that completes
//    the last chunk of a system finishes it (normalisation, grain-boundary sliding) while the
//    others have moved on -- the warps never meet at a barrier.
//  * the next chunk's orientations and fractions (ten rows of 256 B) are fetched by TMA bulk
//    copies (cp.async.bulk -> mbarrier) into the warp's prefetch rows while it integrates the
//    current chunk.
//  * L(t) over the interval is a polynomial (DREX_L_CHEBYSHEV): the preparing warp converts the
//    node table to Chebyshev coefficients once, and the three stage matrices of a step cost a
//    warp one Clenshaw recurrence on 30 lanes.
//  * FP64 throughout; clip inside every right-hand side as utils.extract_vars does
//    (utils.py:183-190).
#pragma once
#include "drex_common.cuh"
#include "drex_derivatives.cuh"
#include "drex_grain.cuh"

namespace drex {

// A warp's unit of work is a CHUNK of 32 * INT_G grains, INT_G per lane, with 512 / INT_G threads
// per CTA.  INT_G = 1 is the product.  INT_G = 2 (-DDREX_INT_G=2: half as many warps with 255
// registers and two independent dependency chains each; step control, stage flows and chunk
// bookkeeping paid once per 64 grains) is kept as a measured experiment: the right-hand side
// alone runs at 95% of the one-grain rate that way, but the whole kernel takes 5.1 ms instead of
// 3.7 ms per launch on the bench workload -- the sections around the stages are latency chains,
// whose throughput is warps / latency, so halving both their count and the warps gains nothing,
// and a 64-grain chunk takes the step its worst grain needs (5% more right-hand sides).
#ifndef DREX_INT_G
#define DREX_INT_G 1
#endif
constexpr int INT_G = DREX_INT_G;
constexpr int INT_CHUNK = 32 * INT_G;
#ifndef DREX_INT_BLOCK
#define DREX_INT_BLOCK (512 / DREX_INT_G)
#endif
constexpr int INT_BLOCK = DREX_INT_BLOCK;
constexpr int INT_WARPS = INT_BLOCK / 32;
constexpr int INT_CTAS_PER_SM = 1;
constexpr int INT_NCTX = 3;   // systems in flight per CTA
constexpr int INT_KMAX = 33;  // L(t) tables up to this many nodes live in shared memory
// per-warp shared-memory rows of 32 grains.  Per sub-chunk j (the j-th grain of every lane) a
// row set of 36: 0-8 state, 9-17 FSAL slope, 18-26 new state, 27-35 error combination (the state
// and new-state rows swap roles on an accepted step); then, per sub-chunk, the ten rows of the
// prefetched next chunk (orientation rows 0-8 of o_in, row 9 of f_in; TMA target).  Keeping the
// accumulators here instead of in registers is what keeps the stage loop out of local memory.
constexpr int ROW_A = 0, ROW_K1 = 9, ROW_Y = 18, ROW_E = 27, ROW_SET = 36;
constexpr int ROW_PF = ROW_SET * INT_G, ROW_PF_SET = 10;
constexpr int WARP_ROWS = (ROW_SET + ROW_PF_SET) * INT_G;
constexpr int WARP_DOUBLES = WARP_ROWS * 32;

struct TolArgs {
    double rtol, atol_abs, atol_rel, atol_abs_F, atol_abs_f, first_step_frac, max_step_frac;
    int max_steps, method;
};

struct IntegrateArgs {
    int64_t S, P, N;
    const int32_t *sys_particle, *sys_phase, *sys_fabric, *sys_regime;
    const double *sys_vol_frac;
    double *F;
    const double *o_in, *f_in;  // o_in never aliases o_out inside the kernel (drex_capi.cu)
    const double *o_prev;       // orientations restored by grain-boundary sliding (usually o_in)
    double *o_out, *f_out;
    const double *t0, *t1;
    int L_kind, K;
    const double *L_nodes;
    double gbm_mobility, gbs_threshold;
    TolArgs tol;
    double *h_io;           // [S][n_chunks] first trial step in / next suggestion out (as strain), or null
    const int32_t *order;   // optional queue permutation (longest systems first)
    int32_t *status, *n_rhs, *n_accept, *n_reject;
    int64_t *grain_evals;   // optional: right-hand sides x live grains per system
    double *scratch;        // per CTA and slot: chunk_cap doubles (per-chunk partial sums)
    int64_t chunk_cap;
    uint32_t *slide_mask;   // [S][n_chunks]: lanes of each chunk whose grain slides (gbs_restore_kernel)
    int *queue;             // atomic work counter, zeroed before launch
};

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a PTX) ------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
}
// Spins before a wait is declared stuck (~1 s at 64 ns per spin).  A profiler that instruments the
// kernel instruction by instruction (ncu's source counters) slows it down a hundredfold; build with
// a larger value for such captures.
#ifndef DREX_WATCHDOG_LOG2
#define DREX_WATCHDOG_LOG2 24
#endif
__device__ __noinline__ void watchdog_fire(int where, long long a, long long b, long long c);
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int where = 100) {
    uint32_t done = 0;
    const uint32_t addr = smem_addr(bar);
    unsigned tries = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && ++tries > (1u << (DREX_WATCHDOG_LOG2 - 2))) watchdog_fire(where, parity, 0, 0);
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_global, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(dst_smem)),
        "l"(src_global), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}

// A wait between warps of one CTA (a system being prepared or finished by another warp).  The
// watchdog turns a protocol bug into a launch failure instead of a hung GPU; when a debug buffer
// in mapped host memory is attached (DREX_DEBUG_HOSTPTR, development only) the waiting warp first
// records where it was stuck: the context is lost after the trap, host memory is not.
__device__ unsigned long long *g_drex_debug = nullptr;
__device__ __noinline__ void watchdog_fire(int where, long long a, long long b, long long c) {
    unsigned long long *d = g_drex_debug;
    if (d != nullptr && (threadIdx.x & 31) == 0) {
        const unsigned long long idx = (unsigned long long)blockIdx.x * 16 + (threadIdx.x >> 5);
        if (idx < 4096) {  // one record per warp, plain stores (no atomics across PCIe)
            volatile unsigned long long *r = d + 8 + 8 * idx;
            r[0] = (unsigned long long)where;
            r[1] = blockIdx.x;
            r[2] = threadIdx.x >> 5;
            r[3] = (unsigned long long)a;
            r[4] = (unsigned long long)b;
            r[5] = (unsigned long long)c;
        }
        __threadfence_system();
    }
    __syncwarp();
    __nanosleep(2000000);
    __trap();
}
__device__ __forceinline__ void spin_pause(unsigned &n, int where = 0, long long a = 0, long long b = 0,
                                           long long c = 0) {
    __nanosleep(64);
    if (++n > (1u << DREX_WATCHDOG_LOG2)) watchdog_fire(where, a, b, c);
}

// Step planning: the rest of the interval is cut into n EQUAL steps, n the smallest count
// whose step exceeds the controller's proposal by at most DREX_STRETCH.  A proposal that
// would leave a sliver (remaining = 2.05 h) becomes two steps of 1.025 h instead of three
// attempts.  The error test of every step is unchanged; the stretch only spends part of the
// controller's safety margin (DREX_SAFETY^3 = 0.73 of the tolerance; error ~ h^3, and
// 1.10^3 = 1.33).
#ifndef DREX_STRETCH
#define DREX_STRETCH 1.10
#endif
#ifndef DREX_SAFETY
#define DREX_SAFETY 0.9
#endif
__device__ __forceinline__ double plan_step(double h_prop, double remaining) {
    // the count n only needs float32 (it is a small integer unless the proposal is degenerate);
    // the step itself is remaining / n in double, the last step of an interval is set to what
    // remains, exactly
    const float ratio = __fdividef(fabsf((float)remaining), fabsf((float)h_prop) * (float)DREX_STRETCH);
    if (!(ratio > 1.0f)) return remaining;  // one step reaches the end
    if (!(ratio < 1e9f)) return h_prop;     // degenerate proposal: leave it to the step-underflow test
    return div_fast(remaining, (double)ceilf(ratio));
}

// Bogacki-Shampine 3(2)
#define BS_A21 0.5
#define BS_A32 0.75
#define BS_B1 (2.0 / 9.0)
#define BS_B2 (1.0 / 3.0)
#define BS_B3 (4.0 / 9.0)
#define BS_E1 (-5.0 / 72.0)
#define BS_E2 (1.0 / 12.0)
#define BS_E3 (1.0 / 9.0)
#define BS_E4 (-1.0 / 8.0)

// Flow quantities of one stage time, as the right-hand side consumes them.
struct ChunkFlow {
    double2 DL[9];  // {sym(L)/s, L/s} per component (minerals.py:421,438-439)
    double s;       // strain_rate_max (minerals.py:422)
    double nks;     // -(0.3 if frictional) vf M* s: dv/dt = nks E
};

// Stage flows shared between the chunks of a system.  The step planner cuts what is left of the
// interval into equal steps, so most chunks of a system walk the SAME (t, h) pairs -- and the
// flows at t + h/2, t + 3h/4, t + h are the same for all of them.  The first warp that needs a
// pair computes the flows into an entry, the others read them: bit-identical to computing them
// (same function, same arguments), without the Clenshaw recurrences (19 % of the warps' time on
// the corner-flow workload when every warp evaluated its own).  `key` is claimed with one
// compare-and-swap (0 = free), `ready` is set after the data; an entry is never evicted.
constexpr int FLOW_CACHE = 4;  // 6 and 8 measure the same; ncu's source-level instrumentation needs
                               // ~5.6 KB of the CTA's shared memory for itself, which 6 entries did not leave
struct FlowEntry {
    unsigned long long key;
    double t, h;
    ChunkFlow f[3];
};

// One system in flight (shared memory).  Written by the preparing warp before `ticket` is
// published and read-only afterwards, except the counters and the flow cache.
struct SysSlot {
    double t0, t1, kappa, rscale;
    double s_begin, s_end;         // strain-rate scale at t0 and t1 (step sizes are carried as strain)
    double inv_span2;              // 2 / (t1 - t0), 0 for an empty interval
    double Q[9];                   // regimes 0/1/7: orientation increment common to all grains
    ChunkFlow cflow;               // the flow when L is constant over the interval
    ChunkFlow flow_t0;             // otherwise: the flow at t0 (first right-hand side of every chunk)
    FlowEntry fcache[FLOW_CACHE];
    int fc_ready[FLOW_CACHE];
    double Ltab[INT_KMAX * 9];     // this particle's L(t) node table (K <= INT_KMAX)
    double coef[10][INT_KMAX];     // Chebyshev coefficients of L (0-8) and of the strain-rate scale (9)
    double ctab[2 * INT_KMAX];     // cos(pi m / (K - 1))
    double Lst[4][9];              // stage matrices of the deformation-gradient integration
    unsigned long long ticket;     // (sequence number << 32) | next chunk: the slot is ready when the
                                   // upper half equals the sequence number a warp looks for
    unsigned long long count_a;    // right-hand sides (low half) and rejected steps, summed over chunks
    unsigned long long count_b;    // accepted steps (low half) and chunks completed
    int sys, regime, fab, mode, n_chunks, prep_at, const_flow, s_table, f_active, status0, K;
    int status, rhs_last;          // worst chunk status; right-hand sides of the (ragged) last chunk
    int seq;                       // sequence number of the system in this CTA
    int busy;                      // from the start of prepare to the end of finalize
};

struct CtaShared {
    SysSlot slot[INT_NCTX];
    uint64_t mbar[INT_WARPS];   // next-chunk prefetch
    // warp-uniform state of the chunk loop.  It lives here because it is needed once per chunk
    // and would otherwise sit in registers across the whole integration of a chunk, i.e. in
    // local memory: per-thread copies whose reloads miss the (small) L1 each time
    int ws_nxt_seq[INT_WARPS], ws_nxt_q[INT_WARPS], ws_pend_slot[INT_WARPS], ws_pend_acc[INT_WARPS];
    int ws_cur_seq[INT_WARPS], ws_cur_q[INT_WARPS];
    int hand_seq;  // sequence number of the system currently handing out chunks
    int end_seq;   // number of systems this CTA processes (INT_MAX until the queue runs dry)
};

__device__ __forceinline__ int ld_vol(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void st_vol(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }
__device__ __forceinline__ unsigned long long ld_vol64(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// strain_rate_max (drex_common.cuh) with the divisions and the square root replaced by the
// ~1 ulp fast versions of drex_math.cuh.
__device__ inline double strain_rate_max_fast(const double *L) {
    const double d00 = L[0], d11 = L[4], d22 = L[8];
    const double d01 = 0.5 * (L[1] + L[3]), d02 = 0.5 * (L[2] + L[6]), d12 = 0.5 * (L[5] + L[7]);
    const double p1 = d01 * d01 + d02 * d02 + d12 * d12;
    const double q = (d00 + d11 + d22) * (1.0 / 3.0);
    const double b00 = d00 - q, b11 = d11 - q, b22 = d22 - q;
    const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
    const double p = sqrt_fast(p2 * (1.0 / 6.0));
    if (!(p > 0.0)) return fabs(q);  // multiple of the identity (or NaN, which propagates)
    const double ip = div_fast(1.0, p);
    const double c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip;
    const double c01 = d01 * ip, c02 = d02 * ip, c12 = d12 * ip;
    double r = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) +
                      c02 * (c01 * c12 - c11 * c02));
    r = fmin(1.0, fmax(-1.0, r));
    const double cphi = cos(acos(r) * (1.0 / 3.0));  // phi in [0, pi/3]
    const double sphi = sqrt_fast(fmax(0.0, 1.0 - cphi * cphi));
    const double e_hi = q + 2.0 * p * cphi;
    const double e_lo = q - p * (cphi + 1.7320508075688772935 * sphi);
    return fmax(fabs(e_hi), fabs(e_lo));
}

// sum_k a_k T_k(xi) by Clenshaw's recurrence (a_0 and a_{K-1} already carry their halves)
__device__ __forceinline__ double clenshaw(const double *a, int K, double xi) {
    double b1 = 0.0, b2 = 0.0;
    const double x2 = xi + xi;
    for (int k = K - 1; k >= 1; --k) {
        const double b0 = fma(x2, b1, a[k] - b2);
        b2 = b1;
        b1 = b0;
    }
    return fma(xi, b1, a[0] - b2);
}

// One component c of L(t) for the providers without a coefficient table: linear, piecewise
// linear, and Chebyshev tables too large for shared memory (barycentric form).
__device__ __noinline__ double eval_L_generic(const IntegrateArgs &A, const double *table, double t,
                                              double ta, double tb, int c) {
    const double *nodes = table + c;
    if (A.L_kind == 0 || A.K == 1) return nodes[0];
    const double span = tb - ta;
    const double x = (span != 0.0) ? (t - ta) / span : 0.0;
    if (A.L_kind == 1) return nodes[0] + x * (nodes[9] - nodes[0]);
    if (A.L_kind == 3) {
        const double pos = fmin(fmax(x, 0.0), 1.0) * (A.K - 1);
        int j = (int)pos;
        if (j > A.K - 2) j = A.K - 2;
        return nodes[j * 9] + (pos - j) * (nodes[(j + 1) * 9] - nodes[j * 9]);
    }
    // Chebyshev-Lobatto nodes x_j = (1 - cos(pi j/(K-1)))/2 on [0, 1]; barycentric weights
    // (-1)^j, halved at both ends
    double numer = 0.0, denom = 0.0;
    const int K = A.K;
    for (int j = 0; j < K; ++j) {
        const double xj = 0.5 * (1.0 - cospi((double)j / (double)(K - 1)));
        const double d = x - xj;
        if (fabs(d) < 1e-15) return nodes[j * 9];
        double w = (j & 1) ? -1.0 : 1.0;
        if (j == 0 || j == K - 1) w *= 0.5;
        w /= d;
        denom += w;
        numer += w * nodes[j * 9];
    }
    return numer / denom;
}

// Component c (0-8) of L(t) for the system in `sl`.
__device__ __forceinline__ double slot_L(const IntegrateArgs &A, const SysSlot &sl, double t, int c) {
    if (A.L_kind == 0 || A.K == 1) return sl.Ltab[c];
    if (A.L_kind == 2 && A.K <= INT_KMAX) {
        return clenshaw(sl.coef[c], A.K, fma(t - sl.t0, sl.inv_span2, -1.0));
    }
    const double *table = (A.K <= INT_KMAX) ? sl.Ltab : A.L_nodes + (int64_t)A.sys_particle[sl.sys] * A.K * 9;
    return eval_L_generic(A, table, t, sl.t0, sl.t1, c);
}

// The flow quantities of n <= 3 times from per-lane values: lane 10 i + c holds component c of
// L(t_i) (c < 9) and, when have_s, the strain-rate scale in c = 9.  Every lane of the warp calls.
__device__ __forceinline__ void store_flows(ChunkFlow *dst, int n, double val, bool have_s,
                                            double kappa, int lane) {
    const int si = lane / 10, c = lane - 10 * si;
    const bool active = si < n;
    const int base = active ? 10 * si : 0;
    if (!have_s) {  // warp-uniform
        double Lm[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) Lm[j] = __shfl_sync(0xffffffffu, val, base + j);
        if (active && c == 9) val = strain_rate_max_fast(Lm);
    }
    const double sv = __shfl_sync(0xffffffffu, val, base + 9);
    const int ct = (c < 9) ? 3 * (c % 3) + c / 3 : 0;  // transposed component
    const double lt = __shfl_sync(0xffffffffu, val, base + ct);
    if (active) {
        if (c < 9) {
            dst[si].DL[c] = make_double2(div_fast(0.5 * (val + lt), sv), div_fast(val, sv));
        } else {
            dst[si].s = sv;
            dst[si].nks = -kappa * sv;
        }
    }
    __syncwarp();
}

// Stage flows at n <= 3 times for a warp stepping a chunk of the system in `sl`: the hot case, a
// Chebyshev table with a validated scale series (sl.s_table).  One Clenshaw recurrence per lane,
// components and scale alike, one shared reciprocal; touches shared memory only and is small
// enough to sit inside the step loop (as a call its instructions were fetched from L2 every time).
__device__ __forceinline__ void warp_flows_cheb(const SysSlot &sl, ChunkFlow *dst, int n, double ta,
                                                double tb, double tc, int lane) {
    const int si = lane / 10, c = lane - 10 * si;
    const bool active = si < n;
    const double t = (si == 0) ? ta : ((si == 1) ? tb : tc);
    const double xi = fma(t - sl.t0, sl.inv_span2, -1.0);
    const double val = active ? clenshaw(sl.coef[c], sl.K, xi) : 0.0;
    const int base = active ? 10 * si : 0;
    const double sv = __shfl_sync(0xffffffffu, val, base + 9);
    const int ct = (c < 9) ? 3 * (c % 3) + c / 3 : 0;  // transposed component
    const double lt = __shfl_sync(0xffffffffu, val, base + ct);
    double r;  // 1 / s to ~1 ulp
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(sv));
    r = fma(r, fma(-sv, r, 1.0), r);
    r = fma(r, fma(-sv, r, 1.0), r);
    if (active) {
        if (c < 9) {
            const double sym = 0.5 * (val + lt);
            double q1 = sym * r, q2 = val * r;
            q1 = fma(fma(-sv, q1, sym), r, q1);
            q2 = fma(fma(-sv, q2, val), r, q2);
            dst[si].DL[c] = make_double2(q1, q2);
        } else {
            dst[si].s = sv;
            dst[si].nks = -sl.kappa * sv;
        }
    }
    __syncwarp();
}

// The same for every other provider (linear, piecewise linear, large tables, a scale with
// kinks): components one by one, closed-form scale per stage time.
__device__ __noinline__ void warp_flows(const IntegrateArgs &A, const SysSlot &sl, ChunkFlow *dst,
                                        int n, double ta, double tb, double tc, int lane) {
    const int si = lane / 10, c = lane - 10 * si;
    const double t = (si == 0) ? ta : ((si == 1) ? tb : tc);
    double val = 0.0;
    if (si < n && c < 9) val = slot_L(A, sl, t, c);
    store_flows(dst, n, val, false, sl.kappa, lane);
}

// The flows of the three stage times of a step (t + h/2, t + 3h/4, t + h): from the system's cache
// when another chunk has taken the same step before, else computed -- into a free cache entry when
// the step is one the planner produced (`cacheable`: a retry after a rejection has a step size of
// its own and would only use up an entry), into the warp's private buffer otherwise.
__device__ __forceinline__ const ChunkFlow *stage_flows(const IntegrateArgs &A, SysSlot &sl, ChunkFlow *wf,
                                                        double t, double h, bool cacheable, bool cheb,
                                                        int lane) {
    const double ta = fma(BS_A21, h, t), tb = fma(BS_A32, h, t), tc = t + h;
    const unsigned long long key =
        (((unsigned long long)__double_as_longlong(t) * 0x9E3779B97F4A7C15ull) ^
         (unsigned long long)__double_as_longlong(h)) | 1ull;  // never 0 (= free); t and h are compared too
    const bool in = lane < FLOW_CACHE;
    const unsigned long long tag = in ? ld_vol64(&sl.fcache[lane].key) : 0ull;
    const unsigned match = __ballot_sync(0xffffffffu, in && tag == key);
    ChunkFlow *dst = wf;
    int claimed = -1;
    if (match) {
        const int e = __ffs(match) - 1;
        if (ld_vol(&sl.fc_ready[e]) == 2) {
            __threadfence_block();  // the entry's data was written before `ready`
            if (sl.fcache[e].t == t && sl.fcache[e].h == h) return sl.fcache[e].f;
        }
        // (still being filled by another warp, or a different step with the same key: private copy)
    } else if (cacheable) {
        const unsigned empty = __ballot_sync(0xffffffffu, in && tag == 0ull);
        if (empty) {
            const int e = __ffs(empty) - 1;
            unsigned long long old = 1ull;
            if (lane == 0) old = atomicCAS(&sl.fcache[e].key, 0ull, key);
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == 0ull) {
                claimed = e;
                dst = sl.fcache[e].f;
            }
        }
    }
    if (cheb) warp_flows_cheb(sl, dst, 3, ta, tb, tc, lane);
    else warp_flows(A, sl, dst, 3, ta, tb, tc, lane);
    if (claimed >= 0) {
        if (lane == 0) {
            sl.fcache[claimed].t = t;
            sl.fcache[claimed].h = h;
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) st_vol(&sl.fc_ready[claimed], 2);
    }
    return dst;
}

// Clip to [-1, 1] (utils.py:187) on the integer pipe: the FP64 pipe is the bottleneck and
// the clip almost never fires.  |x| >= 1 (including inf/NaN) becomes +-1 exactly; a NaN
// state is still caught by the error test through the unclipped new state.
__device__ __forceinline__ double clip_int(double x) {
    const uint32_t hi = (uint32_t)__double2hiint(x);
    const bool big = (hi & 0x7fffffffu) >= 0x3ff00000u;
    const int nhi = big ? (int)((hi & 0x80000000u) | 0x3ff00000u) : (int)hi;
    const int nlo = big ? 0 : __double2loint(x);
    return __hiloint2double(nhi, nlo);
}
// All nine components at once; every lane of the warp must call it.  The test is a running
// integer maximum and one vote, the per-component clamp runs only if some lane needs it.
__device__ __forceinline__ void clip9(double (&a)[9]) {
    uint32_t m = 0;
#pragma unroll
    for (int c = 0; c < 9; ++c) m = max(m, (uint32_t)__double2hiint(a[c]) & 0x7fffffffu);
    if (__any_sync(0xffffffffu, m >= 0x3ff00000u)) {
#pragma unroll
        for (int c = 0; c < 9; ++c) a[c] = clip_int(a[c]);
    }
}

// Rotation rate (times strain_rate_max, minerals.py:450; times 0.3 when yielding, core.py:459)
// and strain energy of one grain at one stage.
template <int MODE>
__device__ __forceinline__ void chunk_rhs(const double (&a)[9], const ChunkFlow &fl, double rscale,
                                          const FabricConsts &fc, const double *__restrict__ mt,
                                          double (&k)[9], double &energy) {
    double D[9], L[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        const double2 dl = fl.DL[c];
        D[c] = dl.x;
        L[c] = dl.y;
    }
    GrainOut out;
    grain_rhs<false, MODE>(a, D, L, fc, mt, out, nullptr, 0.5 * rscale * fl.s);
#pragma unroll
    for (int c = 0; c < 9; ++c) k[c] = out.da[c];
    energy = out.energy;
}

// Weighted error ratio |err| / (atol + relw |y_new|) as float bits.  Only its size relative to
// 1 steers the step, so the reciprocal is the 2^-23-accurate hardware approximation; the
// running maximum is taken on the BITS (non-negative floats order like unsigned integers and a
// NaN sorts above infinity, so a non-finite state reaches the accept test).
__device__ __forceinline__ uint32_t err_bits(double err, double ynew, double atol_abs, double relw) {
    const double sc = fma(relw, fabs(ynew), atol_abs);
    double rc;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(sc));
    return __float_as_uint((float)(fabs(err) * rc)) & 0x7fffffffu;
}

// Development aid (scripts/phase_timers.py, build with -DDREX_PHASE_TIMERS): every warp's clock
// at the boundaries of its activities, summed over the launch into 8 counters behind the work
// queue: 0 chunk tickets (incl. waiting for a system to be prepared), 1 preparing systems,
// 2 waiting for the prefetched rows, 3 integrating chunks (stages), 4 chunk results, 5 finishing
// systems and restoring sliding grains, 6 stage flows, 7 error norm + step controller.
#ifdef DREX_PHASE_TIMERS
#define DREX_MARK(slot)                          \
    {                                            \
        const long long now_ = clock64();        \
        ptimer[slot] += now_ - pmark;            \
        pmark = now_;                            \
    }
#define DREX_TIMER_ARGS , ptimer, pmark
#else
#define DREX_MARK(slot)
#define DREX_TIMER_ARGS
#endif

struct ChunkResult {
    double v[INT_G], h_next;
    int status, n_rhs, n_acc, n_rej, row;
};

// One chunk of 32 * INT_G grains from t0 to t1: state in rows ROW_A.. of each row set of `rows`
// (already loaded), INT_G grains per lane.  Returns with the final orientations in rows res.row..
// and v_g in res.v.  Everything per grain is written as a loop over j < INT_G in straight-line
// code, so that the two dependency chains of a lane interleave.
template <int MODE>
__device__ __forceinline__ void integrate_chunk(const IntegrateArgs &A, SysSlot &sl,
                                                const FabricConsts &fc, const double *__restrict__ mt,
                                                double *rows, ChunkFlow *wf, int lane,
                                                const double (&f0)[INT_G], double h, ChunkResult &res
#ifdef DREX_PHASE_TIMERS
                                                , long long (&ptimer)[8], long long &pmark
#endif
                                                ) {
    const double t_begin = sl.t0, t_end = sl.t1, span = t_end - t_begin;
    const double relw = A.tol.atol_rel + A.tol.rtol, atol_abs = A.tol.atol_abs;
    const bool f_active = sl.f_active != 0, const_flow = sl.const_flow != 0, cheb_flow = sl.s_table != 0;
    // flows of stages 0..2 (f3: the step's) and of stage -1 (fm1: the interval's start)
    const ChunkFlow *f3 = &sl.cflow, *fm1 = const_flow ? &sl.cflow : &sl.flow_t0;
    const int fstride = const_flow ? 0 : 1;
    const double rscale = sl.rscale;
    // row offsets inside a row set; the state and new-state rows swap on an accepted step
    int off_a = ROW_A * 32 + lane, off_y = ROW_Y * 32 + lane;
    const int off_k = ROW_K1 * 32 + lane, off_e = ROW_E * 32 + lane;
    double t = t_begin, h_carry = h;
    double v[INT_G], kv1[INT_G];
#pragma unroll
    for (int j = 0; j < INT_G; ++j) v[j] = kv1[j] = 0.0;
    // attempts left and rejections: right-hand sides and accepted steps follow from them.  (With
    // three counters ptxas kept them on the stack and every step waited for their reloads.  The
    // rejections as a counter in shared memory measured the same, but ncu's source-level
    // instrumentation could not run that kernel.)
    int left = A.tol.max_steps, status = 0;
    int n_rej = 0;
    bool need_k1 = true, last_rejected = false;
    while (span != 0.0) {
        if (left <= 0) {
            status = 1;  // DREX_STATUS_MAX_STEPS
            break;
        }
        if (!const_flow) f3 = stage_flows(A, sl, wf, t, h, !last_rejected, cheb_flow, lane);
        DREX_MARK(6)
        double k[INT_G][9], st[INT_G][9], e[INT_G], kv[INT_G], vy[INT_G], ve[INT_G];
#pragma unroll
        for (int j = 0; j < INT_G; ++j) {
            const double *r = rows + j * ROW_SET * 32;
            kv[j] = vy[j] = ve[j] = 0.0;
            if (need_k1) {
#pragma unroll
                for (int c = 0; c < 9; ++c) st[j][c] = r[off_a + c * 32];
            } else {
#pragma unroll
                for (int c = 0; c < 9; ++c) st[j][c] = fma(h * BS_A21, r[off_k + c * 32], r[off_a + c * 32]);
            }
        }
        // The stages run as a ROLLED loop: one copy of the right-hand side in the hot loop keeps
        // it inside the instruction cache.  The bookkeeping between the stages differs per stage
        // and sits behind warp-uniform branches: rows y / e hold the partial new state
        // y + h (b1 k1 + b2 k2) and the partial error e1 k1 + e2 k2 after stage 2, the full ones
        // after stage 3.  Stage -1 (first step of the interval only) is the slope at t0.
#pragma unroll 1
        for (int stg = need_k1 ? -1 : 0; stg < 3; ++stg) {
            const ChunkFlow &sf = (stg < 0) ? *fm1 : f3[stg * fstride];
#pragma unroll
            for (int j = 0; j < INT_G; ++j) clip9(st[j]);
#pragma unroll
            for (int j = 0; j < INT_G; ++j) chunk_rhs<MODE>(st[j], sf, rscale, fc, mt, k[j], e[j]);
#pragma unroll
            for (int j = 0; j < INT_G; ++j) {
                double *r = rows + j * ROW_SET * 32;
                kv[j] = f_active ? sf.nks * e[j] : 0.0;
                if (stg < 0) {
#pragma unroll
                    for (int c = 0; c < 9; ++c) {
                        r[off_k + c * 32] = k[j][c];
                        st[j][c] = fma(h * BS_A21, k[j][c], r[off_a + c * 32]);
                    }
                    kv1[j] = kv[j];
                } else if (stg == 0) {
#pragma unroll
                    for (int c = 0; c < 9; ++c) {
                        const double ac = r[off_a + c * 32], kc = r[off_k + c * 32];
                        r[off_y + c * 32] = fma(h * BS_B2, k[j][c], fma(h * BS_B1, kc, ac));
                        r[off_e + c * 32] = fma(BS_E2, k[j][c], BS_E1 * kc);
                        st[j][c] = fma(h * BS_A32, k[j][c], ac);  // stage 3 starts from y + 3h/4 k2
                    }
                    vy[j] = fma(h * BS_B2, kv[j], fma(h * BS_B1, kv1[j], v[j]));
                    ve[j] = fma(BS_E2, kv[j], BS_E1 * kv1[j]);
                } else if (stg == 1) {
#pragma unroll
                    for (int c = 0; c < 9; ++c) {
                        const double yn = fma(h * BS_B3, k[j][c], r[off_y + c * 32]);
                        r[off_y + c * 32] = yn;
                        r[off_e + c * 32] = fma(BS_E3, k[j][c], r[off_e + c * 32]);
                        st[j][c] = yn;  // stage 4 is evaluated at the new state itself
                    }
                    vy[j] = fma(h * BS_B3, kv[j], vy[j]);
                    ve[j] = fma(BS_E3, kv[j], ve[j]);
                }
            }
        }
        DREX_MARK(3)
        uint32_t eb = 0;
#pragma unroll
        for (int j = 0; j < INT_G; ++j) {
            const double *r = rows + j * ROW_SET * 32;
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                const double err = h * fma(BS_E4, k[j][c], r[off_e + c * 32]);
                eb = max(eb, err_bits(err, r[off_y + c * 32], atol_abs, relw));
            }
            if (f_active) {
                // the fraction components: f_g ~ f_g(t0) e^{v_g}, err_f = f_g err_v
                const double fe = f0[j] * exp_tab(vy[j], mt);
                eb = max(eb, err_bits(fe * (h * fma(BS_E4, kv[j], ve[j])), fe, A.tol.atol_abs_f, relw));
            }
        }
        eb = __reduce_max_sync(0xffffffffu, eb);
        --left;
        need_k1 = false;
        const float emax = __uint_as_float(eb);
        if (!(emax < INFINITY)) {
            status = 3;  // DREX_STATUS_NOT_FINITE
            ++n_rej;
            break;
        }
        const bool accept = emax <= 1.0f;
        // safety * err^(-1/3) in float32: only the size of the step depends on it
        float facf = (emax > 0.0f) ? (float)DREX_SAFETY * exp2f(-0.33333334f * __log2f(emax)) : 5.0f;
        facf = fminf(accept ? (last_rejected ? 1.0f : 5.0f) : 1.0f, fmaxf(0.2f, facf));
        double h_next = h * (double)facf;
        h_carry = h_next;
        if (accept) {
            last_rejected = false;
            t += h;
            const int tmp = off_a;
            off_a = off_y;
            off_y = tmp;
#pragma unroll
            for (int j = 0; j < INT_G; ++j) {
                double *r = rows + j * ROW_SET * 32;
                v[j] = vy[j];
                kv1[j] = kv[j];
#pragma unroll
                for (int c = 0; c < 9; ++c) r[off_k + c * 32] = k[j][c];
            }
            const double remaining = t_end - t;
            if (fabs(remaining) <= 1e-13 * fmax(fabs(t_end), fabs(span))) break;
            if (A.tol.max_step_frac > 0.0 && fabs(h_next) > A.tol.max_step_frac * fabs(span))
                h_next = A.tol.max_step_frac * span;
            if (fabs(h_next) >= fabs(remaining) * (1.0 - 1e-10)) h_next = remaining;
            else if (DREX_STRETCH > 1.0) h_next = plan_step(h_next, remaining);
        } else {
            ++n_rej;
            last_rejected = true;
            if (fabs(h_next) <= 1e-14 * fmax(fabs(t), fabs(span))) {
                status = 2;  // DREX_STATUS_STEP_UNDERFLOW
                break;
            }
        }
        h = h_next;
        DREX_MARK(7)
    }
    DREX_MARK(7)
#pragma unroll
    for (int j = 0; j < INT_G; ++j) res.v[j] = v[j];
    res.h_next = h_carry;
    res.status = status;
    const int attempts = A.tol.max_steps - left;
    res.n_rhs = attempts ? 3 * attempts + 1 : 0;  // the first attempt evaluates the slope at t0 as well
    res.n_acc = attempts - n_rej;
    res.n_rej = n_rej;
    res.row = (off_a == ROW_A * 32 + lane) ? ROW_A : ROW_Y;
}

// ---- per-system prologue (one warp, once per system and interval) ------------------------

// The deformation gradient dF/dt = L(t) F (minerals.py:426) from t0 to t1 with its own BS3(2)
// steps; lane 0 does the arithmetic, 27 lanes evaluate the stage matrices.  In the regimes
// without slip every grain turns at the same rate R(t) -- s V(t), V the left stretch of L F
// (regime 1: core.py:392-405, minerals.py:427-429), or s I (regimes 0 and 7: core.py:386-391)
// -- so Q = int R dt is integrated here as nine more unknowns and each grain just adds it.
__device__ __noinline__ int integrate_gradient(const IntegrateArgs &A, SysSlot &sl, int lane) {
    const int64_t s = sl.sys;
    const int regime = sl.regime;
    const bool with_Q = regime != 4 && regime != 6;
    const double t_begin = sl.t0, t_end = sl.t1, span = t_end - t_begin;
    const double relw = A.tol.atol_rel + A.tol.rtol, atol_abs = A.tol.atol_abs_F;
    const int si = lane / 9, c = lane - 9 * si;
    double F[9], kF[9], Q[9], R1[9];
    int status = 0;
    auto rate = [&](const double *Lm, const double *LF, double *R) {
        const double sv = strain_rate_max(Lm);
        if (regime == 1) {
            left_stretch(LF, R);
            for (int i = 0; i < 9; ++i) R[i] *= sv;
        } else {
            for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? sv : 0.0;
        }
    };
    if (lane < 9) sl.Lst[3][lane] = slot_L(A, sl, t_begin, lane);
    __syncwarp();
    if (lane == 0) {
        for (int i = 0; i < 9; ++i) {
            F[i] = A.F[s * 9 + i];
            Q[i] = 0.0;
        }
        matmul3(sl.Lst[3], F, kF);
        if (with_Q) rate(sl.Lst[3], kF, R1);
    }
    double t = t_begin, h = span;
    int steps = 0;
    while (span != 0.0) {
        if (steps++ >= A.tol.max_steps) {
            status = 1;
            break;
        }
        const double ts = (si == 0) ? fma(BS_A21, h, t) : ((si == 1) ? fma(BS_A32, h, t) : t + h);
        if (lane < 27) sl.Lst[si][c] = slot_L(A, sl, ts, c);
        __syncwarp();
        int verdict = 0;  // 1 accept, 2 accept and done, 0 reject, <0 failure
        double h_next = h;
        if (lane == 0) {
            double F2[9], F3[9], Fn[9], k2[9], k3[9], k4[9], R2[9], R3[9], R4[9], Qn[9];
            for (int i = 0; i < 9; ++i) F2[i] = F[i] + h * BS_A21 * kF[i];
            matmul3(sl.Lst[0], F2, k2);
            for (int i = 0; i < 9; ++i) F3[i] = F[i] + h * BS_A32 * k2[i];
            matmul3(sl.Lst[1], F3, k3);
            for (int i = 0; i < 9; ++i) Fn[i] = F[i] + h * (BS_B1 * kF[i] + BS_B2 * k2[i] + BS_B3 * k3[i]);
            matmul3(sl.Lst[2], Fn, k4);
            double em = 0.0;
            for (int i = 0; i < 9; ++i) {
                const double err = h * (BS_E1 * kF[i] + BS_E2 * k2[i] + BS_E3 * k3[i] + BS_E4 * k4[i]);
                const double r = fabs(err) / (atol_abs + relw * fabs(Fn[i]));
                em = (r != r) ? INFINITY : fmax(em, r);
            }
            if (with_Q) {
                rate(sl.Lst[0], k2, R2);
                rate(sl.Lst[1], k3, R3);
                rate(sl.Lst[2], k4, R4);
                for (int i = 0; i < 9; ++i) {
                    Qn[i] = Q[i] + h * (BS_B1 * R1[i] + BS_B2 * R2[i] + BS_B3 * R3[i]);
                    const double err = h * (BS_E1 * R1[i] + BS_E2 * R2[i] + BS_E3 * R3[i] + BS_E4 * R4[i]);
                    const double r = fabs(err) / (A.tol.atol_abs + relw * fabs(Qn[i]));
                    em = (r != r) ? INFINITY : fmax(em, r);
                }
            }
            if (!(em < INFINITY)) {
                verdict = -3;
            } else {
                const bool accept = em <= 1.0;
                double fac = (em > 0.0) ? DREX_SAFETY * rcbrt(em) : 5.0;
                fac = fmin(accept ? 5.0 : 1.0, fmax(0.2, fac));
                h_next = h * fac;
                if (accept) {
                    verdict = 1;
                    for (int i = 0; i < 9; ++i) {
                        F[i] = Fn[i];
                        kF[i] = k4[i];
                        if (with_Q) {
                            Q[i] = Qn[i];
                            R1[i] = R4[i];
                        }
                    }
                    const double remaining = t_end - (t + h);
                    if (fabs(remaining) <= 1e-13 * fmax(fabs(t_end), fabs(span))) verdict = 2;
                    else if (fabs(h_next) >= fabs(remaining) * (1.0 - 1e-10)) h_next = remaining;
                    else h_next = plan_step(h_next, remaining);
                } else if (fabs(h_next) <= 1e-14 * fmax(fabs(t), fabs(span))) {
                    verdict = -2;
                }
            }
        }
        verdict = __shfl_sync(0xffffffffu, verdict, 0);
        h_next = __shfl_sync(0xffffffffu, h_next, 0);
        if (verdict < 0) {
            status = -verdict;
            break;
        }
        if (verdict >= 1) t += h;
        if (verdict == 2) break;
        h = h_next;
    }
    if (lane == 0) {
        for (int i = 0; i < 9; ++i) {
            A.F[s * 9 + i] = F[i];
            sl.Q[i] = Q[i];
        }
    }
    return status;
}

// Take the next system from the queue and set slot (seq mod INT_NCTX) up for it; publishes the
// slot by storing its ticket last.  Called by one whole warp.
__device__ __noinline__ void prepare_system(const IntegrateArgs &A, const FabricTable &T, CtaShared &sh,
                                            int seq, int lane) {
    SysSlot &sl = sh.slot[seq % INT_NCTX];
    unsigned spins = 0;
    while (ld_vol(&sl.busy))  // the system that last used this slot is being finished
        spin_pause(spins, 1, seq, sl.seq, sl.sys);
    int s = -1;
    if (lane == 0) {
        const int ticket = atomicAdd(A.queue, 1);
        if (ticket < A.S) s = (A.order != nullptr) ? A.order[ticket] : ticket;
    }
    s = __shfl_sync(0xffffffffu, s, 0);
    if (s < 0) {
        if (lane == 0) st_vol(&sh.end_seq, seq);
        return;
    }
    const int64_t ip = A.sys_particle[s];
    const int regime = A.sys_regime[s], fab = A.sys_fabric[s];
    const int K = A.K;
    const bool slip = regime == 4 || regime == 6;
    const int phase = T.fab[fab].phase;
    const double rscale = (regime == 6) ? 0.3 : 1.0;  // core.py:459,464
    // grain-boundary migration active? (enstatite strain energy is identically zero,
    // core.py:674-684 loops over systems 0..2 only)
    const double kappa = (slip && phase == 0) ? rscale * A.sys_vol_frac[s] * A.gbm_mobility : 0.0;
    const int n_chunks = (int)((A.N + INT_CHUNK - 1) / INT_CHUNK);
    if (lane == 0) {
        sl.busy = 1;
        sl.sys = s;
        sl.regime = regime;
        sl.fab = fab;
        sl.mode = !slip ? 4 : (phase != 0 ? GRAIN_ENSTATITE
                                          : ((T.fab[fab].default_exponents && T.fab[fab].inf_index == 3)
                                                 ? GRAIN_OLIVINE_DEFAULT
                                                 : GRAIN_OLIVINE_GENERIC));
        sl.t0 = A.t0[ip];
        sl.t1 = A.t1[ip];
        sl.kappa = kappa;
        sl.rscale = rscale;
        sl.inv_span2 = (sl.t1 != sl.t0) ? 2.0 / (sl.t1 - sl.t0) : 0.0;
        sl.f_active = kappa != 0.0;
        sl.n_chunks = n_chunks;
        sl.prep_at = n_chunks / 2;
        sl.const_flow = (A.L_kind == 0 || K == 1) ? 1 : 0;
        sl.K = K;
        sl.s_table = 0;
        sl.status = 0;
        sl.rhs_last = 0;
        sl.count_a = 0ull;
        sl.count_b = 0ull;
        sl.seq = seq;
    }
    if (K <= INT_KMAX)
        for (int j = lane; j < K * 9; j += 32) sl.Ltab[j] = A.L_nodes[ip * K * 9 + j];
    __syncwarp();
    if (A.L_kind == 0 || K == 1) {
        store_flows(&sl.cflow, 1, (lane < 9) ? sl.Ltab[lane] : 0.0, false, kappa, lane);
    } else if (A.L_kind == 2 && K <= INT_KMAX) {
        // Chebyshev coefficients of the interpolant through the K Lobatto nodes
        // x_j = (1 - cos(pi j / n)) / 2, n = K - 1, i.e. xi_j = 2 x_j - 1 = cos(pi (n - j) / n):
        //   a_k = (2 / n) sum''_j f_j T_k(xi_j),  T_k(xi_j) = cos(pi k (n - j) / n),
        // first and last term of the sum halved, a_0 and a_n halved.
        const int n = K - 1;
        for (int m = lane; m < 2 * n; m += 32) sl.ctab[m] = cospi((double)m / (double)n);
        __syncwarp();
        auto coefficient = [&](const double *f, int stride, int k) {
            double acc = 0.0;
            for (int j = 0; j <= n; ++j) {
                const double w = (j == 0 || j == n) ? 0.5 : 1.0;
                acc = fma(w * f[j * stride], sl.ctab[(k * (n - j)) % (2 * n)], acc);
            }
            return acc * (((k == 0 || k == n) ? 1.0 : 2.0) / (double)n);
        };
        for (int idx = lane; idx < 9 * K; idx += 32) {
            const int c = idx / K, k = idx - c * K;
            sl.coef[c][k] = coefficient(sl.Ltab + c, 9, k);
        }
        // the strain-rate scale s(t) = max |eig sym L(t)| as a tenth series, kept only if it
        // reproduces s at the midpoints between the nodes (s has kinks where the dominant
        // eigenvalue changes; then every stage evaluates the closed form instead)
        double *snode = sl.Lst[0];  // K <= 33 <= 36 doubles of scratch
        for (int j = lane; j < K; j += 32) snode[j] = strain_rate_max(sl.Ltab + 9 * j);
        __syncwarp();
        for (int k = lane; k < K; k += 32) sl.coef[9][k] = coefficient(snode, 1, k);
        __syncwarp();
        double worst = 0.0;
        for (int j = lane; j < n; j += 32) {
            const double xi = -cospi(((double)j + 0.5) / (double)n);
            double Lm[9];
            for (int c = 0; c < 9; ++c) Lm[c] = clenshaw(sl.coef[c], K, xi);
            const double se = strain_rate_max(Lm), si = clenshaw(sl.coef[9], K, xi);
            const double rel = fabs(si - se) / se;
            worst = (rel == rel) ? fmax(worst, rel) : INFINITY;
        }
        worst = warp_max(worst);
        __syncwarp();
        // 1e-9: s only enters through the strain energy (the rotation rate is homogeneous of
        // degree one in (D, L), so D / s and L / s times s do not depend on s)
        if (lane == 0) sl.s_table = (worst <= 1e-9) ? 1 : 0;
        __syncwarp();
    }
    {
        // carried step sizes are kept in units of STRAIN (time x strain-rate scale): along a
        // pathline the strain rate changes from one interval to the next and a step carried as
        // time would be rejected whenever the flow speeds up
        double Lm[9], sb = 1.0, se = 1.0;
        if (lane < 2) {
            for (int c = 0; c < 9; ++c) Lm[c] = slot_L(A, sl, lane == 0 ? sl.t0 : sl.t1, c);
            const double sv = strain_rate_max(Lm);
            if (lane == 0) sb = (sv > 0.0 && sv < INFINITY) ? sv : 1.0;
            else se = (sv > 0.0 && sv < INFINITY) ? sv : 1.0;
        }
        if (lane == 0) sl.s_begin = sb;
        if (lane == 1) sl.s_end = se;
        __syncwarp();
    }
    if (!sl.const_flow) {
        if (sl.s_table) warp_flows_cheb(sl, &sl.flow_t0, 1, sl.t0, sl.t0, sl.t0, lane);
        else warp_flows(A, sl, &sl.flow_t0, 1, sl.t0, sl.t0, sl.t0, lane);
    }
    if (lane < FLOW_CACHE) {
        sl.fcache[lane].key = 0ull;
        sl.fc_ready[lane] = 0;
    }
    const int status0 = integrate_gradient(A, sl, lane);
    if (lane == 0) sl.status0 = status0;
    __threadfence_block();
    __syncwarp();
    if (lane == 0)
        *reinterpret_cast<volatile unsigned long long *>(&sl.ticket) = (unsigned long long)(unsigned)seq << 32;
}

// Per-slot scratch in global memory (L2-resident): per-chunk sums.
struct SlotScratch {
    double *psum;
};
__device__ __forceinline__ SlotScratch slot_scratch(const IntegrateArgs &A, int slot_index) {
    SlotScratch sc;
    sc.psum = A.scratch + ((int64_t)blockIdx.x * INT_NCTX + slot_index) * A.chunk_cap;
    return sc;
}

// End of the interval for one system, by the warp that completed its last chunk:
// extract_vars -> apply_gbs against the interval-start orientations -> extract_vars
// (minerals.py:464-474, 536-538; utils.py:159-190).  f_out holds u_g = max(f_g(t0), 0) e^{v_g}
// and psum the per-chunk sums of u; every sum is formed in a fixed order.  One warp walks
// 40 KB of fractions twice, so the loads of FIN_BATCH chunks are issued together: the pass is
// a chain of L2 round trips, not arithmetic.  The ORIENTATIONS of sliding grains are not
// restored here -- that is a copy of up to 360 KB per system, latency-bound for a single warp --
// but by gbs_restore_kernel, which streams them at memory bandwidth after this kernel; this
// pass leaves it one lane mask per chunk.
constexpr int FIN_BATCH = 8;
__device__ __noinline__ void finalize_system(const IntegrateArgs &A, SysSlot &sl, const SlotScratch sc,
                                             int lane) {
    const int64_t N = A.N, s = sl.sys;
    const int nq = (int)((N + 31) / 32);  // the sweeps and the masks go by 32 grains
    double *f_out = A.f_out + s * N;
    uint32_t *mask_out = A.slide_mask + s * nq;
    double acc = 0.0;
    for (int q = lane; q < sl.n_chunks; q += 32) acc += __ldcg(sc.psum + q);
    const double S1 = warp_sum(acc);
    const double floor_v = A.gbs_threshold / (double)N;
    // first sweep: S2 = sum of the fractions after sliding grains were set to chi / N
    acc = 0.0;
    for (int q0 = 0; q0 < nq; q0 += FIN_BATCH) {
        double u[FIN_BATCH];
#pragma unroll
        for (int j = 0; j < FIN_BATCH; ++j) {
            const int64_t g = (int64_t)(q0 + j) * 32 + lane;
            u[j] = (q0 + j < nq && g < N) ? __ldcg(f_out + g) : -1.0;
        }
#pragma unroll
        for (int j = 0; j < FIN_BATCH; ++j) {
            const double fh = u[j] / S1;
            acc += (u[j] < 0.0) ? 0.0 : ((fh < floor_v) ? floor_v : fh);  // NaN passes through
        }
    }
    const double S2 = warp_sum(acc);
    // The last renormalisation of the reference divides by sum_g (f_g / S2), which is 1 up to
    // a few ulp whatever the order of summation; it is taken as exactly 1 here.
    // second sweep: the new fractions, and which grains slide
    for (int q0 = 0; q0 < nq; q0 += FIN_BATCH) {
        double u[FIN_BATCH];
#pragma unroll
        for (int j = 0; j < FIN_BATCH; ++j) {
            const int64_t g = (int64_t)(q0 + j) * 32 + lane;
            u[j] = (q0 + j < nq && g < N) ? __ldcg(f_out + g) : -1.0;
        }
        uint32_t mine = 0;  // lane j keeps the mask of chunk q0 + j: one coalesced store per batch
#pragma unroll
        for (int j = 0; j < FIN_BATCH; ++j) {
            const int64_t g = (int64_t)(q0 + j) * 32 + lane;
            bool slide = false;
            if (!(u[j] < 0.0)) {  // live grain (or NaN, which is written through)
                const double fh = u[j] / S1;
                slide = fh < floor_v;
                f_out[g] = fmax((slide ? floor_v : fh) / S2, 0.0);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, slide);
            if (lane == j) mine = m;
        }
        if (lane < FIN_BATCH && q0 + lane < nq) mask_out[q0 + lane] = mine;
    }
    if (lane == 0) {
        const int status = max(sl.status, sl.status0);
        const unsigned long long ca = sl.count_a, cb = sl.count_b;
        const long long n_rhs = (long long)(ca & 0xffffffffull);
        const int64_t live_last = N - (int64_t)(sl.n_chunks - 1) * INT_CHUNK;
        A.status[s] = status;
        if (A.n_rhs) A.n_rhs[s] = (int32_t)n_rhs;
        if (A.n_accept) A.n_accept[s] = (int32_t)(cb & 0xffffffffull);
        if (A.n_reject) A.n_reject[s] = (int32_t)(ca >> 32);
        if (A.grain_evals) A.grain_evals[s] = INT_CHUNK * n_rhs - (INT_CHUNK - live_last) * (int64_t)sl.rhs_last;
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) st_vol(&sl.busy, 0);
}

// ---- chunk tickets -------------------------------------------------------------------------

struct ChunkRef {
    int seq, q;  // seq < 0: none (-1: the CTA is done, -2: would have to wait)
};

// Next chunk of the system currently handing out work.
__device__ __forceinline__ ChunkRef grab_chunk(CtaShared &sh, bool blocking, int lane) {
    int seq = -1, q = -1;
    if (lane == 0) {
        unsigned spins = 0;
        for (;;) {
            const int hs = ld_vol(&sh.hand_seq);
            if (hs >= ld_vol(&sh.end_seq)) break;
            SysSlot &sl = sh.slot[hs % INT_NCTX];
            const unsigned long long cur = ld_vol64(&sl.ticket);
            if ((int)(cur >> 32) != hs) {  // still being prepared
                if (!blocking) {
                    seq = -2;
                    break;
                }
                spin_pause(spins, 2, hs, (long long)cur, ld_vol(&sh.end_seq));
                continue;
            }
            const int t = (int)(cur & 0xffffffffull);
            if (t >= sl.n_chunks) {
                atomicCAS(&sh.hand_seq, hs, hs + 1);
                continue;
            }
            if (atomicCAS(&sl.ticket, cur, cur + 1ull) == cur) {
                seq = hs;
                q = t;
                break;
            }
        }
    }
    seq = __shfl_sync(0xffffffffu, seq, 0);
    q = __shfl_sync(0xffffffffu, q, 0);
    __threadfence_block();  // the slot's fields were written before its ticket was published
    return ChunkRef{seq, q};
}


__global__ void __launch_bounds__(INT_BLOCK, INT_CTAS_PER_SM)
integrate_kernel(const __grid_constant__ IntegrateArgs A, const __grid_constant__ FabricTable T) {
    extern __shared__ __align__(128) double dyn_smem[];
    __shared__ CtaShared sh;
    __shared__ __align__(16) double math_tab[MATH_TABLE_DOUBLES];
    load_math_tables(math_tab);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
    const int64_t N = A.N;
    double *rows = dyn_smem + warp * WARP_DOUBLES;
    ChunkFlow *wf = reinterpret_cast<ChunkFlow *>(dyn_smem + INT_WARPS * WARP_DOUBLES) + 3 * warp;
    uint64_t *mbar = &sh.mbar[warp];
    uint32_t parity = 0;
    if (lane == 0) mbar_init(mbar, 1);
    if (tid == 0) {
        sh.hand_seq = 0;
        sh.end_seq = 0x7fffffff;
        for (int i = 0; i < INT_NCTX; ++i) {  // shared memory comes with whatever the last kernel left
            sh.slot[i].ticket = ~0ull;
            sh.slot[i].busy = 0;
            sh.slot[i].seq = -1;
        }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
#ifdef DREX_PHASE_TIMERS
    long long ptimer[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long pmark = clock64();
#endif
    if (warp == 0) prepare_system(A, T, sh, 0, lane);
    DREX_MARK(1)

    // TMA needs 16-byte aligned rows: N even and aligned bases; otherwise plain loads
#ifdef DREX_NO_TMA_PREFETCH
    const bool rows_aligned = false;
#else
    const bool rows_aligned = ((N & 1) == 0) && ((reinterpret_cast<uintptr_t>(A.o_in) & 15) == 0) &&
                              ((reinterpret_cast<uintptr_t>(A.f_in) & 15) == 0);
#endif
    auto prefetch = [&](const ChunkRef &ch) {
        // ten rows of INT_CHUNK grains (nine of o_in, one of f_in) of chunk ch into rows ROW_PF..,
        // component c at row ROW_PF + c * INT_G: one copy per lane, issued by ten lanes at once
        if (!rows_aligned) return;
        const int64_t s = sh.slot[ch.seq % INT_NCTX].sys;
        const int64_t g0 = (int64_t)ch.q * INT_CHUNK;
        const uint32_t row_bytes = (uint32_t)(((N - g0 < INT_CHUNK) ? (N - g0) : INT_CHUNK) * sizeof(double));
        if (lane == 0) mbar_expect_tx(mbar, 10 * row_bytes);
        __syncwarp();
        if (lane < 9) bulk_g2s(rows + (ROW_PF + lane * INT_G) * 32, A.o_in + (s * 9 + lane) * N + g0, row_bytes, mbar);
        if (lane == 9) bulk_g2s(rows + (ROW_PF + 9 * INT_G) * 32, A.f_in + s * N + g0, row_bytes, mbar);
    };

    // deferred completion count of the last chunk this warp finished (see below)
    if (lane == 0) sh.ws_pend_slot[warp] = -1;
    __syncwarp();
    auto flush_pending = [&]() -> int {
        // returns the slot of a system this warp has to finish (it reported the last chunk), or -1
        const int pend_slot = ld_vol(&sh.ws_pend_slot[warp]);
        if (pend_slot < 0) return -1;
        SysSlot &ps = sh.slot[pend_slot];
        __threadfence_block();  // (one chunk later: nothing of this warp is in flight any more)
        __syncwarp();
        unsigned long long cb = 0ull;
        if (lane == 0) {
            cb = atomicAdd(&ps.count_b, (unsigned long long)ld_vol(&sh.ws_pend_acc[warp]) | (1ull << 32));
            st_vol(&sh.ws_pend_slot[warp], -1);
        }
        const int d = __shfl_sync(0xffffffffu, (int)(cb >> 32), 0);
        if (d != ps.n_chunks - 1) return -1;
        __threadfence_block();
        return pend_slot;
    };

    ChunkRef cur = grab_chunk(sh, true, lane);
    if (cur.seq >= 0) prefetch(cur);
    DREX_MARK(0)
    while (cur.seq >= 0) {
        SysSlot &sl = sh.slot[cur.seq % INT_NCTX];
        const int64_t s = sl.sys;
        const int64_t g = (int64_t)cur.q * INT_CHUNK + lane;  // + 32 j for the lane's j-th grain
        // The warp that drew chunk `prep_at` of a system prepares the NEXT system before starting
        // on its chunk, so that the slot is ready long before the current system runs out.  (Not
        // at the time of the draw: a warp still holding an unfinished chunk must never wait for
        // another system to be finished.)
        if (cur.q == sl.prep_at) {
            // (first everything this warp still owes to older systems: preparing waits for a free
            // slot, i.e. for an older system to be finished completely)
            const int fs = flush_pending();
            if (fs >= 0) finalize_system(A, sh.slot[fs], slot_scratch(A, fs), lane);
            prepare_system(A, T, sh, cur.seq + 1, lane);
        }
        DREX_MARK(1)
        // ---- this chunk's state into rows ROW_A.., then the prefetch rows are free again
        double f0[INT_G];
        if (rows_aligned) {
            mbar_wait(mbar, parity);
            parity ^= 1u;
#pragma unroll
            for (int j = 0; j < INT_G; ++j) {
                const bool live = g + 32 * j < N;
                const double *pf = rows + (ROW_PF + j) * 32 + lane;
                double *r = rows + (j * ROW_SET + ROW_A) * 32 + lane;
#pragma unroll
                for (int c = 0; c < 9; ++c) r[c * 32] = live ? pf[c * INT_G * 32] : 0.0;
                f0[j] = live ? pf[9 * INT_G * 32] : 0.0;
            }
            // generic-proxy reads of the prefetch rows, then an async-proxy overwrite of them
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        } else {
#pragma unroll
            for (int j = 0; j < INT_G; ++j) {
                const int64_t gj = g + 32 * j;
                const bool live = gj < N;
                double *r = rows + (j * ROW_SET + ROW_A) * 32 + lane;
#pragma unroll
                for (int c = 0; c < 9; ++c) r[c * 32] = live ? A.o_in[(s * 9 + c) * N + gj] : 0.0;
                f0[j] = live ? A.f_in[s * N + gj] : 0.0;
            }
        }
        __syncwarp();
        DREX_MARK(2)
        ChunkRef nxt = grab_chunk(sh, false, lane);
        if (nxt.seq >= 0) prefetch(nxt);
        DREX_MARK(0)

        // park the warp-uniform loop state in shared memory across the integration
        if (lane == 0) {
            sh.ws_cur_seq[warp] = cur.seq;
            sh.ws_cur_q[warp] = cur.q;
            sh.ws_nxt_seq[warp] = nxt.seq;
            sh.ws_nxt_q[warp] = nxt.q;
        }
        __syncwarp();
        ChunkResult res;
#pragma unroll
        for (int j = 0; j < INT_G; ++j) {
            f0[j] = fmax(f0[j], 0.0);  // utils.py:188
            res.v[j] = 0.0;
        }
        res.h_next = 0.0;
        res.status = res.n_rhs = res.n_acc = res.n_rej = 0;
        res.row = ROW_A;
        const int mode = sl.mode;
        if (mode != 4) {
            // first trial step of this chunk: carried from the previous interval, else the
            // whole interval (or first_step_frac of it)
            const double span = sl.t1 - sl.t0;
            double h = (A.h_io != nullptr) ? A.h_io[s * sl.n_chunks + cur.q] / sl.s_begin : 0.0;
            if (!(fabs(h) > 0.0)) h = (A.tol.first_step_frac > 0.0 ? A.tol.first_step_frac : 1.0) * span;
            h = copysign(fabs(h), span);
            if (A.tol.max_step_frac > 0.0 && fabs(h) > A.tol.max_step_frac * fabs(span))
                h = A.tol.max_step_frac * span;
            if (!(fabs(h) < fabs(span))) h = span;  // a step carried over from a longer interval
            if (DREX_STRETCH > 1.0 && span != 0.0) h = plan_step(h, span);
            // broadcast from lane 0: provably warp-uniform, so the fabric constants are addressed
            // through uniform registers / the constant bank instead of ~30 vector registers
            const int fab_u = __shfl_sync(0xffffffffu, sl.fab, 0);
            const FabricConsts &fc = T.fab[fab_u];
            if (mode == GRAIN_ENSTATITE) integrate_chunk<GRAIN_ENSTATITE>(A, sl, fc, math_tab, rows, wf, lane, f0, h, res DREX_TIMER_ARGS);
            else if (mode == GRAIN_OLIVINE_DEFAULT) integrate_chunk<GRAIN_OLIVINE_DEFAULT>(A, sl, fc, math_tab, rows, wf, lane, f0, h, res DREX_TIMER_ARGS);
            else integrate_chunk<GRAIN_OLIVINE_GENERIC>(A, sl, fc, math_tab, rows, wf, lane, f0, h, res DREX_TIMER_ARGS);
        }
        DREX_MARK(7)
        cur.seq = ld_vol(&sh.ws_cur_seq[warp]);
        cur.q = ld_vol(&sh.ws_cur_q[warp]);
        nxt.seq = ld_vol(&sh.ws_nxt_seq[warp]);
        nxt.q = ld_vol(&sh.ws_nxt_q[warp]);
        const int fin_slot = flush_pending();  // the previous chunk's completion (fast by now)
        // ---- results of the chunk: clipped orientations (utils.py:187), u_g = clip(f_g) e^{v_g}
        // (everything is derived again from the reloaded chunk reference, so that nothing but f0
        // has to survive the integration in registers)
        {
            SysSlot &sr = sh.slot[cur.seq % INT_NCTX];
            const int64_t ss = sr.sys;
            const bool slip = sr.mode != 4;
            double usum = 0.0;
#pragma unroll
            for (int j = 0; j < INT_G; ++j) {
                const int64_t gg = (int64_t)cur.q * INT_CHUNK + 32 * j + lane;
                const bool alive = gg < N;
                const double *fin = rows + (j * ROW_SET + res.row) * 32 + lane;
                double *o_out = A.o_out + ss * 9 * N + gg;
                if (alive) {
#pragma unroll
                    for (int c = 0; c < 9; ++c) {
                        const double y = slip ? fin[c * 32] : fin[c * 32] + sr.Q[c];
                        o_out[c * N] = clip_int(y);
                    }
                }
                const double u = sr.f_active ? f0[j] * exp_tab(res.v[j], math_tab) : f0[j];
                if (alive) A.f_out[ss * N + gg] = u;
                usum += alive ? u : 0.0;
            }
            const double part = warp_sum(usum);
            if (lane == 0) {
                slot_scratch(A, cur.seq % INT_NCTX).psum[cur.q] = part;
                if (A.h_io != nullptr && slip) A.h_io[ss * sr.n_chunks + cur.q] = res.h_next * sr.s_end;
                if (cur.q == sr.n_chunks - 1) sr.rhs_last = res.n_rhs;
                if (res.status) atomicMax(&sr.status, res.status);
                atomicAdd(&sr.count_a, (unsigned long long)res.n_rhs | ((unsigned long long)res.n_rej << 32));
            }
        }
        // This chunk's completion is reported when the warp finishes its NEXT chunk (or runs out
        // of work): the count must follow the stores above, and a fence right here waits a full
        // round trip to L2 for them; one chunk later they have long landed.
        if (lane == 0) {
            sh.ws_pend_slot[warp] = cur.seq % INT_NCTX;
            sh.ws_pend_acc[warp] = res.n_acc;
        }
        __syncwarp();
        DREX_MARK(4)
        if (fin_slot >= 0) finalize_system(A, sh.slot[fin_slot], slot_scratch(A, fin_slot), lane);
        DREX_MARK(5)
        if (nxt.seq == -2) {
            // the next system is still being prepared: report first, then wait for it
            const int fs = flush_pending();
            if (fs >= 0) finalize_system(A, sh.slot[fs], slot_scratch(A, fs), lane);
            nxt = grab_chunk(sh, true, lane);
            if (nxt.seq >= 0) prefetch(nxt);
        }
        DREX_MARK(0)
        cur = nxt;
    }
    {
        const int fs = flush_pending();
        if (fs >= 0) finalize_system(A, sh.slot[fs], slot_scratch(A, fs), lane);
    }
    DREX_MARK(5)
#ifdef DREX_PHASE_TIMERS
    if (lane == 0) {
        unsigned long long *out = reinterpret_cast<unsigned long long *>(A.queue) + 8;
        for (int i = 0; i < 8; ++i) atomicAdd(out + i, (unsigned long long)ptimer[i]);
    }
#endif
}

// Grain-boundary sliding, orientations: a sliding grain keeps its interval-start orientation
// (utils.py:170), i.e. o_prev is copied over what the integration wrote for it.  Runs after
// integrate_kernel on the same stream: one warp per chunk whose mask is not empty, nine
// coalesced loads and masked stores; pure streaming at memory bandwidth (up to 144 B per
// sliding grain), nothing when no grain slides.
constexpr int RESTORE_BLOCK = 256;
__global__ void __launch_bounds__(RESTORE_BLOCK)
gbs_restore_kernel(const uint32_t *__restrict__ slide_mask, const double *__restrict__ o_prev,
                   double *__restrict__ o_out, int64_t S, int64_t N) {
    const int lane = threadIdx.x & 31;
    const int64_t n_chunks = (N + 31) / 32, total = S * n_chunks;
    const int64_t warp0 = ((int64_t)blockIdx.x * RESTORE_BLOCK + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * RESTORE_BLOCK) >> 5;
    for (int64_t base = warp0 * 32; base < total; base += n_warps * 32) {
        // 32 masks per warp and round, one per lane; then the warp walks the non-empty ones
        const uint32_t mine = (base + lane < total) ? slide_mask[base + lane] : 0u;
        uint32_t todo = __ballot_sync(0xffffffffu, mine != 0u);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t m = __shfl_sync(0xffffffffu, mine, j);
            const int64_t idx = base + j, s = idx / n_chunks, q = idx - s * n_chunks;
            if ((m >> lane) & 1u) {
                const int64_t off = s * 9 * N + q * 32 + lane;
                double a[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) a[c] = o_prev[off + c * N];
#pragma unroll
                for (int c = 0; c < 9; ++c) o_out[off + c * N] = clip1(a[c]);
            }
        }
    }
}

// Work-queue order of the next launch: systems by measured cost, most expensive first, ties by
// index.  One CTA sorts 64-bit keys (complement of the cost | index) with the all-ascending
// bitonic network over exactly S entries (virtual padding stays at the end and is never touched).
constexpr int ORDER_THREADS = 1024;
constexpr int ORDER_MAX = 16384;
__global__ void __launch_bounds__(ORDER_THREADS)
next_order_kernel(const int64_t *__restrict__ grain_evals, const int32_t *__restrict__ sys_phase, int n,
                  int32_t *__restrict__ order) {
    extern __shared__ __align__(16) unsigned long long order_keys[];
    for (int i = threadIdx.x; i < n; i += ORDER_THREADS) {
        unsigned long long cost = (unsigned long long)grain_evals[i] * (sys_phase[i] == 0 ? 4ull : 1ull);
        if (cost >> 49) cost = (1ull << 49) - 1;
        order_keys[i] = ((~cost & ((1ull << 49) - 1)) << 14) | (unsigned long long)i;
    }
    __syncthreads();
    int npad = 2;
    while (npad < n) npad <<= 1;
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k; j > 0; j = (j == k) ? (k >> 2) : (j >> 1)) {
            // first step of a merge: mirror image inside the block of k; then strides k/4, k/8, ...
            for (int i = threadIdx.x; i < n; i += ORDER_THREADS) {
                const int p = (j == k) ? (i ^ (k - 1)) : (i ^ j);
                if (p > i && p < n) {
                    const unsigned long long a = order_keys[i], b = order_keys[p];
                    if (b < a) {
                        order_keys[i] = b;
                        order_keys[p] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += ORDER_THREADS) order[i] = (int32_t)(order_keys[i] & 0x3fffull);
}

}  // namespace drex
