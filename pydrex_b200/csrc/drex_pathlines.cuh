// drex_pathlines.cuh -- analytic steady flows and batched backward pathline tracing (sm_100a).
//
// Replaces the host-side input generators of the reference for whole particle clouds:
//   velocity._simple_shear_2d/_cell_2d/_corner_2d and their gradients (velocity.py:17-97)
//   pathlines.get_pathline (pathlines.py:10-116): from each particle's final location,
//     follow the steady flow BACKWARDS in time until the accumulated strain
//     (utils.strain_increment: |dt| * max|eig(sym L)|, utils.py:144-156) reaches
//     `max_strain` or the particle leaves the box, then hand out regular time stamps
//     (numpy.linspace(t_start, 0, steps + 1)) with the positions there.
// The reference calls scipy's LSODA once per particle (rtol 1e-5, atol 1e-8), accumulates the
// strain with a rectangle rule inside the event callback and resamples the dense output.
// Here one thread owns one particle: P independent 4-component ODEs (x, remaining strain)
// integrated with Dormand-Prince 5(4), the strain integrated as a state component to the
// same tolerance as the position.  Pass 1 finds t_start (events located by bisection on the
// Runge-Kutta step itself, no dense output needed), pass 2 walks from t = 0 to t_start again,
// landing exactly on every time the CPO integrator will ask for: the time stamps and the
// Chebyshev-Lobatto nodes inside each interval, where it stores L directly in the layout
// drex_integrate_f64 consumes (DREX_L_CHEBYSHEV).  The host never sees a position.
//
// This is latency-bound scalar work (10^4 particles x ~3000 field evaluations, < 1 ms); it is
// on the device to remove ~1 ms of host LSODA per particle and the table upload, not because
// it is heavy.  Output arrays are interval-major ([step][particle]...) so that thread p
// writes coalesced and each interval's table is one contiguous block.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "drex_common.cuh"

namespace drex {

struct FlowField {
    int kind;  // 0 simple_shear_2d, 1 cell_2d, 2 corner_2d
    int i0;    // direction / horizontal axis
    int i1;    // deformation plane / vertical axis
    double a;  // strain_rate / velocity_edge / plate_speed
    double b;  // - / edge_length / -
};

enum { FLOW_SIMPLE_SHEAR = 0, FLOW_CELL = 1, FLOW_CORNER = 2 };

enum {
    PATH_OK = 0,
    PATH_NOT_FINITE = 1,   // the field returned NaN (corner origin, outside the cell's domain)
    PATH_MAX_STEPS = 2,
    PATH_EMPTY = 3,        // the final location is outside the box or the backward path leaves at once
    PATH_TIME_LIMIT = 4    // neither event before t_limit: path ends there (the reference's 100 Myr)
};

// velocity (velocity.py:24-29, 56-71, 74-85); v may be null
__device__ __forceinline__ void flow_eval(const FlowField &fl, const double *x, double *v, double *L) {
    if (L) {
#pragma unroll
        for (int i = 0; i < 9; ++i) L[i] = 0.0;
    }
    if (v) v[0] = v[1] = v[2] = 0.0;
    const int a = fl.i0, b = fl.i1;
    if (fl.kind == FLOW_SIMPLE_SHEAR) {
        if (v) v[a] = x[b] * fl.a;
        if (L) L[3 * a + b] = 2.0 * fl.a;
    } else if (fl.kind == FLOW_CELL) {
        const double lim = 0.5 * fl.b;
        if (fabs(x[a]) > lim || fabs(x[b]) > lim) {  // reference raises ValueError
            if (v) v[0] = v[1] = v[2] = NAN;
            if (L) for (int i = 0; i < 9; ++i) L[i] = NAN;
            return;
        }
        double sx, cx, sz, cz;
        sincos(M_PI * x[a] / fl.b, &sx, &cx);
        sincos(M_PI * x[b] / fl.b, &sz, &cz);
        if (v) {
            v[a] = fl.a * cx * sz;
            v[b] = -fl.a * sx * cz;
        }
        if (L) {
            const double g = M_PI / fl.b;
            L[3 * a + a] = fl.a * (-g * sz * sx);
            L[3 * a + b] = fl.a * (g * cz * cx);
            L[3 * b + b] = fl.a * (-g * cz * cx);
            L[3 * b + a] = fl.a * (g * sz * sx);
        }
    } else {
        const double h = x[a], w = x[b];
        if (fabs(h) < 1e-15 && fabs(w) < 1e-15) {
            if (v) v[0] = v[1] = v[2] = NAN;
            if (L) for (int i = 0; i < 9; ++i) L[i] = NAN;
            return;
        }
        const double r2 = h * h + w * w;
        if (v) {
            const double pf = 2.0 * fl.a / M_PI;
            v[a] = pf * (atan2(h, -w) + h * w / r2);
            v[b] = pf * w * w / r2;
        }
        if (L) {
            const double pf = 4.0 * fl.a / (M_PI * r2 * r2);
            L[3 * a + a] = pf * (-(h * h) * w);
            L[3 * a + b] = pf * (h * h * h);
            L[3 * b + a] = pf * (-h * w * w);
            L[3 * b + b] = pf * (h * h * w);
        }
    }
}

// d/dt of (x, remaining strain): the strain left to accumulate shrinks going backwards
__device__ __forceinline__ void path_rhs(const FlowField &fl, const double *y, double *dy) {
    double L[9], D[9];
    flow_eval(fl, y, dy, L);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) D[3 * i + j] = 0.5 * (L[3 * i + j] + L[3 * j + i]);
    dy[3] = strain_rate_max(D);
}

struct PathBox {
    double lo[3], hi[3];
};

// pathlines._is_inside (pathlines.py:133-138): points ON the boundary are inside
__device__ __forceinline__ bool path_alive(const PathBox &box, const double *y) {
    bool inside = true;
#pragma unroll
    for (int i = 0; i < 3; ++i) inside = inside && !(y[i] < box.lo[i]) && !(y[i] > box.hi[i]);
    return inside && y[3] > 0.0;
}

// One Dormand-Prince 5(4) step of size h from (y, k1 = f(y)).  Returns the weighted error
// norm (max over components of |err| / (atol + rtol max(|y|,|ynew|))); k7 = f(ynew) (FSAL).
__device__ __forceinline__ double dopri_step(const FlowField &fl, const double *y, const double *k1,
                                             double h, double rtol, double atol, double *ynew,
                                             double *k7) {
    double k2[4], k3[4], k4[4], k5[4], k6[4], ys[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ys[i] = y[i] + h * (0.2 * k1[i]);
    path_rhs(fl, ys, k2);
#pragma unroll
    for (int i = 0; i < 4; ++i) ys[i] = y[i] + h * (3.0 / 40.0 * k1[i] + 9.0 / 40.0 * k2[i]);
    path_rhs(fl, ys, k3);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        ys[i] = y[i] + h * (44.0 / 45.0 * k1[i] - 56.0 / 15.0 * k2[i] + 32.0 / 9.0 * k3[i]);
    path_rhs(fl, ys, k4);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        ys[i] = y[i] + h * (19372.0 / 6561.0 * k1[i] - 25360.0 / 2187.0 * k2[i] +
                            64448.0 / 6561.0 * k3[i] - 212.0 / 729.0 * k4[i]);
    path_rhs(fl, ys, k5);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        ys[i] = y[i] + h * (9017.0 / 3168.0 * k1[i] - 355.0 / 33.0 * k2[i] + 46732.0 / 5247.0 * k3[i] +
                            49.0 / 176.0 * k4[i] - 5103.0 / 18656.0 * k5[i]);
    path_rhs(fl, ys, k6);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        ynew[i] = y[i] + h * (35.0 / 384.0 * k1[i] + 500.0 / 1113.0 * k3[i] + 125.0 / 192.0 * k4[i] -
                              2187.0 / 6784.0 * k5[i] + 11.0 / 84.0 * k6[i]);
    path_rhs(fl, ynew, k7);
    double err = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double e = h * (71.0 / 57600.0 * k1[i] - 71.0 / 16695.0 * k3[i] + 71.0 / 1920.0 * k4[i] -
                              17253.0 / 339200.0 * k5[i] + 22.0 / 525.0 * k6[i] - 1.0 / 40.0 * k7[i]);
        const double sc = atol + rtol * fmax(fabs(y[i]), fabs(ynew[i]));
        const double r = fabs(e) / sc;
        err = (r > err || r != r) ? r : err;  // NaN wins
    }
    return err;
}

__device__ __forceinline__ double dopri_factor(double err) {
    if (err == 0.0) return 5.0;
    return fmin(5.0, fmax(0.2, 0.9 * pow(err, -0.2)));
}

struct PathlineArgs {
    FlowField field;
    PathBox box;
    const double *final_locations;  // [P][3]
    int64_t P;
    double max_strain, t_limit, rtol, atol;
    int64_t n_intervals;    // S
    int n_nodes;            // K (>= 2) Chebyshev-Lobatto nodes per interval, or 0
    int timestamps_given;   // 1: timestamps[] holds the caller's ascending stamps (pass 1 skipped
                            //    unless t_start is requested)
    int max_steps;
    double *t_start;         // [P] out (may be null)
    double *timestamps;      // [S+1][P]
    double *positions;       // [S+1][P][3]
    double *strains;         // [S+1][P] strain accumulated since t_start (may be null)
    double *node_positions;  // [S][P][K][3] (may be null)
    double *L_nodes;         // [S][P][K][9] (may be null)
    double *step_times;      // [P][max_record] accepted times of pass 1, descending (may be null)
    int max_record;
    int *n_taken;            // [P] recorded times of pass 1: accepted steps + the event time (may be null)
    int *status;             // [P]
};

// advance (t, y, k1) to exactly t_target (< t) with adaptive steps; h is the signed step
// suggestion carried along.  Returns false on failure (status set).
__device__ __forceinline__ bool path_advance_to(const PathlineArgs &A, double t_target, double &t,
                                                double *y, double *k1, double &h, int &steps,
                                                int &status) {
    while (t > t_target) {
        const double remaining = t_target - t;  // negative
        bool last = false;
        double hs = h;
        if (!(fabs(hs) < fabs(remaining) * (1.0 - 1e-9))) {
            hs = remaining;
            last = true;
        }
        double ynew[4], k7[4];
        const double err = dopri_step(A.field, y, k1, hs, A.rtol, A.atol, ynew, k7);
        if (err != err) { status = PATH_NOT_FINITE; return false; }
        if (++steps > A.max_steps) { status = PATH_MAX_STEPS; return false; }
        if (err <= 1.0) {
            t = last ? t_target : t + hs;
#pragma unroll
            for (int i = 0; i < 4; ++i) { y[i] = ynew[i]; k1[i] = k7[i]; }
            // a step cut short by the target says nothing about the step the error allows
            const double grown = hs * dopri_factor(err);
            h = last ? ((fabs(grown) > fabs(h)) ? grown : h) : grown;
        } else {
            h = hs * dopri_factor(err);
        }
    }
    return true;
}

__global__ void __launch_bounds__(128) pathlines_kernel(PathlineArgs A) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.P) return;
    const int64_t P = A.P, S = A.n_intervals;
    const int K = A.n_nodes;
    int status = PATH_OK;
    double y0[4] = {A.final_locations[3 * p], A.final_locations[3 * p + 1],
                    A.final_locations[3 * p + 2], A.max_strain};
    double f0[4];
    path_rhs(A.field, y0, f0);
    if (!(f0[0] == f0[0]) || !(f0[3] == f0[3])) status = PATH_NOT_FINITE;

    // initial step: a hundredth of the time the strain budget or the box crossing would take
    double h0;
    {
        double span = 0.0, speed = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            span = fmax(span, A.box.hi[i] - A.box.lo[i]);
            speed = fmax(speed, fabs(f0[i]));
        }
        double scale = fabs(A.t_limit);
        if (f0[3] > 0.0) scale = fmin(scale, A.max_strain / f0[3]);
        if (speed > 0.0) scale = fmin(scale, span / speed);
        h0 = -0.01 * scale;
    }

    double t_start = 0.0;
    int taken = 0;
    if (status == PATH_OK && !A.timestamps_given) {
        // ---- pass 1: find t_start --------------------------------------------------------
        if (!path_alive(A.box, y0)) status = PATH_EMPTY;
        double t = 0.0, h = h0, y[4], k1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { y[i] = y0[i]; k1[i] = f0[i]; }
        int steps = 0;
        while (status == PATH_OK) {
            bool clipped = false;
            double hs = h;
            if (t + hs < A.t_limit) { hs = A.t_limit - t; clipped = true; }
            double ynew[4], k7[4];
            const double err = dopri_step(A.field, y, k1, hs, A.rtol, A.atol, ynew, k7);
            if (++steps > A.max_steps) { status = PATH_MAX_STEPS; break; }
            if (err != err) {
                // a trial point hit a singular spot of the field: retry much shorter, give up
                // once the step is negligible
                if (fabs(hs) < 1e-12 * fabs(A.t_limit)) { status = PATH_NOT_FINITE; break; }
                h = 0.1 * hs;
                continue;
            }
            if (err > 1.0) { h = hs * dopri_factor(err); continue; }
            if (!path_alive(A.box, ynew)) {
                // event inside this step: bisect on the step length (alive at 0, not at hs)
                double lo = 0.0, hi = hs;
                for (int it = 0; it < 64; ++it) {
                    const double mid = 0.5 * (lo + hi);
                    if (mid == lo || mid == hi) break;
                    double ym[4], km[4];
                    dopri_step(A.field, y, k1, mid, A.rtol, A.atol, ym, km);
                    if (path_alive(A.box, ym)) lo = mid; else hi = mid;
                }
                t_start = t + lo;
                if (t_start == 0.0) status = PATH_EMPTY;
                break;
            }
            t = clipped ? A.t_limit : t + hs;
#pragma unroll
            for (int i = 0; i < 4; ++i) { y[i] = ynew[i]; k1[i] = k7[i]; }
            if (A.step_times && taken < A.max_record) A.step_times[p * A.max_record + taken] = t;
            ++taken;
            if (clipped) { t_start = t; status = PATH_TIME_LIMIT; break; }
            h = hs * dopri_factor(err);
        }
        if (status == PATH_OK) {  // the event time closes the list
            if (A.step_times && taken < A.max_record) A.step_times[p * A.max_record + taken] = t_start;
            ++taken;
        }
    } else if (status == PATH_OK) {
        t_start = A.timestamps[p];
    }
    if (A.t_start) A.t_start[p] = t_start;
    if (A.n_taken) A.n_taken[p] = taken;
    const bool usable = (status == PATH_OK || status == PATH_TIME_LIMIT);

    // ---- time stamps: numpy.linspace(t_start, 0, S + 1) ---------------------------------------
    if (!A.timestamps_given && A.timestamps) {
        const double step = (S > 0) ? (0.0 - t_start) / (double)S : 0.0;
        for (int64_t j = 0; j <= S; ++j) {
            double tj = __dadd_rn(__dmul_rn((double)j, step), t_start);  // numpy rounds twice: no FMA
            if (j == S && S > 0) tj = 0.0;
            A.timestamps[j * P + p] = usable ? tj : NAN;
        }
    }
    if (!usable || !A.positions) {
        if (A.positions)
            for (int64_t j = 0; j <= S; ++j)
                for (int c = 0; c < 3; ++c) A.positions[(j * P + p) * 3 + c] = NAN;
        A.status[p] = status;
        return;
    }

    // ---- pass 2: from the last stamp back to the first, landing on every stamp and node ---
    double t = A.timestamps[S * P + p], h = h0, y[4], k1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { y[i] = y0[i]; k1[i] = f0[i]; }
    int steps = 0, st2 = PATH_OK;
    if (t < 0.0) {  // caller's stamps need not end at 0
        const double t_last = t;
        t = 0.0;
        path_advance_to(A, t_last, t, y, k1, h, steps, st2);
    }
    auto put_stamp = [&](int64_t j) {
        for (int c = 0; c < 3; ++c) A.positions[(j * P + p) * 3 + c] = y[c];
        if (A.strains) A.strains[j * P + p] = y[3];  // rebased below
    };
    auto put_node = [&](int64_t j, int k) {
        if (A.node_positions)
            for (int c = 0; c < 3; ++c) A.node_positions[((j * P + p) * K + k) * 3 + c] = y[c];
        if (A.L_nodes) {
            double L[9];
            flow_eval(A.field, y, nullptr, L);
            for (int c = 0; c < 9; ++c) A.L_nodes[((j * P + p) * K + k) * 9 + c] = L[c];
        }
    };
    put_stamp(S);
    for (int64_t j = S - 1; j >= 0 && st2 == PATH_OK; --j) {
        const double ta = A.timestamps[j * P + p], tb = A.timestamps[(j + 1) * P + p];
        if (K >= 2) {
            put_node(j, K - 1);
            for (int k = K - 2; k >= 1 && st2 == PATH_OK; --k) {
                const double xk = 0.5 * (1.0 - cospi((double)k / (double)(K - 1)));
                const double tk = ta + (tb - ta) * xk;
                if (path_advance_to(A, tk, t, y, k1, h, steps, st2)) put_node(j, k);
            }
        }
        if (st2 != PATH_OK) break;
        if (!path_advance_to(A, ta, t, y, k1, h, steps, st2)) break;
        if (K >= 2) put_node(j, 0);
        put_stamp(j);
    }
    if (st2 != PATH_OK) status = st2;
    if (A.strains && st2 == PATH_OK) {
        const double s_first = A.strains[p];  // remaining budget at t_start
        for (int64_t j = 0; j <= S; ++j) A.strains[j * P + p] -= s_first;
    }
    A.status[p] = status;
}

// velocity / gradient at arbitrary points (the callables of velocity.py evaluated in bulk)
__global__ void flow_field_kernel(FlowField fl, const double *__restrict__ x, int64_t P,
                                  double *__restrict__ v, double *__restrict__ L) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const double xp[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
    double vp[3], Lp[9];
    flow_eval(fl, xp, v ? vp : nullptr, L ? Lp : nullptr);
    if (v) for (int c = 0; c < 3; ++c) v[3 * p + c] = vp[c];
    if (L) for (int c = 0; c < 9; ++c) L[9 * p + c] = Lp[c];
}

}  // namespace drex
