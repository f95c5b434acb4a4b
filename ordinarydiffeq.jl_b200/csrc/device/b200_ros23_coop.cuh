// b200_ros23_coop.cuh — Rosenbrock23 for the lane-group kernel (b200_coop.cuh): B200_L >= n lanes of a warp own one
// trajectory, lane g holds component g of every vector and ROW g of W = J - I/(dt*gamma) in registers, and the dense
// linear algebra runs on warp shuffles inside the group:
//   * LU with partial pivoting: the pivot search is a shuffle reduction over the lanes (largest |W[r][k]|, smallest row
//     on ties — the scan order of a sequential search), the row exchange is a lane exchange, the pivot row is broadcast
//     entry by entry, every lane eliminates its own row;
//   * W \ b: forward substitution broadcasts y_k down the lanes, back substitution walks the rows upward with each
//     lane folding its own row in ascending column order.
// Every lane performs exactly the operations the sequential partial-pivot LU performs on its row, so the results equal
// the one-thread kernel's (and the oracle's) LU path bit for bit.  This is the "warp-per-trajectory LU with shuffles
// for larger n" of north_star; the one-thread kernel needs 254 registers plus 0.6-1 KB of local memory at n = 8.
//
// Reference behaviour reproduced: perform_step!(…, ::Rosenbrock23ConstantCache)
// (lib/OrdinaryDiffEqRosenbrock/src/rosenbrock_perform_step.jl:249-332), calc_rosenbrock_differentiation / calc_W
// (lib/OrdinaryDiffEqDifferentiation/src/derivative_utils.jl:937-1002,1050-1110), interpolant
// (rosenbrock_interpolants.jl:46-61); same stats accounting as b200_rosenbrock.cuh (nw += 1, njacs += 2 per attempt).
// Source contract (program option B200ODE_OPT_COMPONENT_RHS with a Rosenbrock algorithm):
//   rhs:    real NAME(int i, const real* u, const real* p, const real t)            -> du_i
//   jac:    real NAME(int i, int j, const real* u, const real* p, const real t)     -> J[i][j]
//   tgrad:  real NAME(int i, const real* u, const real* p, const real t)            -> dT_i      (optional)
#pragma once
#if B200_VLEN != 1
#error "the lane-group Rosenbrock23 keeps one component per lane: B200_L must be >= n"
#endif
#if B200_N < 2 || B200_N > 16
#error "the lane-group Rosenbrock23 serves 2 <= n <= 16"
#endif

struct B200Ros23Coop {
    real k1[1], k2[1];          // dense output rows, my component
    real f0[1], f2[1];          // fsalfirst, fsallast
    bool need_f0;               // initialize! is deferred to the first attempt (the RHS is a whole-warp operation)
    B200_STEPPER_EXTRA_MEMBERS

    static B200_D int order() { return 2; }
    static B200_D real qsteady_min() { return (real)1; }
    static B200_D real qsteady_max() { return (real)1.2; }

    B200_D real shfl(real v, int src) const {
#if B200_F32
        return __shfl_sync(gmask, v, src, B200_L);
#else
        const int lo = __shfl_sync(gmask, __double2loint(v), src, B200_L);
        const int hi = __shfl_sync(gmask, __double2hiint(v), src, B200_L);
        return __hiloint2double(hi, lo);
#endif
    }
    B200_D real shfl_xor(real v, int m) const {
#if B200_F32
        return __shfl_xor_sync(gmask, v, m, B200_L);
#else
        const int lo = __shfl_xor_sync(gmask, __double2loint(v), m, B200_L);
        const int hi = __shfl_xor_sync(gmask, __double2hiint(v), m, B200_L);
        return __hiloint2double(hi, lo);
#endif
    }

    // lu(W) with partial pivoting, rows across lanes.  row[] holds my row of W on entry and of the factors on exit.
    B200_D bool factor(real* row, int* piv) const {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < B200_N; ++k) {
            // pr = first row in k..n-1 with the largest |W[r][k]| (a NaN is skipped, except in row k itself)
            real v = (g >= k && g < B200_N) ? b200_abs(row[k]) : (real)-1;
            if (v != v) v = (g == k) ? b200_inf() : (real)-1;
            int idx = g;
#pragma unroll
            for (int off = B200_L / 2; off > 0; off >>= 1) {
                const real ov = shfl_xor(v, off);
                const int oi = __shfl_xor_sync(gmask, idx, off, B200_L);
                if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
            }
            const int pr = idx;
            piv[k] = pr;
            if (pr != k) {                                  // group-uniform
                const int partner = (g == k) ? pr : (g == pr ? k : g);
#pragma unroll
                for (int j = 0; j < B200_N; ++j) row[j] = shfl(row[j], partner);
            }
            const real pivot = shfl(row[k], k);
            if (pivot == (real)0) { ok = false; continue; } // group-uniform
            const bool below = (g > k && g < B200_N);
            real l = (real)0;
            if (below) { l = row[k] / pivot; row[k] = l; }
#pragma unroll
            for (int j = k + 1; j < B200_N; ++j) {
                const real pkj = shfl(row[j], k);
                if (below) row[j] = row[j] - l * pkj;
            }
        }
        return ok;
    }

    // x_g = (W \ b)_g for b distributed over the lanes
    B200_D real solve(const real* row, const int* piv, real b) const {
        real y = b;
#pragma unroll
        for (int k = 0; k < B200_N; ++k) {
            const int pr = piv[k];
            if (pr != k) {
                const int partner = (g == k) ? pr : (g == pr ? k : g);
                y = shfl(y, partner);
            }
            const real yk = shfl(y, k);
            if (g > k && g < B200_N) y = y - row[k] * yk;
        }
#pragma unroll
        for (int i = B200_N - 1; i >= 0; --i) {
            real s = y;
#pragma unroll
            for (int j = i + 1; j < B200_N; ++j) {
                const real yj = shfl(y, j);                 // rows below are final
                s = s - row[j] * yj;
            }
            const real xi = s / row[i];
            if (g == i) y = xi;
        }
        return y;
    }

    B200_D void init(const real*, const real*, real, int&) { need_f0 = true; }

    B200_D real attempt(const real* uprev, real* u, const real* p, real t, real dt, real reltol, real abstol,
                        int& nf, int& njacs, int& nw, int& nsolve, bool stepping) {
        const real d = (real)0.2928932188134525;       // convert(T, 1/(2+sqrt(2)))
        const real c32 = (real)7.414213562373095;      // convert(T, 6+sqrt(2))
        const real dtg = dt * d;
        const real ninv = -((real)1 / dtg);
        const real dto2 = dt / (real)2;
        const real dto6 = dt / (real)6;
        const bool mine = (g < B200_N);
        // initialize!(integrator, ::Rosenbrock23ConstantCache): fsalfirst = f(uprev, p, t); nf += 1 — on the first attempt
        // of a trajectory (uprev = u0, t = t0), evaluated by the whole warp if any group needs it
        if (__any_sync(0xffffffffu, need_f0)) {
            real tmp0[1];
            B200_RHS(tmp0, uprev, p, t);
            if (need_f0) { f0[0] = tmp0[0]; if (stepping) nf += 1; need_f0 = !stepping; }
        }
        // J and dT at (uprev, t): publish uprev, every lane evaluates its row
        real* Ub = sm + sbuf * B200_N;
        if (mine) Ub[g] = uprev[0];
        __syncwarp();
        real row[B200_N];
        int piv[B200_N];
        real dT = (real)0;
        {
            B200UserExact e;
#pragma unroll
            for (int j = 0; j < B200_N; ++j) row[j] = mine ? e.B200_USER_JAC_NAME(g, j, Ub, p, t) : (real)0;
#ifdef B200_USER_TGRAD_NAME
            if (mine) dT = e.B200_USER_TGRAD_NAME(g, Ub, p, t);
#endif
        }
        sbuf ^= 1;
        njacs += 2;
        nw += 1;
        // (members are committed at the end and only by a group that is really stepping: the other groups of the warp
        //  run through the same shuffles and barriers on stale state)
        const real lam = (real)1 / dtg;
#pragma unroll
        for (int j = 0; j < B200_N; ++j) if (g == j) row[j] = row[j] - lam;      // W = J - I * inv(dtgamma)
        const bool ok = factor(row, piv);
        // k1 = (W \ (fsalfirst + dtγ dT)) * (-1/dtγ)
        real x = solve(row, piv, b200_fma(dtg, dT, f0[0]));
        const real k1n = x * ninv;
        real tmp[1], f1[1], f2n[1];
        tmp[0] = b200_fma(dto2, k1n, uprev[0]);
        B200_RHS(f1, tmp, p, t + dto2);
        x = solve(row, piv, f1[0] - k1n);
        const real k2n = b200_fma(x, ninv, k1n);
        u[0] = b200_fma(dt, k2n, uprev[0]);
        B200_RHS(f2n, u, p, t + dt);
        x = solve(row, piv, b200_fma(dt, dT, b200_fma((real)-2, k1n - f0[0], b200_fma(-c32, k2n - f1[0], f2n[0]))));
        const real k3 = x * ninv;
        if (ok) { nf += 2; nsolve += 3; }       // a singular W returns before any solve (rosenbrock_perform_step.jl:271-274)
        if (stepping && ok) { k1[0] = k1n; k2[0] = k2n; f2[0] = f2n[0]; }
        const real ut = dto6 * (b200_fma((real)-2, k2n, k1n) + k3);
        // calculate_residuals + ODE_DEFAULT_NORM
        real res[1];
        res[0] = mine ? ut / b200_fma(b200_max_fast(b200_abs(uprev[0]), b200_abs(u[0])), reltol, abstol) : (real)0;
        const real EEst = B200_NORM(res, u);
        return ok ? EEst : (real)2;             // singular W: the step is rejected (rosenbrock_perform_step.jl:271-274)
    }

    B200_D void accept() { f0[0] = f2[0]; }
    B200_D void dense_prepare(const real*, const real*, const real*, real, real) {}

    B200_D void interp(real th, real dt, const real* y0, const real* /*y1*/, real* out) const {
        const real d = (real)0.2928932188134525;
        const real den = b200_fma((real)-2, d, (real)1);
        const real c1 = th * ((real)1 - th) / den;
        const real c2 = th * b200_fma((real)-2, d, th) / den;
        out[0] = b200_fma(dt, b200_fma(c2, k2[0], c1 * k1[0]), y0[0]);
    }
};
