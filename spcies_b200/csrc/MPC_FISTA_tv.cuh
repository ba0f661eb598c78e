// MPC_FISTA_tv.cuh -- batched laxMPC (SPCIES_TERMINAL == 1) / equMPC (== 0) FISTA solver with a *per-instance* model (options.time_varying, `#define TIME_VARYING 1`),
// hand-written for sm_100a.  Every instance brings its own A, B (column-major, as MATLAB passes them), diagonal Q, R and bounds:
// the "one shared model" of the other kernels becomes "a model per instance", and the factorisation that the generator does off
// line for a fixed model runs on the device, once per instance:
//
//   QRi = -1 ./ [Q; R];  A Q^-1 A', B R^-1 B'                                          code_laxMPC_FISTA_C.c:109-153
//   block-Cholesky recursion of W = G H^-1 G' (block tridiagonal): Beta_0, Alpha_0, then
//   Beta_h = chol(A Qi A' + B Ri B' + Qi - Alpha_{h-1}' Alpha_{h-1}) (diagonal stored inverted), Alpha_h = Beta_h^-T (-Qi A'),
//   Beta_{N-1} with the terminal weight                                                :155-262
//   then the FISTA loop of the constant-model solver on that instance's Alpha / Beta / [A B]     :275-389, :471-651
//
// The shared-matrix MMA formulation of MPC_FISTA_mma.cuh does not apply (there is no shared matrix): one thread owns one
// instance, its model, factor and iterates live in the per-instance state of the persistent skeleton ([element][thread],
// coalesced; ~9 KB per instance at N = 10, L2 resident).  Operation order is the reference's throughout, so Arith<EXACT> is
// bit-identical to the template compiled with -DTIME_VARYING=1.
#pragma once
#include "spcies_kernel.cuh"

#if !defined(TIME_VARYING) || TIME_VARYING != 1
#error "MPC_FISTA_tv.cuh is the TIME_VARYING == 1 path"
#endif
#if defined(VAR_BOUNDS)
#error "TIME_VARYING with per-stage bounds is rejected by the generator (cons_laxMPC_FISTA_C.m:51)"
#endif

namespace spcies {
namespace fista_tv {

constexpr int n = nn_, m = mm_, nm = nm_, N = NN_;

struct Solver {
    typedef SPCIES_REAL real;
    static constexpr int OFF_AB = 0;                            // [n][nm]
    static constexpr int OFF_Q = OFF_AB + n * nm;               // Q (negated after the factorisation), [n]
    static constexpr int OFF_R = OFF_Q + n;                     // [m]
    static constexpr int OFF_QRI = OFF_R + m;                   // [nm]
    static constexpr int OFF_ALPHA = OFF_QRI + nm;              // [N-1][n][n]
    static constexpr int OFF_BETA = OFF_ALPHA + (N - 1) * n * n;   // [N][n][n]
    static constexpr int OFF_LB = OFF_BETA + N * n * n;         // [nm]
    static constexpr int OFF_UB = OFF_LB + nm;
    static constexpr int OFF_Z = OFF_UB + nm;                   // z[N-1][nm]
    static constexpr int OFF_Z0 = OFF_Z + (N - 1) * nm;         // [m]
    static constexpr int OFF_ZN = OFF_Z0 + m;                   // [n]
    static constexpr int OFF_Y = OFF_ZN + n;                    // [N][n]
    static constexpr int OFF_LAM = OFF_Y + N * n;
    static constexpr int OFF_LAM1 = OFF_LAM + N * n;
    static constexpr int OFF_DL = OFF_LAM1 + N * n;             // residual / d_lambda
    static constexpr int OFF_B = OFF_DL + N * n;                // [n]
    static constexpr int OFF_QV = OFF_B + n;                    // q [nm]
    static constexpr int OFF_QT = OFF_QV + nm;                  // qT [n]
    static constexpr int OFF_T = OFF_QT + n;                    // t
    static constexpr int OFF_TMP = OFF_T + 1;                   // A Qi A' [n][n], B Ri B' [n][n], Q_i [n], R_i [m] (factorisation only)
    static constexpr int STATE = OFF_TMP + 2 * n * n + n + m;
    static constexpr int STATE_VARB = STATE;
    static constexpr bool HAS_VARB = true;                      // LB_in / UB_in are part of the signature

    template <class A, bool VARB, class ST> struct Ctx {
        const spcies_consts *C;
        ST s;
        const BatchIO &io;
        __device__ Ctx(const spcies_consts *C_, ST s_, const BatchIO &io_) : C(C_), s(s_), io(io_) {}

        __device__ __forceinline__ real ab(int i, int j) const { return s.ld(OFF_AB + i * nm + j); }
        __device__ __forceinline__ real al(int h, int i, int j) const { return s.ld(OFF_ALPHA + (h * n + i) * n + j); }
        __device__ __forceinline__ real be(int h, int i, int j) const { return s.ld(OFF_BETA + (h * n + i) * n + j); }
        __device__ __forceinline__ void set_al(int h, int i, int j, real v) const { s.st(OFF_ALPHA + (h * n + i) * n + j, v); }
        __device__ __forceinline__ void set_be(int h, int i, int j, real v) const { s.st(OFF_BETA + (h * n + i) * n + j, v); }

        // Beta_h (upper triangular, diagonal inverted) from `base`(i, j) [- Alpha_{h-1}' Alpha_{h-1}]; `diag`(i) is added to the
        // diagonal before the square root (Q_i, or -Ti for the last block)                                :155-176, :190-216, :237-261
        template <class Base, class Diag> __device__ void beta_block(int h, bool with_alpha, Base base, Diag diag) {
#pragma unroll 1
            for (int i = 0; i < n; ++i)
#pragma unroll 1
                for (int j = i; j < n; ++j) {
                    real v = base(i, j);
                    if (with_alpha) {
#pragma unroll 1
                        for (int k = 0; k < n; ++k) v = A::nmsub(v, al(h - 1, k, i), al(h - 1, k, j));
                    }
#pragma unroll 1
                    for (int l = 1; l <= i; ++l) v = A::nmsub(v, be(h, l - 1, i), be(h, l - 1, j));
                    if (i == j) {
                        v = diag(i, v);
                        v = A::div(real(1), A::sqrt(v));
                    } else {
                        v = A::mul(v, be(h, i, i));
                    }
                    set_be(h, i, j, v);
                }
        }
        // Alpha_h = Beta_h^-T (-Qi A')                                                                   :178-188, :219-232
        __device__ void alpha_block(int h) {
            const int QI = OFF_TMP + 2 * n * n;
#pragma unroll 1
            for (int i = 0; i < n; ++i)
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    real v = A::mul(-s.ld(QI + i), ab(j, i));
#pragma unroll 1
                    for (int l = 1; l <= i; ++l) v = A::nmsub(v, be(h, l - 1, i), al(h, l - 1, j));
                    set_al(h, i, j, A::mul(v, be(h, i, i)));
                }
        }

        __device__ void init(long long inst) {
            const int AQ = OFF_TMP, BR = OFF_TMP + n * n, QI = OFF_TMP + 2 * n * n, RI = QI + n;
            real x0[n], xr[n], ur[m];
#pragma unroll
            for (int i = 0; i < n; ++i) {
                x0[i] = (real)eng_x(C, io.x0, inst, n, i);
                xr[i] = (real)eng_x(C, io.xr, inst, n, i);
            }
#pragma unroll
            for (int i = 0; i < m; ++i) ur[i] = (real)eng_u(C, io.ur, inst, m, i);
            const double *Ain = io.ex[0] + inst * (long long)(n * n), *Bin = io.ex[1] + inst * (long long)(n * m);
            // bounds, weights, [A B], QRi                                                                  :83-137
#pragma unroll 1
            for (int i = 0; i < nm; ++i) {
#if defined(in_engineering) && in_engineering == 1
                const double sc = i < n ? (double)C->scaling_x[i] : (double)C->scaling_u[i - n];
                const double op = i < n ? (double)C->OpPoint_x[i] : (double)C->OpPoint_u[i - n];
                s.st(OFF_LB + i, (real)__dmul_rn(sc, __dsub_rn(io.LB[inst * nm + i], op)));
                s.st(OFF_UB + i, (real)__dmul_rn(sc, __dsub_rn(io.UB[inst * nm + i], op)));
#else
                s.st(OFF_LB + i, (real)io.LB[inst * nm + i]);
                s.st(OFF_UB + i, (real)io.UB[inst * nm + i]);
#endif
            }
#pragma unroll 1
            for (int i = 0; i < n; ++i) {
                const real q = (real)io.ex[2][inst * n + i];
                s.st(OFF_Q + i, q);
                const real qi = A::div(real(1), q);
                s.st(QI + i, qi);
                s.st(OFF_QRI + i, -qi);
                for (int j = 0; j < n; ++j) s.st(OFF_AB + i * nm + j, (real)Ain[i + j * n]);
                for (int j = 0; j < m; ++j) s.st(OFF_AB + i * nm + n + j, (real)Bin[i + j * n]);
            }
#pragma unroll 1
            for (int j = 0; j < m; ++j) {
                const real r = (real)io.ex[3][inst * m + j];
                s.st(OFF_R + j, r);
                const real ri = A::div(real(1), r);
                s.st(RI + j, ri);
                s.st(OFF_QRI + n + j, -ri);
            }
            // A Qi A', B Ri B'  (every term is (A_ik * Qi_k) * A_jk, summed in k order)                    :146-153
#pragma unroll 1
            for (int i = 0; i < n; ++i)
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    real a = real(0), b = real(0);
#pragma unroll 1
                    for (int k = 0; k < n; ++k) a = A::add(a, A::mul(A::mul((real)Ain[i + k * n], s.ld(QI + k)), (real)Ain[j + k * n]));
#pragma unroll 1
                    for (int k = 0; k < m; ++k) b = A::add(b, A::mul(A::mul((real)Bin[i + k * n], s.ld(RI + k)), (real)Bin[j + k * n]));
                    s.st(AQ + i * n + j, a);
                    s.st(BR + i * n + j, b);
                }
#pragma unroll 4
            for (int e = 0; e < (2 * N - 1) * n * n; ++e) s.st(OFF_ALPHA + e, real(0));   // memset(Alpha), memset(Beta)   :141-142
            // the recursion
            beta_block(0, false, [&](int i, int j) { return s.ld(BR + i * n + j); }, [&](int i, real v) { return A::add(v, s.ld(QI + i)); });
            alpha_block(0);
#pragma unroll 1
            for (int h = 1; h < N - 1; ++h) {
                beta_block(h, true, [&](int i, int j) { return A::add(s.ld(AQ + i * n + j), s.ld(BR + i * n + j)); },
                           [&](int i, real v) { return A::add(v, s.ld(QI + i)); });
                alpha_block(h);
            }
#if SPCIES_TERMINAL
            beta_block(N - 1, true, [&](int i, int j) { return A::add(s.ld(AQ + i * n + j), s.ld(BR + i * n + j)); },
                       [&](int i, real v) { return A::sub(v, C->Ti[i]); });                 // Ti is stored negated   :250-252
#else
            beta_block(N - 1, true, [&](int i, int j) { return A::add(s.ld(AQ + i * n + j), s.ld(BR + i * n + j)); },
                       [&](int, real v) { return v; });                                     // code_equMPC_FISTA_C.c:236-254
#endif
            // Q, R <- -Q, -R                                                                               :264-270
#pragma unroll
            for (int i = 0; i < n; ++i) s.st(OFF_Q + i, -s.ld(OFF_Q + i));
#pragma unroll
            for (int i = 0; i < m; ++i) s.st(OFF_R + i, -s.ld(OFF_R + i));
            // b, q, qT                                                                                     :274-289
#pragma unroll 1
            for (int j = 0; j < n; ++j) {
                real b = real(0);
#pragma unroll
                for (int i = 0; i < n; ++i) b = A::sub(b, A::mul(ab(j, i), x0[i]));
                s.st(OFF_B + j, b);
                s.st(OFF_QV + j, A::mul(s.ld(OFF_Q + j), xr[j]));
#if SPCIES_TERMINAL
                s.st(OFF_QT + j, A::mul(C->T[j], xr[j]));
#else
                s.st(OFF_QT + j, xr[j]);                       // equMPC: the terminal state is the reference   code_equMPC_FISTA_C.c:549
#endif
            }
#pragma unroll
            for (int j = 0; j < m; ++j) s.st(OFF_QV + n + j, A::mul(s.ld(OFF_R + j), ur[j]));
#pragma unroll 4
            for (int e = 0; e < 4 * N * n; ++e) s.st(OFF_Y + e, real(0));              // y = lambda = lambda1 = d_lambda = 0
            s.st(OFF_T, real(1));
            // initial steps                                                                               :298-320
            compute_z(OFF_LAM);
            residual();
            solve_W();
#pragma unroll 4
            for (int e = 0; e < N * n; ++e) {
                const real l = A::add(s.ld(OFF_LAM + e), s.ld(OFF_DL + e));
                s.st(OFF_LAM + e, l);
                s.st(OFF_Y + e, l);
            }
        }

        // z(lambda)                                                                                        :471-539
        __device__ void compute_z(int off_lam) {
#pragma unroll
            for (int j = 0; j < m; ++j) {
                real v = s.ld(OFF_QV + n + j);
#pragma unroll
                for (int i = 0; i < n; ++i) v = A::sub(v, A::mul(ab(i, n + j), s.ld(off_lam + i)));
                v = A::mul(v, s.ld(OFF_QRI + n + j));
                s.st(OFF_Z0 + j, clip(v, s.ld(OFF_LB + n + j), s.ld(OFF_UB + n + j)));
            }
#pragma unroll 1
            for (int l = 0; l < N - 1; ++l)
#pragma unroll 1
                for (int j = 0; j < nm; ++j) {
                    real v = s.ld(OFF_QV + j);
#pragma unroll
                    for (int i = 0; i < n; ++i) v = A::sub(v, A::mul(ab(i, j), s.ld(off_lam + (l + 1) * n + i)));
                    if (j < n) v = A::add(v, s.ld(off_lam + l * n + j));
                    v = A::mul(v, s.ld(OFF_QRI + j));
                    s.st(OFF_Z + l * nm + j, clip(v, s.ld(OFF_LB + j), s.ld(OFF_UB + j)));
                }
#if SPCIES_TERMINAL
#pragma unroll
            for (int j = 0; j < n; ++j) {
                real v = A::add(s.ld(OFF_QT + j), s.ld(off_lam + (N - 1) * n + j));
                v = A::mul(v, C->Ti[j]);
                s.st(OFF_ZN + j, clip(v, s.ld(OFF_LB + j), s.ld(OFF_UB + j)));
            }
#endif
        }
        // residual -> d_lambda                                                                             :546-574
        __device__ void residual() {
#pragma unroll 1
            for (int j = 0; j < n; ++j) {
                real v = A::add(s.ld(OFF_B + j), s.ld(OFF_Z + j));
#pragma unroll
                for (int i = 0; i < m; ++i) v = A::sub(v, A::mul(ab(j, n + i), s.ld(OFF_Z0 + i)));
                s.st(OFF_DL + j, v);
            }
#pragma unroll 1
            for (int l = 1; l < N; ++l)
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    real v = l < N - 1 ? s.ld(OFF_Z + l * nm + j) : (SPCIES_TERMINAL ? s.ld(OFF_ZN + j) : s.ld(OFF_QT + j));   // z_N | xr
#pragma unroll
                    for (int i = 0; i < nm; ++i) v = A::sub(v, A::mul(ab(j, i), s.ld(OFF_Z + (l - 1) * nm + i)));
                    s.st(OFF_DL + l * n + j, v);
                }
        }
        // W mu = r with the instance's block-Cholesky factor                                               :577-651
        __device__ void solve_W() {
#pragma unroll 1
            for (int l = 0; l < N; ++l)
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    real v = s.ld(OFF_DL + l * n + j);
                    if (l > 0) {
#pragma unroll 1
                        for (int i = 0; i < n; ++i) v = A::sub(v, A::mul(al(l - 1, i, j), s.ld(OFF_DL + (l - 1) * n + i)));
                    }
#pragma unroll 1
                    for (int i = 0; i < j; ++i) v = A::sub(v, A::mul(be(l, i, j), s.ld(OFF_DL + l * n + i)));
                    s.st(OFF_DL + l * n + j, A::mul(be(l, j, j), v));
                }
#pragma unroll 1
            for (int l = N - 1; l >= 0; --l)
#pragma unroll 1
                for (int j = n - 1; j >= 0; --j) {
                    real v = s.ld(OFF_DL + l * n + j);
                    if (l < N - 1) {
#pragma unroll 1
                        for (int i = n - 1; i >= 0; --i) v = A::sub(v, A::mul(al(l, j, i), s.ld(OFF_DL + (l + 1) * n + i)));
                    }
#pragma unroll 1
                    for (int i = n - 1; i >= j + 1; --i) v = A::sub(v, A::mul(be(l, j, i), s.ld(OFF_DL + l * n + i)));
                    s.st(OFF_DL + l * n + j, A::mul(be(l, j, j), v));
                }
        }

        __device__ bool iterate(int k) {
            // lambda1 = lambda, t1 = t                                                                      :326-328
#pragma unroll 4
            for (int e = 0; e < N * n; ++e) s.st(OFF_LAM1 + e, s.ld(OFF_LAM + e));
            const real t1 = s.ld(OFF_T);
            compute_z(OFF_Y);
            residual();
            bool over = false;
#pragma unroll 4
            for (int e = 0; e < N * n; ++e) over |= exceeds(s.ld(OFF_DL + e), (real)tol);
            if (!over || k >= k_max) return !over;                  // "the rest of the steps are unnecessary if done == 1"
            solve_W();
            const real t = A::mul(real(0.5), A::add(real(1), A::sqrt(A::add(real(1), A::mul(A::mul(real(4), t1), t1)))));
            s.st(OFF_T, t);
#pragma unroll 4
            for (int e = 0; e < N * n; ++e) {
                const real l = A::add(s.ld(OFF_Y + e), s.ld(OFF_DL + e));
                s.st(OFF_LAM + e, l);
                s.st(OFF_Y + e, A::add(l, A::div(A::mul(A::sub(t1, real(1)), A::sub(l, s.ld(OFF_LAM1 + e))), t)));
            }
            return false;
        }

        __device__ void finish(long long inst, int k, int ef) {
#pragma unroll
            for (int j = 0; j < m; ++j) io.u[inst * m + j] = eng_u_out(C, (double)s.ld(OFF_Z0 + j), j);
            io.k[inst] = k;
            io.e[inst] = ef;
            if (io.sol) {   // sol_<name>: z = (z_0, z, z_N), lambda = y                                     :411-446
                double *o = io.sol + inst * (long long)(sizeof(SPCIES_SOL_T) / sizeof(double));
                int c = 0;
                for (int j = 0; j < m; ++j) o[c++] = (double)s.ld(OFF_Z0 + j);
                for (int e = 0; e < (N - 1) * nm; ++e) o[c++] = (double)s.ld(OFF_Z + e);
#if SPCIES_TERMINAL
                for (int j = 0; j < n; ++j) o[c++] = (double)s.ld(OFF_ZN + j);
#endif
                for (int e = 0; e < N * n; ++e) o[c++] = (double)s.ld(OFF_Y + e);
                for (; c < (int)(sizeof(SPCIES_SOL_T) / sizeof(double)); ++c) o[c] = 0.0;
            }
        }
    };
};

typedef PolicyTraits<Solver> Traits;

}  // namespace fista_tv
}  // namespace spcies

#define SPCIES_TRAITS ::spcies::fista_tv::Traits
#define SPCIES_TIME_VARYING_ABI 1
#include "spcies_entry.cuh"
