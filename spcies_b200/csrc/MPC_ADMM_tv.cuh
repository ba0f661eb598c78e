// MPC_ADMM_tv.cuh -- batched equMPC (SPCIES_TERMINAL == 0) / laxMPC (== 1) ADMM solver with a *per-instance* model (options.time_varying, `#define TIME_VARYING 1`),
// hand-written for sm_100a.  Every instance brings its own A, B (column-major, as MATLAB passes them), diagonal Q, R and bounds;
// the factorisation the generator does off line for a fixed model runs on the device, once per instance:
//
//   Q_rho_i = 1 ./ (Q + rho), R_rho_i = 1 ./ (R + rho), Hi;  A Q_rho_i A', B R_rho_i B'                code_equMPC_ADMM_C.c:110-153
//   block-Cholesky recursion of W = G H_hat^-1 G': Beta_0, Alpha_0, Beta_h = chol(A Qi A' + B Ri B' + Qi - Alpha_{h-1}' Alpha_{h-1})
//   (diagonal stored inverted), Alpha_h = Beta_h^-T (-Qi A'), Beta_{N-1} without the Qi term          :155-255
//   then the ADMM loop of the constant-model solver on that instance's Alpha / Beta / [A B] / Hi      :291-553
// laxMPC (code_laxMPC_ADMM_C.c:108-633): the terminal state x_N with the dense Hi_N = T_rho_i = (T + rho I)^-1 (a generated
// constant: T is not an argument), added to every entry of the last diagonal block before its factorisation.
//
// One thread owns one instance; its model, factor and iterates live in the per-instance state of the persistent skeleton
// ([element][thread], coalesced, L2 resident).  Operation order is the reference's throughout, so Arith<EXACT> is bit-identical to
// the template compiled with -DTIME_VARYING=1 (the oracle of tests/test_time_varying_gpu.py).
#pragma once
#include "spcies_kernel.cuh"

#if !defined(TIME_VARYING) || TIME_VARYING != 1
#error "MPC_ADMM_tv.cuh is the TIME_VARYING == 1 path"
#endif
#if defined(VAR_BOUNDS) || !defined(SCALAR_RHO)
#error "TIME_VARYING needs fixed bounds along the horizon and a scalar rho (cons_equMPC_ADMM_C.m:46-51)"
#endif
#if SPCIES_TERMINAL != 0 && SPCIES_TERMINAL != 1
#error "MPC_ADMM_tv.cuh: the equMPC and laxMPC formulations"
#endif

namespace spcies {
namespace admm_tv {

constexpr int n = nn_, m = mm_, nm = nm_, N = NN_;
constexpr bool LAX = SPCIES_TERMINAL == 1;
constexpr int ZN = m + (N - 1) * nm;          // first element of the terminal state (laxMPC)
constexpr int ZLEN = ZN + (LAX ? n : 0);

struct Solver {
    typedef SPCIES_REAL real;
    static constexpr int OFF_AB = 0;                            // [n][nm]
    static constexpr int OFF_Q = OFF_AB + n * nm;               // Q (negated after the factorisation), [n]
    static constexpr int OFF_R = OFF_Q + n;                     // [m]
    static constexpr int OFF_HI = OFF_R + m;                    // Q_rho_i, R_rho_i: the rows of Hi (all equal), Hi_0 = its input part
    static constexpr int OFF_ALPHA = OFF_HI + nm;               // [N-1][n][n]
    static constexpr int OFF_BETA = OFF_ALPHA + (N - 1) * n * n;   // [N][n][n]
    static constexpr int OFF_LB = OFF_BETA + N * n * n;         // [nm]
    static constexpr int OFF_UB = OFF_LB + nm;
    static constexpr int OFF_V = OFF_UB + nm;                   // v_0[m], v[N-1][nm]
    static constexpr int OFF_LAM = OFF_V + ZLEN;
    static constexpr int OFF_Z = OFF_LAM + ZLEN;
    static constexpr int OFF_V1 = OFF_Z + ZLEN;
    static constexpr int OFF_MU = OFF_V1 + ZLEN;                // [N][n]
    static constexpr int OFF_B = OFF_MU + N * n;                // [n]
    static constexpr int OFF_QV = OFF_B + n;                    // q [nm]
    static constexpr int OFF_XR = OFF_QV + nm;                  // xr [n] (equMPC) | qT = T xr [n] (laxMPC)
    static constexpr int OFF_TMP = OFF_XR + n;                  // A Qi A' [n][n], B Ri B' [n][n] (factorisation only)
    static constexpr int STATE = OFF_TMP + 2 * n * n;
    static constexpr int STATE_VARB = STATE;
    static constexpr bool HAS_VARB = true;                      // LB_in / UB_in are part of the signature

    template <class A, bool VARB, class ST> struct Ctx {
        const spcies_consts *C;
        ST s;
        const BatchIO &io;
        __device__ Ctx(const spcies_consts *C_, ST s_, const BatchIO &io_) : C(C_), s(s_), io(io_) {}

        __device__ __forceinline__ real ab(int i, int j) const { return s.ld(OFF_AB + i * nm + j); }
        __device__ __forceinline__ real hi(int j) const { return s.ld(OFF_HI + j); }
        __device__ __forceinline__ real al(int h, int i, int j) const { return s.ld(OFF_ALPHA + (h * n + i) * n + j); }
        __device__ __forceinline__ real be(int h, int i, int j) const { return s.ld(OFF_BETA + (h * n + i) * n + j); }
        __device__ __forceinline__ void set_al(int h, int i, int j, real v) const { s.st(OFF_ALPHA + (h * n + i) * n + j, v); }
        __device__ __forceinline__ void set_be(int h, int i, int j, real v) const { s.st(OFF_BETA + (h * n + i) * n + j, v); }
        __device__ __forceinline__ real mu(int l, int j) const { return s.ld(OFF_MU + l * n + j); }
        __device__ __forceinline__ void set_mu(int l, int j, real v) const { s.st(OFF_MU + l * n + j, v); }
        // element of z / v / lambda / v1: the first m inputs, then stage l
        static __device__ __forceinline__ int e0(int j) { return j; }
        static __device__ __forceinline__ int el(int l, int j) { return m + l * nm + j; }

        // Beta_h (upper triangular, diagonal inverted) from `base`(i, j) [- Alpha_{h-1}' Alpha_{h-1}]; Q_rho_i is added to the diagonal
        // before the square root except in the last block                                             :155-176, :190-216, :234-255
        template <class Base> __device__ void beta_block(int h, bool with_alpha, bool with_qi, Base base, bool with_T = false) {
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
#if SPCIES_TERMINAL == 1
                    if (with_T) v = A::add(v, (real)C->T_rho_i[i][j]);      // dense: added to every entry   code_laxMPC_ADMM_C.c:258
#endif
                    if (i == j) {
                        if (with_qi) v = A::add(v, hi(i));
                        v = A::div(real(1), A::sqrt(v));
                    } else {
                        v = A::mul(v, be(h, i, i));
                    }
                    set_be(h, i, j, v);
                }
        }
        // Alpha_h = Beta_h^-T (-Q_rho_i A')                                                            :178-188, :218-230
        __device__ void alpha_block(int h) {
#pragma unroll 1
            for (int i = 0; i < n; ++i)
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    real v = A::mul(-hi(i), ab(j, i));
#pragma unroll 1
                    for (int l = 1; l <= i; ++l) v = A::nmsub(v, be(h, l - 1, i), al(h, l - 1, j));
                    set_al(h, i, j, A::mul(v, be(h, i, i)));
                }
        }

        __device__ void init(long long inst) {
            const int AQ = OFF_TMP, BR = OFF_TMP + n * n;
            const real rho_ = (real)rho;
            real x0[n], xr[n], ur[m];
#pragma unroll
            for (int i = 0; i < n; ++i) {
                x0[i] = (real)eng_x(C, io.x0, inst, n, i);
                xr[i] = (real)eng_x(C, io.xr, inst, n, i);
            }
#pragma unroll
            for (int i = 0; i < m; ++i) ur[i] = (real)eng_u(C, io.ur, inst, m, i);
            const double *Ain = io.ex[0] + inst * (long long)(n * n), *Bin = io.ex[1] + inst * (long long)(n * m);
            // bounds, weights, [A B], Hi                                                               :83-137
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
                s.st(OFF_HI + i, A::div(real(1), A::add(q, rho_)));
                for (int j = 0; j < n; ++j) s.st(OFF_AB + i * nm + j, (real)Ain[i + j * n]);
                for (int j = 0; j < m; ++j) s.st(OFF_AB + i * nm + n + j, (real)Bin[i + j * n]);
            }
#pragma unroll 1
            for (int j = 0; j < m; ++j) {
                const real r = (real)io.ex[3][inst * m + j];
                s.st(OFF_R + j, r);
                s.st(OFF_HI + n + j, A::div(real(1), A::add(r, rho_)));
            }
            // A Qi A', B Ri B'  (every term is (A_ik * Qi_k) * A_jk, accumulated in k order on zero-initialised arrays)   :142-151
#pragma unroll 1
            for (int i = 0; i < n; ++i)
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    real a = real(0), b = real(0);
#pragma unroll 1
                    for (int k = 0; k < n; ++k) a = A::add(a, A::mul(A::mul((real)Ain[i + k * n], hi(k)), (real)Ain[j + k * n]));
#pragma unroll 1
                    for (int k = 0; k < m; ++k) b = A::add(b, A::mul(A::mul((real)Bin[i + k * n], hi(n + k)), (real)Bin[j + k * n]));
                    s.st(AQ + i * n + j, a);
                    s.st(BR + i * n + j, b);
                }
#pragma unroll 4
            for (int e = 0; e < (2 * N - 1) * n * n; ++e) s.st(OFF_ALPHA + e, real(0));   // Alpha, Beta start as zeros   :55-56
            // the recursion
            beta_block(0, false, true, [&](int i, int j) { return s.ld(BR + i * n + j); });
            alpha_block(0);
#pragma unroll 1
            for (int h = 1; h < N - 1; ++h) {
                beta_block(h, true, true, [&](int i, int j) { return A::add(s.ld(AQ + i * n + j), s.ld(BR + i * n + j)); });
                alpha_block(h);
            }
            beta_block(N - 1, true, false, [&](int i, int j) { return A::add(s.ld(AQ + i * n + j), s.ld(BR + i * n + j)); }, LAX);
            // Q, R <- -Q, -R                                                                           :257-264
#pragma unroll
            for (int i = 0; i < n; ++i) s.st(OFF_Q + i, -s.ld(OFF_Q + i));
#pragma unroll
            for (int i = 0; i < m; ++i) s.st(OFF_R + i, -s.ld(OFF_R + i));
            // b, q                                                                                     :268-283
#pragma unroll 1
            for (int j = 0; j < n; ++j) {
                real b = real(0);
#pragma unroll
                for (int i = 0; i < n; ++i) b = A::sub(b, A::mul(ab(j, i), x0[i]));
                s.st(OFF_B + j, b);
                s.st(OFF_QV + j, A::mul(s.ld(OFF_Q + j), xr[j]));
#if SPCIES_TERMINAL == 1
                real qT = real(0);                                                               // qT = T xr (T dense, negated)   :292-295
#pragma unroll
                for (int i = 0; i < n; ++i) qT = A::add(qT, A::mul((real)C->T[j][i], xr[i]));
                s.st(OFF_XR + j, qT);
#else
                s.st(OFF_XR + j, xr[j]);
#endif
            }
#pragma unroll
            for (int j = 0; j < m; ++j) s.st(OFF_QV + n + j, A::mul(s.ld(OFF_R + j), ur[j]));
#pragma unroll 4
            for (int e = 0; e < 4 * ZLEN + N * n; ++e) s.st(OFF_V + e, real(0));        // v = lambda = z = v1 = 0, mu = 0
        }

        // one ADMM iteration                                                                            :291-553
        __device__ bool iterate(int) {
            const real rho_ = (real)rho, rhoi_ = (real)rho_i, tol_ = (real)tol;
            // v1 = v; q_hat -> z                                                                         :295-321
#pragma unroll 4
            for (int e = 0; e < ZLEN; ++e) {
                const real v = s.ld(OFF_V + e);
                s.st(OFF_V1 + e, v);
                const real q = e < m ? s.ld(OFF_QV + n + e) : (e < ZN ? s.ld(OFF_QV + (e - m) % nm) : s.ld(OFF_XR + e - ZN));
                s.st(OFF_Z + e, A::sub(A::add(q, s.ld(OFF_LAM + e)), A::mul(rho_, v)));
            }
            // r.h.s. of the W system                                                                     :326-353
#pragma unroll 1
            for (int j = 0; j < n; ++j) {
                real r = A::sub(A::mul(hi(j), s.ld(OFF_Z + el(0, j))), s.ld(OFF_B + j));
#pragma unroll 1
                for (int i = 0; i < m; ++i) r = A::sub(r, A::mul(A::mul(ab(j, i + n), hi(n + i)), s.ld(OFF_Z + e0(i))));
                set_mu(0, j, r);
            }
#pragma unroll 1
            for (int l = 1; l < N - 1; ++l)
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    real r = A::mul(hi(j), s.ld(OFF_Z + el(l, j)));
#pragma unroll 1
                    for (int i = 0; i < nm; ++i) r = A::sub(r, A::mul(A::mul(ab(j, i), hi(i)), s.ld(OFF_Z + el(l - 1, i))));
                    set_mu(l, j, r);
                }
#pragma unroll 1
            for (int j = 0; j < n; ++j) {
                real r = real(0);
#if SPCIES_TERMINAL == 1
#pragma unroll 1
                for (int i = 0; i < n; ++i) r = A::add(r, A::mul((real)C->T_rho_i[j][i], s.ld(OFF_Z + ZN + i)));      // Hi_N z_N   :373-381
#endif
#pragma unroll 1
                for (int i = 0; i < nm; ++i) r = A::sub(r, A::mul(A::mul(ab(j, i), hi(i)), s.ld(OFF_Z + el(N - 2, i))));
                set_mu(N - 1, j, LAX ? r : A::sub(r, s.ld(OFF_XR + j)));
            }
            // forward substitution                                                                       :358-388
#pragma unroll 1
            for (int l = 0; l < N; ++l)
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    real v = mu(l, j);
                    if (l > 0) {
#pragma unroll 1
                        for (int i = 0; i < n; ++i) v = A::sub(v, A::mul(al(l - 1, i, j), mu(l - 1, i)));
                    }
#pragma unroll 1
                    for (int i = 0; i < j; ++i) v = A::sub(v, A::mul(be(l, i, j), mu(l, i)));
                    set_mu(l, j, A::mul(be(l, j, j), v));
                }
            // backward substitution                                                                      :392-422
#pragma unroll 1
            for (int l = N - 1; l >= 0; --l)
#pragma unroll 1
                for (int j = n - 1; j >= 0; --j) {
                    real v = mu(l, j);
                    if (l < N - 1) {
#pragma unroll 1
                        for (int i = n - 1; i >= 0; --i) v = A::sub(v, A::mul(al(l, j, i), mu(l + 1, i)));
                    }
#pragma unroll 1
                    for (int i = n - 1; i > j; --i) v = A::sub(v, A::mul(be(l, j, i), mu(l, i)));
                    set_mu(l, j, A::mul(be(l, j, j), v));
                }
            // z                                                                                          :426-445
#pragma unroll 1
            for (int j = 0; j < m; ++j) {
                real z = s.ld(OFF_Z + e0(j));
#pragma unroll 1
                for (int i = 0; i < n; ++i) z = A::add(z, A::mul(ab(i, j + n), mu(0, i)));
                s.st(OFF_Z + e0(j), A::mul(-hi(n + j), z));
            }
#pragma unroll 1
            for (int l = 0; l < N - 1; ++l)
#pragma unroll 1
                for (int j = 0; j < nm; ++j) {
                    real z = s.ld(OFF_Z + el(l, j));
                    if (j < n) z = A::sub(z, mu(l, j));
#pragma unroll 1
                    for (int i = 0; i < n; ++i) z = A::add(z, A::mul(ab(i, j), mu(l + 1, i)));
                    s.st(OFF_Z + el(l, j), A::mul(-hi(j), z));
                }
#if SPCIES_TERMINAL == 1
            {   // z_N = -Hi_N (z_N - mu_{N-1})                                                            code_laxMPC_ADMM_C.c:476-485
                real aux[n];
#pragma unroll
                for (int j = 0; j < n; ++j) aux[j] = A::sub(s.ld(OFF_Z + ZN + j), mu(N - 1, j));
#pragma unroll 1
                for (int j = 0; j < n; ++j) {
                    real z = real(0);
#pragma unroll
                    for (int i = 0; i < n; ++i) z = A::sub(z, A::mul((real)C->T_rho_i[j][i], aux[i]));
                    s.st(OFF_Z + ZN + j, z);
                }
            }
#endif
            // v, lambda, residuals                                                                       :449-524
            bool over = false;
#pragma unroll 4
            for (int e = 0; e < ZLEN; ++e) {
                const int c = e < m ? n + e : (e < ZN ? (e - m) % nm : e - ZN);
                const real z = s.ld(OFF_Z + e), lam = s.ld(OFF_LAM + e);
                const real v = clip(A::add(z, A::mul(rhoi_, lam)), s.ld(OFF_LB + c), s.ld(OFF_UB + c));
                s.st(OFF_V + e, v);
                s.st(OFF_LAM + e, A::add(lam, A::mul(rho_, A::sub(z, v))));
                over = over || exceeds(A::sub(s.ld(OFF_V1 + e), v), tol_) || exceeds(A::sub(z, v), tol_);
            }
            return !over;
        }

        __device__ void finish(long long inst, int k, int ef) {
#pragma unroll
            for (int j = 0; j < m; ++j) io.u[inst * m + j] = eng_u_out(C, (double)s.ld(OFF_V + j), j);     // u_opt = v_0   :557-566
            io.k[inst] = k;
            io.e[inst] = ef;
            if (io.sol) {   // sol_<name>: z, v, lambda [N nm - n] (header_equMPC_ADMM_C.h), then the four times
                double *o = io.sol + inst * (long long)(sizeof(SPCIES_SOL_T) / sizeof(double));
                for (int e = 0; e < ZLEN; ++e) {
                    o[e] = (double)s.ld(OFF_Z + e);
                    o[ZLEN + e] = (double)s.ld(OFF_V + e);
                    o[2 * ZLEN + e] = (double)s.ld(OFF_LAM + e);
                }
                for (int e = 3 * ZLEN; e < (int)(sizeof(SPCIES_SOL_T) / sizeof(double)); ++e) o[e] = 0.0;
            }
        }
    };
};

typedef PolicyTraits<Solver> Traits;

}  // namespace admm_tv
}  // namespace spcies

#define SPCIES_TRAITS ::spcies::admm_tv::Traits
#include "spcies_entry.cuh"
