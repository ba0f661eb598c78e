// MPCT_ADMM_semiband.cuh -- batched ADMM solver for the MPC-for-tracking formulation whose equality-constrained QP step exploits
// the semi-banded structure of the problem (banded + low rank: the Woodbury identity applied to the Hessian and to the Schur
// complement of the equality constraints), hand-written for sm_100a.  Scalar rho, hard constraints, no constrained output (the
// defaults of def_options_MPCT_ADMM_semiband.m).
//
// Decision vector z = (x_0, u_0, ..., x_{N-1}, u_{N-1}, x_s, u_s), (N + 1) nm entries.  Per instance and iteration it performs
// exactly the arithmetic of formulations/+MPCT/code_MPCT_ADMM_semiband_C.c:123-1180 in the reference's order:
//   p = lambda - rho v (+ q on the last block, q = -(T xr, S ur))                                      :141-187
//   xi:  z1_a = blkdiag(Q_rho_i.., T_rho_i, S_rho_i) p;  z2_a = M_hat z1_a;  z3_a = blkdiag(..) (U_hat z2_a);  xi = z1_a - z3_a   :191-339
//   mu:  rhs = -(G xi + b);  z1_b = banded Cholesky solve;  z2_b = M_tilde z1_b;  z3_b = Cholesky solve of U_tilde z2_b;
//        mu = z1_b - z3_b                                                                              :342-515
//   z:   p <- -(G' mu + p);  z1_c = blkdiag(..) p;  z2_c = M_hat z1_c;  z3_c = blkdiag(..) (U_hat z2_c);  z = z1_c - z3_c   :518-770
//   v = clip(z + lambda / rho): x_0 to +-inf, (x_l, u_l) to [LB, UB], (x_s, u_s) to [LB + eps, UB - eps]    :775-830
//   lambda += rho (z - v);  exit on |v - v_prev| <= tol_d and |z - v| <= tol_p                          :1052-1121
// `solve_banded_Chol` (:1191-1268) and `solve_banded_QRST_sys` (:1271-1316) keep their loop order, so Arith<EXACT> is
// bit-identical.  The template re-uses `v` as scratch three times; here the scratch is its own array, which keeps the previous v
// alive for the fixed-point residual without a `v_old` copy.
#pragma once
// the generated header defines a macro named `inf` (cons_MPCT_ADMM_semiband_C.m:87): take its value before any system header sees it
#ifdef inf
static const double spcies_inf_value = inf;
#undef inf
#else
static const double spcies_inf_value = 1e6;
#endif
#include "spcies_dense_mma.cuh"

#if !defined(SCALAR_RHO) || SOFT_CONSTRAINTS != 0 || CONSTRAINED_OUTPUT != 0
#error "MPCT_ADMM_semiband.cuh implements scalar rho, hard constraints, no constrained output"
#endif

namespace spcies {
namespace mpct_sb {

constexpr int n = nn_, m = mm_, nm = nm_, N = NN_;
constexpr int L = (N + 1) * nm;      // decision vector
constexpr int LM = (N + 2) * n;      // multipliers of the equality constraints

struct Solver {
    typedef SPCIES_REAL real;
    static constexpr int OFF_Z = 0;
    static constexpr int OFF_V = OFF_Z + L;
    static constexpr int OFF_LAM = OFF_V + L;
    static constexpr int OFF_P = OFF_LAM + L;
    static constexpr int OFF_XI = OFF_P + L;
    static constexpr int OFF_Z3 = OFF_XI + L;
    static constexpr int OFF_W = OFF_Z3 + L;          // the template's three uses of `v` as scratch
    static constexpr int OFF_MU = OFF_W + L;          // [(N + 2) n]
    static constexpr int OFF_Z2 = OFF_MU + LM;        // [2 nm]
    static constexpr int OFF_Q = OFF_Z2 + 2 * nm;     // [nm]
    static constexpr int OFF_X0 = OFF_Q + nm;         // [n]
    static constexpr int STATE = OFF_X0 + n;
    static constexpr int STATE_VARB = STATE;
    static constexpr bool HAS_VARB = false;

    template <class A, bool VARB, class ST> struct Ctx {
        const spcies_consts *C;
        ST s;
        const BatchIO &io;
        __device__ Ctx(const spcies_consts *C_, ST s_, const BatchIO &io_) : C(C_), s(s_), io(io_) {}

        __device__ void init(long long inst) {
            real xr[n], ur[m];
#pragma unroll
            for (int i = 0; i < n; ++i) {
                s.st(OFF_X0 + i, (real)eng_x(C, io.x0, inst, n, i));
                xr[i] = (real)eng_x(C, io.xr, inst, n, i);
            }
#pragma unroll
            for (int i = 0; i < m; ++i) ur[i] = (real)eng_u(C, io.ur, inst, m, i);
#pragma unroll
            for (int i = 0; i < n; ++i) {                                  // q[i] -= T[i][j] xr[j]          :91-99
                real q = real(0);
#pragma unroll
                for (int j = 0; j < n; ++j) q = A::nmsub(q, C->T[i][j], xr[j]);
                s.st(OFF_Q + i, q);
            }
#pragma unroll
            for (int i = 0; i < m; ++i) {                                  // q[n + i] -= S[i][j] ur[j]      :101-109
                real q = real(0);
#pragma unroll
                for (int j = 0; j < m; ++j) q = A::nmsub(q, C->S[i][j], ur[j]);
                s.st(OFF_Q + n + i, q);
            }
#pragma unroll 4
            for (int e = 0; e < 3 * L; ++e) s.st(OFF_Z + e, real(0));      // z = v = lambda = 0
        }

        // out = blkdiag(Q_rho_i, R_rho_i, ..., T_rho_i, S_rho_i) d      (out starts at zero)                :1271-1316
        __device__ void qrst(int off_out, int off_d) const {
#pragma unroll 1
            for (int i = 0; i <= N; ++i) {
                const bool last = i == N;
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    real a = real(0);
#pragma unroll
                    for (int k = 0; k < n; ++k) a = A::madd(a, last ? C->T_rho_i[j][k] : C->Q_rho_i[j][k], s.ld(off_d + i * nm + k));
                    s.st(off_out + i * nm + j, a);
                }
#pragma unroll
                for (int j = 0; j < m; ++j) {
                    real a = real(0);
#pragma unroll
                    for (int k = 0; k < m; ++k) a = A::madd(a, last ? C->S_rho_i[j][k] : C->R_rho_i[j][k], s.ld(off_d + i * nm + n + k));
                    s.st(off_out + i * nm + n + j, a);
                }
            }
        }
        // z2 = M_hat x (sparse: the four distinct blocks of M_hat)                                          :196-283, :617-704
        __device__ void mhat(int off_x) const {
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    real a = real(0);
#pragma unroll 1
                    for (int l = 0; l < N; ++l)
#pragma unroll
                        for (int j = 0; j < n; ++j) a = A::madd(a, C->M_hat_x1[half * n + i][j], s.ld(off_x + l * nm + j));
#pragma unroll
                    for (int j = 0; j < n; ++j) a = A::madd(a, C->M_hat_x2[half * n + i][j], s.ld(off_x + N * nm + j));
                    s.st(OFF_Z2 + half * nm + i, a);
                }
#pragma unroll
                for (int i = 0; i < m; ++i) {
                    real a = real(0);
#pragma unroll 1
                    for (int l = 0; l < N; ++l)
#pragma unroll
                        for (int j = 0; j < m; ++j) a = A::madd(a, C->M_hat_u1[half * m + i][j], s.ld(off_x + l * nm + n + j));
#pragma unroll
                    for (int j = 0; j < m; ++j) a = A::madd(a, C->M_hat_u2[half * m + i][j], s.ld(off_x + N * nm + n + j));
                    s.st(OFF_Z2 + half * nm + n + i, a);
                }
            }
        }
        // w = U_hat z2 (sparse)                                                                             :296-325, :707-736
        __device__ void uhat() const {
#pragma unroll 1
            for (int l = 0; l < N; ++l) {
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    real a = real(0);
#pragma unroll
                    for (int j = 0; j < n; ++j) a = A::nmsub(a, C->Q[i][j], s.ld(OFF_Z2 + j));
                    s.st(OFF_W + l * nm + i, a);
                }
#pragma unroll
                for (int i = 0; i < m; ++i) {
                    real a = real(0);
#pragma unroll
                    for (int j = 0; j < m; ++j) a = A::nmsub(a, C->R[i][j], s.ld(OFF_Z2 + n + j));
                    s.st(OFF_W + l * nm + n + i, a);
                }
            }
#pragma unroll
            for (int i = 0; i < nm; ++i) s.st(OFF_W + N * nm + i, s.ld(OFF_Z2 + nm + i));
        }
        // in-place solve with the block-bidiagonal Cholesky factor of Gamma_tilde (N + 2 blocks)            :1191-1268
        __device__ void chol(int off) const {
#pragma unroll 1
            for (int k = 0; k < N + 2; ++k)
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    real d = s.ld(off + k * n + i);
#pragma unroll
                    for (int p_ = 0; p_ < n; ++p_)
                        if (p_ < i) d = A::nmsub(d, C->Beta[k][p_][i], s.ld(off + k * n + p_));
                    if (k > 0) {
#pragma unroll
                        for (int p_ = 0; p_ < n; ++p_) d = A::nmsub(d, C->Alpha[k - 1][p_][i], s.ld(off + (k - 1) * n + p_));
                    }
                    s.st(off + k * n + i, A::mul(d, C->Beta[k][i][i]));
                }
#pragma unroll 1
            for (int k = N + 1; k >= 0; --k)
#pragma unroll
                for (int i = n - 1; i >= 0; --i) {
                    real d = s.ld(off + k * n + i);
#pragma unroll
                    for (int p_ = 0; p_ < n; ++p_)
                        if (p_ > i) d = A::nmsub(d, C->Beta[k][i][p_], s.ld(off + k * n + p_));
                    if (k < N + 1) {
#pragma unroll
                        for (int p_ = 0; p_ < n; ++p_) d = A::nmsub(d, C->Alpha[k][i][p_], s.ld(off + (k + 1) * n + p_));
                    }
                    s.st(off + k * n + i, A::mul(d, C->Beta[k][i][i]));
                }
        }

        __device__ bool iterate(int /*k*/) {
            const real rho_ = (real)rho, rho_i_ = (real)rho_i;
            // p = lambda - rho v (+ q on the last block)                                                    :141-187
#pragma unroll 4
            for (int i = 0; i < L; ++i) s.st(OFF_P + i, A::sub(s.ld(OFF_LAM + i), A::mul(rho_, s.ld(OFF_V + i))));
#pragma unroll
            for (int i = 0; i < nm; ++i) s.st(OFF_P + N * nm + i, A::add(s.ld(OFF_P + N * nm + i), s.ld(OFF_Q + i)));
            // ---- xi (9a)                                                                                   :191-339
            qrst(OFF_XI, OFF_P);
            mhat(OFF_XI);
            uhat();
            qrst(OFF_Z3, OFF_W);
#pragma unroll 4
            for (int i = 0; i < L; ++i) s.st(OFF_XI + i, A::sub(s.ld(OFF_XI + i), s.ld(OFF_Z3 + i)));
            // ---- mu (9b): rhs = -(G xi + b)                                                                :342-393
#pragma unroll
            for (int i = 0; i < n; ++i) s.st(OFF_MU + i, -A::add(s.ld(OFF_X0 + i), s.ld(OFF_XI + i)));
#pragma unroll 1
            for (int l = 1; l <= N + 1; ++l) {
                const int src = (l <= N ? l - 1 : N) * nm;            // stage whose (x, u) enter through [A B]
                const int dst = (l <= N ? l : N) * nm;                // stage whose x enters with +1
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    real a = real(0);
#pragma unroll
                    for (int j = 0; j < n; ++j) a = A::nmsub(a, C->A[i][j], s.ld(OFF_XI + src + j));
                    a = A::add(a, s.ld(OFF_XI + dst + i));
#pragma unroll
                    for (int j = 0; j < m; ++j) a = A::nmsub(a, C->B[i][j], s.ld(OFF_XI + src + n + j));
                    s.st(OFF_MU + l * n + i, a);
                }
            }
            chol(OFF_MU);                                                                                   // z1_b   :395
            // z2_b = M_tilde z1_b (the shortened M_tilde of the scalar-rho case)                             :400-423
#pragma unroll 1
            for (int i = 0; i < 2 * nm; ++i) {
                real a = real(0);
#pragma unroll
                for (int j = 0; j < n; ++j) a = A::madd(a, C->M_tilde[i][j], s.ld(OFF_MU + j));
#pragma unroll 1
                for (int l = 0; l < N - 1; ++l)
#pragma unroll
                    for (int j = n; j < 2 * n; ++j) a = A::madd(a, C->M_tilde[i][j], s.ld(OFF_MU + l * n + j));
#pragma unroll
                for (int j = 2 * n; j < 4 * n; ++j) a = A::madd(a, C->M_tilde[i][j], s.ld(OFF_MU + (N - 2) * n + j));
                s.st(OFF_Z2 + i, a);
            }
            // U_tilde z2_b (the shortened U_tilde), then z3_b, mu = z1_b - z3_b                              :452-515
#pragma unroll 1
            for (int r = 0; r < LM; ++r) {
                const int j = r < n ? r : (r < N * n ? n + (r - n) % n : r - (N - 2) * n);     // row of the shortened U_tilde
                real a = real(0);
#pragma unroll
                for (int i = 0; i < 2 * nm; ++i) a = A::madd(a, C->U_tilde[j][i], s.ld(OFF_Z2 + i));
                s.st(OFF_W + r, a);
            }
            chol(OFF_W);
#pragma unroll 4
            for (int i = 0; i < LM; ++i) s.st(OFF_MU + i, A::sub(s.ld(OFF_MU + i), s.ld(OFF_W + i)));
            // ---- z (9c): p <- -(G' mu + p)                                                                 :518-609
#pragma unroll 1
            for (int l = 1; l <= N + 1; ++l) {
                // block l - 1 of p for l <= N: -(p - mu_{l-1}') ... ; the last block (x_s, u_s) for l = N + 1
                const bool last = l == N + 1;
                const int pb = (l - 1) * nm;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    real a;
                    if (l == 1) a = -A::add(s.ld(OFF_P + i), s.ld(OFF_MU + i));                              // :520-530
                    else if (!last) a = -A::sub(s.ld(OFF_P + pb + i), s.ld(OFF_MU + (l - 1) * n + i));         // :546-548
                    else a = A::add(-s.ld(OFF_P + pb + i), A::add(s.ld(OFF_MU + (N + 1) * n + i), s.ld(OFF_MU + N * n + i)));   // :574-576
#pragma unroll
                    for (int j = 0; j < n; ++j) a = A::nmsub(a, C->A[j][i], s.ld(OFF_MU + l * n + j));
                    s.st(OFF_P + pb + i, a);
                }
#pragma unroll
                for (int i = 0; i < m; ++i) {
                    real a = s.ld(OFF_P + pb + n + i);
#pragma unroll
                    for (int j = 0; j < n; ++j) a = A::madd(a, C->B[j][i], s.ld(OFF_MU + l * n + j));
                    s.st(OFF_P + pb + n + i, -a);
                }
            }
            qrst(OFF_Z, OFF_P);                                                                             // z1_c   :613
            mhat(OFF_Z);
            uhat();
            qrst(OFF_Z3, OFF_W);                                                                            // z3_c   :760
            // z, v, lambda, residuals                                                                        :763-830, :1052-1121
            bool over = false;
#pragma unroll 1
            for (int l = 0; l <= N; ++l)
#pragma unroll
                for (int c = 0; c < nm; ++c) {
                    const int i = l * nm + c;
                    const real z = A::sub(s.ld(OFF_Z + i), s.ld(OFF_Z3 + i));
                    const real lam = s.ld(OFF_LAM + i), vo = s.ld(OFF_V + i);
                    real v = A::add(A::mul(rho_i_, lam), z);
                    real lo, hi;
                    if (l == 0 && c < n) {
                        lo = -(real)spcies_inf_value;
                        hi = (real)spcies_inf_value;
                    } else if (l < N) {
                        lo = C->LB[c];
                        hi = C->UB[c];
                    } else {
                        const real eps = c < n ? (real)eps_x : (real)eps_u;
                        lo = A::add(C->LB[c], eps);
                        hi = A::sub(C->UB[c], eps);
                    }
                    v = (v > lo) ? v : lo;
                    v = (v < hi) ? v : hi;
                    s.st(OFF_Z + i, z);
                    s.st(OFF_V + i, v);
                    s.st(OFF_LAM + i, A::madd(lam, rho_, A::sub(z, v)));
                    over |= exceeds(A::sub(v, vo), (real)tol_d) || exceeds(A::sub(z, v), (real)tol_p);
                }
            return !over;
        }

        __device__ void finish(long long inst, int k, int ef) {
#pragma unroll
            for (int j = 0; j < m; ++j) io.u[inst * m + j] = eng_u_out(C, (double)s.ld(OFF_V + n + j), j);   // u_opt = v[n + j]   :1140-1149
            io.k[inst] = k;
            io.e[inst] = ef;
            if (io.sol) {   // sol_<name>: z, v, lambda (header_MPCT_ADMM_semiband_C.h)
                double *o = io.sol + inst * (long long)(sizeof(SPCIES_SOL_T) / sizeof(double));
                for (int e = 0; e < 3 * L; ++e) o[e] = (double)s.ld(OFF_Z + e);
                for (int e = 3 * L; e < (int)(sizeof(SPCIES_SOL_T) / sizeof(double)); ++e) o[e] = 0.0;
            }
        }
    };
};

#include "MPCT_ADMM_semiband_mma.cuh"

typedef dense::DenseTraits<Solver, Engine> Traits;

}  // namespace mpct_sb
}  // namespace spcies

#define SPCIES_TRAITS ::spcies::mpct_sb::Traits
#include "spcies_entry.cuh"
