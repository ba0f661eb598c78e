// HMPC_ADMM_split.cuh -- batched ADMM / symmetric-ADMM solver for harmonic MPC with the (z_hat, s_hat) = (z, s)
// splitting, hand-written for sm_100a.  Dense path (NON_SPARSE: M1, M2), box constraints, diamond-set projections.
//
// Per instance it performs exactly the arithmetic of formulations/+HMPC/code_HMPC_ADMM_split_C.c:99-347:
//   bh[0:n] = -A x0;  q gets -Te xr - QQ x0, -QQ x0, -Se ur                                   :99-129
//   q_hat = [sigma z - q - lambda ; rho s - mu]                                               :149-154
//   primal_hat = M2 bh - M1 q_hat                        dense mat-vec (the hot loop)         :176-188
//   [SADMM] dual += alpha (sigma|rho) (primal_hat - primal)                                   :215-225
//   z = clip(z_hat + lambda/sigma) on the first dim-3nm entries;  s = s_hat + mu/rho          :230-245
//   s_j <- proj_D(s_j; LBy_j, UBy_j) = proj_SOC3(+1, LBy) then proj_SOC3(-1, UBy)             :255-258
//   dual += [alpha] (sigma|rho) (primal_hat - primal)                                         :288-312
//   exit on |primal_prev - primal| <= tol_d and |primal - primal_hat| <= tol_p                :318-334
// q has 2n+m and bh has n non-zero entries; they are kept compactly (adding / subtracting the remaining exact zeros is
// the identity in IEEE arithmetic, so this is bit-exact).  The M1 product is blocked four rows at a time: each q_hat
// element is loaded once per block and every row still accumulates in the reference's j order.
//
// This is the thread-per-instance formulation (state in shared memory or, for N = 50, in the L2-resident scratch; M1
// read through the read-only path): the EXACT / debug-payload path.  The FAST arithmetic runs the batched product
// [M1] x [q_hat of a tile of instances] as a GEMM on the dense tensor-core engine (HMPC_ADMM_split_mma.cuh).
#pragma once
#include <type_traits>
#include "spcies_dense_mma.cuh"
#include "spcies_sparse.cuh"

// NON_SPARSE (default): dense M1 / M2.  Otherwise (solver option sparse = true): the KKT system [Hh, Gh'; Gh, 0] is solved per
// iteration through its L D L' factorisation in CSC form (code_HMPC_ADMM_split_C.c:161-164, :193-209; ldl_solve_csc).
#if defined(COUPLED_CONSTRAINTS) || defined(USE_SOC)
#error "HMPC_ADMM_split.cuh implements box constraints with diamond-set projections (no COUPLED_CONSTRAINTS / USE_SOC)"
#endif

namespace spcies {
namespace hmpc {

constexpr int n = nn_, m = mm_, nm = nm_, N = NN_;
constexpr int DIM = dim, NS = n_s;
constexpr int NP = DIM + NS;
constexpr int Q0 = (N - 1) * nm + m;   // first index of x_e in z
#ifdef IS_SYMMETRIC
constexpr bool SYMMETRIC = true;
#define SPCIES_ALPHA alpha_SADMM
#else
constexpr bool SYMMETRIC = false;
#define SPCIES_ALPHA 1.0
#endif

struct Solver {
    typedef SPCIES_REAL real;
    static constexpr int OFF_P = 0;              // primal = (z, s)
    static constexpr int OFF_D = OFF_P + NP;     // dual = (lambda, mu)
#ifdef NON_SPARSE
    static constexpr int OFF_PH = OFF_D + NP;    // primal_hat
    static constexpr int OFF_QH = OFF_PH + NP;   // q_hat
    static constexpr int OFF_QE = OFF_QH + NP;   // q at x_e  [n]
#else
    static constexpr int OFF_PH = OFF_D + NP;    // rhs[nrow_M] = (q_hat, bh) -> (primal_hat, multipliers): z_hat, s_hat are its head
    static constexpr int OFF_QH = OFF_PH;
    static constexpr int OFF_QE = OFF_PH + nrow_M;
#endif
    static constexpr int OFF_QC = OFF_QE + n;    // q at x_c  [n]
    static constexpr int OFF_QU = OFF_QC + n;    // q at u_e  [m]
    static constexpr int OFF_BH = OFF_QU + m;    // bh[0:n]
    static constexpr int STATE = OFF_BH + n;
    static constexpr int STATE_VARB = STATE;
    static constexpr bool HAS_VARB = false;

    template <class A, bool VARB, class ST> struct Ctx {
        const spcies_consts *C;
        ST s;
        const BatchIO &io;
        __device__ Ctx(const spcies_consts *C_, ST s_, const BatchIO &io_) : C(C_), s(s_), io(io_) {}

        __device__ void init(long long inst) {
            real x0[n], xr[n], ur[m];
#pragma unroll
            for (int i = 0; i < n; ++i) {
                x0[i] = (real)eng_x(C, io.x0, inst, n, i);
                xr[i] = (real)eng_x(C, io.xr, inst, n, i);
            }
#pragma unroll
            for (int i = 0; i < m; ++i) ur[i] = (real)eng_u(C, io.ur, inst, m, i);
#pragma unroll
            for (int j = 0; j < n; ++j) {
                real b = real(0), qe = real(0), qc = real(0);
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    b = A::nmsub(b, C->A[j][i], x0[i]);                                              // :99-104
                    qe = A::sub(qe, A::madd(A::mul(C->Te[j][i], xr[i]), C->QQ[j][i], x0[i]));        // :115-119
                    qc = A::nmsub(qc, C->QQ[j][i], x0[i]);                                           // :120-124
                }
                s.st(OFF_BH + j, b);
                s.st(OFF_QE + j, qe);
                s.st(OFF_QC + j, qc);
            }
#pragma unroll
            for (int j = 0; j < m; ++j) {
                real qu = real(0);
#pragma unroll
                for (int i = 0; i < m; ++i) qu = A::nmsub(qu, C->Se[j][i], ur[i]);                   // :125-129
                s.st(OFF_QU + j, qu);
            }
#pragma unroll 4
            for (int e = 0; e < 2 * NP; ++e) s.st(OFF_P + e, real(0));
        }

        __device__ __forceinline__ real q_at(int idx) const {
            if (idx < Q0) return real(0);
            const int t = idx - Q0;
            if (t < n) return s.ld(OFF_QE + t);
            if (t >= 2 * n && t < 3 * n) return s.ld(OFF_QC + t - 2 * n);
            if (t >= 3 * n && t < 3 * n + m) return s.ld(OFF_QU + t - 3 * n);
            return real(0);
        }

        __device__ bool iterate(int /*k*/) {
            const real sigma_ = C->sigma, sigma_i_ = C->sigma_i, rho_ = C->rho, rho_i_ = C->rho_i;
            // q_hat                                                                                  :149-154
#pragma unroll 1
            for (int j = 0; j < DIM; ++j)
                s.st(OFF_QH + j, A::sub(A::sub(A::mul(sigma_, s.ld(OFF_P + j)), q_at(j)), s.ld(OFF_D + j)));
#pragma unroll 1
            for (int j = 0; j < NS; ++j)
                s.st(OFF_QH + DIM + j, A::sub(A::mul(rho_, s.ld(OFF_P + DIM + j)), s.ld(OFF_D + DIM + j)));
#ifndef NON_SPARSE
            // rhs = (q_hat, bh) with bh[idx_x0[j]] = -A x0;  L D L' rhs = rhs                          :161-164, :193-209
#pragma unroll 1
            for (int j = 0; j < nrow_M - NP; ++j) s.st(OFF_PH + NP + j, C->bh[j]);
#pragma unroll
            for (int j = 0; j < n; ++j) s.st(OFF_PH + NP + C->idx_x0[j], s.ld(OFF_BH + j));
            ldl_solve_csc<A>(s, OFF_PH, (int)nrow_M, C->L_val, C->L_row, C->L_col, C->Dinv);
#else
            // primal_hat = M2 bh - M1 q_hat                                                          :176-188
            real bh[n];
#pragma unroll
            for (int j = 0; j < n; ++j) bh[j] = s.ld(OFF_BH + j);
            constexpr int RB = 4;
#pragma unroll 1
            for (int i0 = 0; i0 < NP; i0 += RB) {
                real acc[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    acc[r] = real(0);
                    const int i = (i0 + r < NP) ? i0 + r : NP - 1;
#pragma unroll
                    for (int j = 0; j < n; ++j) acc[r] = A::madd(acc[r], C->M2[i][j], bh[j]);
                }
#pragma unroll 2
                for (int j = 0; j < NP; ++j) {
                    const real qh = s.ld(OFF_QH + j);
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        const int i = (i0 + r < NP) ? i0 + r : NP - 1;
                        acc[r] = A::nmsub(acc[r], C->M1[i][j], qh);
                    }
                }
#pragma unroll
                for (int r = 0; r < RB; ++r)
                    if (i0 + r < NP) s.st(OFF_PH + i0 + r, acc[r]);
            }
#endif

            bool over = false;
            const real as = SYMMETRIC ? A::mul((real)SPCIES_ALPHA, sigma_) : sigma_;   // alpha_SADMM*sigma
            const real ar = SYMMETRIC ? A::mul((real)SPCIES_ALPHA, rho_) : rho_;
            // z, lambda                                                                 :215-219, :230-238, :288-305, :318-334
#pragma unroll 1
            for (int j = 0; j < DIM; ++j) {
                const real zh = s.ld(OFF_PH + j), zo = s.ld(OFF_P + j);
                real lam = s.ld(OFF_D + j);
                if (SYMMETRIC) lam = A::madd(lam, as, A::sub(zh, zo));
                real z = A::madd(zh, sigma_i_, lam);
                if (j < DIM - 3 * n - 3 * m) z = clip(z, C->LB[j], C->UB[j]);
                s.st(OFF_P + j, z);
                s.st(OFF_D + j, A::madd(lam, as, A::sub(zh, z)));
                over |= exceeds(A::sub(zo, z), (real)tol_d) || exceeds(A::sub(z, zh), (real)tol_p);
            }
            // s, mu: diamond-set projection of each (y_e, y_s, y_c) triple                              :220-225, :241-258, :294-311
#pragma unroll 1
            for (int g = 0; g < nm; ++g) {
                real sv[3], sh[3], so[3], mu[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    sh[c] = s.ld(OFF_PH + DIM + 3 * g + c);
                    so[c] = s.ld(OFF_P + DIM + 3 * g + c);
                    mu[c] = s.ld(OFF_D + DIM + 3 * g + c);
                    if (SYMMETRIC) mu[c] = A::madd(mu[c], ar, A::sub(sh[c], so[c]));
                    sv[c] = A::madd(sh[c], rho_i_, mu[c]);
                }
                proj_soc3<A>(sv, real(1), C->LBy[g]);
                proj_soc3<A>(sv, real(-1), C->UBy[g]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    s.st(OFF_P + DIM + 3 * g + c, sv[c]);
                    s.st(OFF_D + DIM + 3 * g + c, A::madd(mu[c], ar, A::sub(sh[c], sv[c])));
                    over |= exceeds(A::sub(so[c], sv[c]), (real)tol_d) || exceeds(A::sub(sv[c], sh[c]), (real)tol_p);
                }
            }
            return !over;
        }

        __device__ void finish(long long inst, int k, int ef) {
#pragma unroll
            for (int j = 0; j < m; ++j) io.u[inst * m + j] = eng_u_out(C, (double)s.ld(OFF_P + j), j);   // u_opt = z[0..m)   :359-368
            io.k[inst] = k;
            io.e[inst] = ef;
            if (io.sol) {   // sol_<name>: z, s, z_hat, s_hat, lambda, mu (header_HMPC_ADMM_split_C.h)
                double *o = io.sol + inst * (long long)(sizeof(SPCIES_SOL_T) / sizeof(double));
                for (int e = 0; e < NP; ++e) {
                    o[e] = (double)s.ld(OFF_P + e);
                    o[NP + e] = (double)s.ld(OFF_PH + e);
                    o[2 * NP + e] = (double)s.ld(OFF_D + e);
                }
                for (int e = 3 * NP; e < (int)(sizeof(SPCIES_SOL_T) / sizeof(double)); ++e) o[e] = 0.0;
            }
        }
    };
};

#include "HMPC_ADMM_split_mma.cuh"

typedef dense::DenseTraits<Solver, Engine> Traits;

}  // namespace hmpc
}  // namespace spcies

#define SPCIES_TRAITS ::spcies::hmpc::Traits
#include "spcies_entry.cuh"
