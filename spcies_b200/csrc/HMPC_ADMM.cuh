// HMPC_ADMM.cuh -- batched ADMM / symmetric-ADMM solver for harmonic MPC *without* the (z_hat, s_hat) splitting: the solver a
// user gets for formulation 'HMPC' with the toolbox defaults (method 'ADMM', submethod '', classes/Spcies_options.m:94,104 ->
// cons_HMPC_ADMM_C.m), and -- with three state / input references -- the ellipHMPC solver (SPCIES_NREF == 3,
// code_ellipHMPC_ADMM_C.c:18).  Hand-written for sm_100a.
//
// Per instance it performs exactly the arithmetic of formulations/+HMPC/code_HMPC_ADMM_C.c:83-279:
//   b = -A x0;  q gets -Te xr - QQ x0 (x_e), -QQ x0 (x_c), -Se ur (u_e)   [ellipHMPC: + -Th x_rs, -Th x_rc, -Sh u_rs, -Sh u_rc]
//   w = rho (s [- d]) + lambda;  q_hat = q + C' w                       CSR mat-vec with C'            :122-135
//   z = M2 b + M1 q_hat                                                 dense mat-vec (the hot loop)   :143-155
//   Cz = C z [- d]                                                      CSR mat-vec                    :159-169
//   [SADMM] lambda += alpha rho (Cz + s)                                                               :173-180
//   s = -Cz - lambda / rho;  box clip of the first n_box entries;  cone entries: diamond sets
//       proj_SOC3(+1, LBy) then proj_SOC3(-1, UBy)  [USE_SOC: proj_SOC3(+1, 0)]                        :184-206
//   Cz += s;  lambda += [alpha] rho Cz                                                                 :209-229
//   exit on |Cz| <= tol_p and |s - s_prev| <= tol_d                                                    :233-261
// q has 3 (n + m) trailing entries that can be non-zero and b has n: both are kept compactly (adding the remaining exact zeros
// is the identity in IEEE arithmetic).  `s_ant` is not stored: the dual residual is taken when s is overwritten.
#pragma once
#include "spcies_dense_mma.cuh"
#include "spcies_sparse.cuh"

#ifndef SPCIES_NREF
#define SPCIES_NREF 1
#endif

namespace spcies {
namespace hmpc_ns {

constexpr int n = nn_, m = mm_, nm = nm_, N = NN_;
constexpr int DIM = dim, NS = n_s, NBOX = n_box, NY = n_y;
constexpr int Q0 = (N - 1) * nm + m;   // first index of x_e in z: the tail z[Q0..DIM) = (x_e, x_s, x_c, u_e, u_s, u_c)
constexpr int NQ = 3 * nm;
static_assert(Q0 + NQ == DIM, "decision vector layout");
#ifdef IS_SYMMETRIC
constexpr bool SYMMETRIC = true;
#define SPCIES_ALPHA alpha_SADMM
#else
constexpr bool SYMMETRIC = false;
#define SPCIES_ALPHA 1.0
#endif
#ifdef USE_SOC
constexpr bool SOC = true;
constexpr int NCONE = n_soc;
#else
constexpr bool SOC = false;
constexpr int NCONE = n_y;
#endif
static_assert(NBOX + 3 * NCONE == NS, "s = (box part, cone triples)");

struct Solver {
    typedef SPCIES_REAL real;
    static constexpr int OFF_Z = 0;
    static constexpr int OFF_S = OFF_Z + DIM;
    static constexpr int OFF_LAM = OFF_S + NS;
    static constexpr int OFF_QH = OFF_LAM + NS;
    static constexpr int OFF_CZ = OFF_QH + DIM;
    static constexpr int OFF_Q = OFF_CZ + NS;        // q[Q0 .. DIM)
    static constexpr int OFF_B = OFF_Q + NQ;         // b[0 .. n)
    static constexpr int STATE = OFF_B + n;
    static constexpr int STATE_VARB = STATE;
    static constexpr bool HAS_VARB = false;

    template <class A, bool VARB, class ST> struct Ctx {
        const spcies_consts *C;
        ST s;
        const BatchIO &io;
        __device__ Ctx(const spcies_consts *C_, ST s_, const BatchIO &io_) : C(C_), s(s_), io(io_) {}

        __device__ __forceinline__ real d_at(int i) const {
#ifdef USE_SOC
            return C->d[i];
#else
            (void)i;
            return real(0);
#endif
        }

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
            for (int j = 0; j < NQ; ++j) s.st(OFF_Q + j, real(0));
#pragma unroll
            for (int j = 0; j < n; ++j) {
                real b = real(0), qe = real(0), qc = real(0);
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    b = A::nmsub(b, C->A[j][i], x0[i]);                                              // :85-90
                    qe = A::sub(qe, A::madd(A::mul(C->Te[j][i], xr[i]), C->QQ[j][i], x0[i]));        // :93-97
#if SPCIES_NREF == 3
                    qc = A::sub(qc, A::madd(A::mul(C->Th[j][i], (real)io.ex[1][inst * n + i]), C->QQ[j][i], x0[i]));   // ellip :110-114
#else
                    qc = A::nmsub(qc, C->QQ[j][i], x0[i]);                                           // :98-102
#endif
                }
                s.st(OFF_B + j, b);
                s.st(OFF_Q + j, qe);
                s.st(OFF_Q + 2 * n + j, qc);
#if SPCIES_NREF == 3
                real qs = real(0);
#pragma unroll
                for (int i = 0; i < n; ++i) qs = A::nmsub(qs, C->Th[j][i], (real)io.ex[0][inst * n + i]);              // ellip :105-109
                s.st(OFF_Q + n + j, qs);
#endif
            }
#pragma unroll
            for (int j = 0; j < m; ++j) {
                real qu = real(0);
#pragma unroll
                for (int i = 0; i < m; ++i) qu = A::nmsub(qu, C->Se[j][i], ur[i]);                   // :103-107
                s.st(OFF_Q + 3 * n + j, qu);
#if SPCIES_NREF == 3
                real q2 = real(0), q3 = real(0);
#pragma unroll
                for (int i = 0; i < m; ++i) {
                    q2 = A::nmsub(q2, C->Sh[j][i], (real)io.ex[2][inst * m + i]);                      // ellip :121-125
                    q3 = A::nmsub(q3, C->Sh[j][i], (real)io.ex[3][inst * m + i]);                      // ellip :126-130
                }
                s.st(OFF_Q + 3 * n + m + j, q2);
                s.st(OFF_Q + 3 * n + 2 * m + j, q3);
#endif
            }
#pragma unroll 4
            for (int e = 0; e < 2 * NS; ++e) s.st(OFF_S + e, real(0));     // s = lambda = 0
        }

        __device__ bool iterate(int /*k*/) {
            const real rho_ = C->rho, rho_i_ = C->rho_i;
            // w = rho (s - d) + lambda (kept in Cz)                                                          :125-131
#pragma unroll 2
            for (int i = 0; i < NS; ++i) {
                const real sv = SOC ? A::sub(s.ld(OFF_S + i), d_at(i)) : s.ld(OFF_S + i);
                s.st(OFF_CZ + i, A::add(A::mul(rho_, sv), s.ld(OFF_LAM + i)));
            }
            // q_hat = q + C' w                                                                               :132-137
            {
                int j = C->Ct_row[0];
#pragma unroll 1
                for (int i = 0; i < DIM; ++i) {
                    const int jend = C->Ct_row[i + 1];
                    real acc = i >= Q0 ? s.ld(OFF_Q + i - Q0) : real(0);
#pragma unroll 1
                    for (; j < jend; ++j) acc = A::madd(acc, C->Ct_val[j], s.ld(OFF_CZ + C->Ct_col[j]));
                    s.st(OFF_QH + i, acc);
                }
            }
            // z = M2 b + M1 q_hat, four rows at a time (every row accumulates in the reference's order)      :143-155
            real b[n];
#pragma unroll
            for (int j = 0; j < n; ++j) b[j] = s.ld(OFF_B + j);
            constexpr int RB = 4;
#pragma unroll 1
            for (int i0 = 0; i0 < DIM; i0 += RB) {
                real acc[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    acc[r] = real(0);
                    const int i = (i0 + r < DIM) ? i0 + r : DIM - 1;
#pragma unroll
                    for (int j = 0; j < n; ++j) acc[r] = A::madd(acc[r], C->M2[i][j], b[j]);
                }
#pragma unroll 2
                for (int j = 0; j < DIM; ++j) {
                    const real qh = s.ld(OFF_QH + j);
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        const int i = (i0 + r < DIM) ? i0 + r : DIM - 1;
                        acc[r] = A::madd(acc[r], C->M1[i][j], qh);
                    }
                }
#pragma unroll
                for (int r = 0; r < RB; ++r)
                    if (i0 + r < DIM) s.st(OFF_Z + i0 + r, acc[r]);
            }
            // Cz = C z - d                                                                                   :159-169
            {
                int j = C->C_row[0];
#pragma unroll 1
                for (int i = 0; i < NS; ++i) {
                    const int jend = C->C_row[i + 1];
                    real acc = SOC ? -d_at(i) : real(0);
#pragma unroll 1
                    for (; j < jend; ++j) acc = A::madd(acc, C->C_val[j], s.ld(OFF_Z + C->C_col[j]));
                    s.st(OFF_CZ + i, acc);
                }
            }
            bool over = false;
            const real ar = SYMMETRIC ? A::mul((real)SPCIES_ALPHA, rho_) : rho_;          // alpha_SADMM*rho
            // box part                                                                      :173-194, :209-261
#pragma unroll 1
            for (int j = 0; j < NBOX; ++j) {
                const real cz = s.ld(OFF_CZ + j), so = s.ld(OFF_S + j);
                real lam = s.ld(OFF_LAM + j);
                if (SYMMETRIC) lam = A::madd(lam, ar, A::add(cz, so));
                real sv = A::sub(-cz, A::mul(rho_i_, lam));
                sv = clip(sv, C->LB[j], C->UB[j]);
                const real cz2 = A::add(cz, sv);
                s.st(OFF_S + j, sv);
                s.st(OFF_LAM + j, A::madd(lam, ar, cz2));
                over |= exceeds(cz2, (real)tol_p) || exceeds(A::sub(sv, so), (real)tol_d);
            }
            // cone part: one (y_e, y_s, y_c) triple at a time                               :197-206
#pragma unroll 1
            for (int g = 0; g < NCONE; ++g) {
                real cz[3], so[3], lam[3], sv[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int j = NBOX + 3 * g + c;
                    cz[c] = s.ld(OFF_CZ + j);
                    so[c] = s.ld(OFF_S + j);
                    lam[c] = s.ld(OFF_LAM + j);
                    if (SYMMETRIC) lam[c] = A::madd(lam[c], ar, A::add(cz[c], so[c]));
                    sv[c] = A::sub(-cz[c], A::mul(rho_i_, lam[c]));
                }
                if (SOC) {
                    proj_soc3<A>(sv, real(1), real(0));
                } else {
                    proj_soc3<A>(sv, real(1), C->LBy[g]);
                    proj_soc3<A>(sv, real(-1), C->UBy[g]);
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int j = NBOX + 3 * g + c;
                    const real cz2 = A::add(cz[c], sv[c]);
                    s.st(OFF_S + j, sv[c]);
                    s.st(OFF_LAM + j, A::madd(lam[c], ar, cz2));
                    over |= exceeds(cz2, (real)tol_p) || exceeds(A::sub(sv[c], so[c]), (real)tol_d);
                }
            }
            return !over;
        }

        __device__ void finish(long long inst, int k, int ef) {
#pragma unroll
            for (int j = 0; j < m; ++j) io.u[inst * m + j] = eng_u_out(C, (double)s.ld(OFF_Z + j), j);   // u_opt = z[0..m)   :271-282
            io.k[inst] = k;
            io.e[inst] = ef;
            if (io.sol) {   // sol_<name>: z, s, lambda (header_HMPC_ADMM_C.h)
                double *o = io.sol + inst * (long long)(sizeof(SPCIES_SOL_T) / sizeof(double));
                for (int e = 0; e < DIM + 2 * NS; ++e) o[e] = (double)s.ld(OFF_Z + e);
                for (int e = DIM + 2 * NS; e < (int)(sizeof(SPCIES_SOL_T) / sizeof(double)); ++e) o[e] = 0.0;
            }
        }
    };
};

#include "HMPC_ADMM_mma.cuh"

typedef dense::DenseTraits<Solver, Engine> Traits;

}  // namespace hmpc_ns
}  // namespace spcies

#define SPCIES_TRAITS ::spcies::hmpc_ns::Traits
#include "spcies_entry.cuh"
