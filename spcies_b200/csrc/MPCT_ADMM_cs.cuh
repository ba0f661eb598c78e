// MPCT_ADMM_cs.cuh -- batched ADMM solver for the MPC-for-tracking formulation on the extended state space
// (z_j = (x_j, x_s), v_j = (u_j, u_s); the default submethod of MPCT / ADMM, classes/Spcies_options.m:103), hand-written for
// sm_100a.
//
// Per instance it performs exactly the arithmetic of formulations/+MPCT/code_MPCT_ADMM_cs_C.c:57-267:
//   q = (0, Tz xr, 0, Sz ur) per stage (Tz = -T/N, Sz = -S/N)                         :73-82
//   q_hat = q + lambda - rho v                                                          :101-107
//   rhs = (-A Hi) q_hat - b                  CSR mat-vec, b = x0 on the first n rows    :111-119
//   W mu = rhs                               CSC L D L' solve (QDLDL style)             :124-147
//   z = (-Hi) q_hat + (-Hi A') mu            2 x CSR mat-vec                            :153-165
//   v = clip(z + lambda / rho);  lambda += rho (z - v)                                  :169-188
//   exit on |v_prev - v| <= tol and |z - v| <= tol                                      :192-216
// The sparse structure is the generator's (full2CSR / full2CSC / full2LDL, spcies_b200/sp_utils.py); the per-instance vectors
// are dynamically indexed in the [element][thread] state, the operation order is the reference's, so Arith<EXACT> is
// bit-identical.  `v1` (the previous v) is not stored: the fixed-point residual is taken when v is overwritten.
#pragma once
#include "spcies_kernel.cuh"
#include "spcies_dense_mma.cuh"
#include "spcies_sparse.cuh"

namespace spcies {
namespace mpct_cs {

constexpr int n = nn_, m = mm_, nm = nm_, N = NN_;
constexpr int DNM = dnm_;            // 2 (n + m): one extended stage
constexpr int DIM = N * DNM;         // decision vector
constexpr int NR = nrow_AHi;         // rows of the W system

struct Solver {
    typedef SPCIES_REAL real;
    static constexpr int OFF_Z = 0;
    static constexpr int OFF_V = OFF_Z + DIM;
    static constexpr int OFF_LAM = OFF_V + DIM;
    static constexpr int OFF_QH = OFF_LAM + DIM;
    static constexpr int OFF_MU = OFF_QH + DIM;       // rhs / mu [nrow_AHi]
    static constexpr int OFF_Q = OFF_MU + NR;         // q [dnm_]
    static constexpr int OFF_B = OFF_Q + DNM;         // b [n]
    static constexpr int STATE = OFF_B + n;
    static constexpr int STATE_VARB = STATE;
    static constexpr bool HAS_VARB = false;

    template <class A, bool VARB, class ST> struct Ctx {
        const spcies_consts *C;
        ST s;
        const BatchIO &io;
        __device__ Ctx(const spcies_consts *C_, ST s_, const BatchIO &io_) : C(C_), s(s_), io(io_) {}

        __device__ __forceinline__ real rho_at(int j) const {
#ifdef SCALAR_RHO
            return C->rho;
#else
            return C->rho[j];
#endif
        }
        __device__ __forceinline__ real rho_i_at(int j) const {
#ifdef SCALAR_RHO
            return C->rho_i;
#else
            return C->rho_i[j];
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
            for (int j = 0; j < DNM; ++j) s.st(OFF_Q + j, real(0));
#pragma unroll
            for (int j = 0; j < n; ++j) {
                real q = real(0);
#pragma unroll
                for (int i = 0; i < n; ++i) q = A::madd(q, C->Tz[j][i], xr[i]);
                s.st(OFF_Q + n + j, q);
                s.st(OFF_B + j, x0[j]);
            }
#pragma unroll
            for (int j = 0; j < m; ++j) {
                real q = real(0);
#pragma unroll
                for (int i = 0; i < m; ++i) q = A::madd(q, C->Sz[j][i], ur[i]);
                s.st(OFF_Q + 2 * n + m + j, q);
            }
#pragma unroll 4
            for (int e = 0; e < 3 * DIM; ++e) s.st(OFF_Z + e, real(0));   // z = v = lambda = 0
        }

        __device__ bool iterate(int /*k*/) {
            // q_hat = q + lambda - rho v                                                                 :101-107
#pragma unroll 1
            for (int l = 0; l < N; ++l)
#pragma unroll 4
                for (int i = 0; i < DNM; ++i) {
                    const int j = l * DNM + i;
                    s.st(OFF_QH + j, A::nmsub(A::add(s.ld(OFF_Q + i), s.ld(OFF_LAM + j)), rho_at(j), s.ld(OFF_V + j)));
                }
            // rhs = AHi q_hat - b                                                                         :111-119
            spmv_csr<A, false>(s, OFF_MU, OFF_QH, NR, C->AHi_val, C->AHi_col, C->AHi_row);
#pragma unroll
            for (int j = 0; j < n; ++j) s.st(OFF_MU + j, A::sub(s.ld(OFF_MU + j), s.ld(OFF_B + j)));
            // W mu = rhs                                                                                  :124-147
            ldl_solve_csc<A>(s, OFF_MU, NR, C->L_val, C->L_row, C->L_col, C->Dinv);
            // z = Hi q_hat + HiA mu                                                                       :153-165
            spmv_csr<A, false>(s, OFF_Z, OFF_QH, DIM, C->Hi_val, C->Hi_col, C->Hi_row);
            spmv_csr<A, true>(s, OFF_Z, OFF_MU, nrow_HiA, C->HiA_val, C->HiA_col, C->HiA_row);
            // v, lambda, residuals                                                                        :169-216
            bool over = false;
#pragma unroll 1
            for (int j = 0; j < DIM; ++j) {
                const real z = s.ld(OFF_Z + j), lam = s.ld(OFF_LAM + j), vo = s.ld(OFF_V + j);
                const real v = clip(A::madd(z, rho_i_at(j), lam), C->LB[j], C->UB[j]);
                s.st(OFF_V + j, v);
                s.st(OFF_LAM + j, A::madd(lam, rho_at(j), A::sub(z, v)));
                over |= exceeds(A::sub(vo, v), (real)tol) || exceeds(A::sub(z, v), (real)tol);
            }
            return !over;
        }

        __device__ void finish(long long inst, int k, int ef) {
#pragma unroll
            for (int j = 0; j < m; ++j) io.u[inst * m + j] = eng_u_out(C, (double)s.ld(OFF_V + 2 * n + j), j);   // u_opt = v[2 n + j]   :228-239
            io.k[inst] = k;
            io.e[inst] = ef;
            if (io.sol) {   // sol_<name>: z, v, lambda (header_MPCT_ADMM_cs_C.h)
                double *o = io.sol + inst * (long long)(sizeof(SPCIES_SOL_T) / sizeof(double));
                for (int e = 0; e < 3 * DIM; ++e) o[e] = (double)s.ld(OFF_Z + e);
                for (int e = 3 * DIM; e < (int)(sizeof(SPCIES_SOL_T) / sizeof(double)); ++e) o[e] = 0.0;
            }
        }
    };
};

#include "MPCT_ADMM_cs_mma.cuh"

typedef dense::DenseTraits<Solver, Engine> Traits;

}  // namespace mpct_cs
}  // namespace spcies

#define SPCIES_TRAITS ::spcies::mpct_cs::Traits
#include "spcies_entry.cuh"
