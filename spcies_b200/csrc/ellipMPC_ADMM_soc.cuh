// ellipMPC_ADMM_soc.cuh -- batched ADMM solver for the ellipMPC formulation with the terminal ellipsoid imposed as a
// second-order cone, hand-written for sm_100a.
//
// Per instance it performs exactly the arithmetic of formulations/+ellipMPC/code_ellipMPC_ADMM_soc_C.c:84-283:
//   q_hat = [q + lambda - sigma z ; mu - rho s]                                   :149-154
//   rhs   = (-Gh Hh^-1) q_hat - bh                 CSR mat-vec (sp_utils/smv.m)   :157-165
//   W rhs = rhs                                    CSC L D L' solve (LDLsolve.m)  :172-188
//   (z_hat, s_hat) = (-Hh^-1) q_hat + (-Hh^-1 Gh') rhs     2 x CSR mat-vec        :193-205
//   z = clip(z_hat + lambda/sigma) on the first dim-n-1 entries                   :210-217
//   s = proj_SOC(s_hat + mu/rho)                    (sp_utils/proj_SOC.m)         :220-242
//   lambda += sigma (z_hat - z);  mu += rho (s_hat - s)                           :247-254
//   exit on |primal_prev - primal| <= tol_d  and  |primal - primal_hat| <= tol_p  :260-281
// The sparse structure is the generator's (full2CSR / full2CSC / full2LDL); the index arrays travel with the
// constants into shared memory, the per-instance vectors are dynamically indexed in the [element][thread] state.
// `q` and `bh` are not stored: q repeats (Q xr, R ur) per stage and bh has 2n+1 non-zeros (b = -A x0, r, -PhiP xr);
// subtracting the remaining exact zeros of bh is the identity in IEEE arithmetic, so skipping them is bit-exact.
#pragma once
#include "spcies_kernel.cuh"
#include "spcies_mma.cuh"
#include "spcies_sparse.cuh"

namespace spcies {
namespace soc {

constexpr int n = nn_, m = mm_, nm = nm_, N = NN_;
constexpr int DIM = dim, NS = n_s, NEQ = n_eq;
constexpr int NP = DIM + NS;     // primal / dual length
constexpr int NR = NEQ + NS;     // rows of the W system

struct Solver {
    typedef SPCIES_REAL real;
    static constexpr int OFF_P = 0;               // primal = (z, s)
    static constexpr int OFF_D = OFF_P + NP;      // dual = (lambda, mu)
    static constexpr int OFF_PH = OFF_D + NP;     // primal_hat = (z_hat, s_hat)
    static constexpr int OFF_QH = OFF_PH + NP;    // q_hat
    static constexpr int OFF_RHS = OFF_QH + NP;   // rhs[n_eq + n_s]
    static constexpr int OFF_QX = OFF_RHS + NR;   // Q xr   [n]
    static constexpr int OFF_QU = OFF_QX + n;     // R ur   [m]
    static constexpr int OFF_QT = OFF_QU + m;     // T xr   [n]
    static constexpr int OFF_B0 = OFF_QT + n;     // -A x0  [n]
    static constexpr int OFF_BP = OFF_B0 + n;     // -PhiP xr [n]
    static constexpr int OFF_R = OFF_BP + n;      // r_ellip
    static constexpr int STATE = OFF_R + 1;
    static constexpr int STATE_VARB = STATE;
    static constexpr bool HAS_VARB = false;

    template <class A, bool VARB, class ST> struct Ctx {
        const spcies_consts *C;
        ST s;
        const BatchIO &io;
        __device__ Ctx(const spcies_consts *C_, ST s_, const BatchIO &io_) : C(C_), s(s_), io(io_) {}

        // set-up                                                              code_ellipMPC_ADMM_soc_C.c:84-131
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
                real b = real(0), bp = real(0), qx = real(0), qt = real(0);
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    b = A::nmsub(b, C->A[j][i], x0[i]);          // bh[j] -= A[j][i]*x0[i]
                    bp = A::nmsub(bp, C->PhiP[j][i], xr[i]);     // bh[n_eq+1+j] -= PhiP[j][i]*xr[i]
                    qx = A::madd(qx, C->Q[j][i], xr[i]);         // q += Q[j][i]*xr[i]   (Q, R, T stored negated)
                    qt = A::madd(qt, C->T[j][i], xr[i]);
                }
                s.st(OFF_B0 + j, b);
                s.st(OFF_BP + j, bp);
                s.st(OFF_QX + j, qx);
                s.st(OFF_QT + j, qt);
            }
#pragma unroll
            for (int j = 0; j < m; ++j) {
                real qu = real(0);
#pragma unroll
                for (int i = 0; i < m; ++i) qu = A::madd(qu, C->R[j][i], ur[i]);
                s.st(OFF_QU + j, qu);
            }
            s.st(OFF_R, (real)io.r[inst]);
#pragma unroll 4
            for (int e = 0; e < 2 * NP; ++e) s.st(OFF_P + e, real(0));   // primal = dual = 0
        }

        __device__ __forceinline__ real q_at(int idx) const {
            if (idx < m) return s.ld(OFF_QU + idx);
            if (idx >= m + (N - 1) * nm) return (idx < DIM - 1) ? s.ld(OFF_QT + idx - m - (N - 1) * nm) : real(0);
            const int t = (idx - m) % nm;
            return (t < n) ? s.ld(OFF_QX + t) : s.ld(OFF_QU + t - n);
        }

        __device__ bool iterate(int /*k*/) {
            const real sigma_ = C->sigma, sigma_i_ = C->sigma_i, rho_ = C->rho, rho_i_ = C->rho_i;
            // q_hat                                                                              :149-154
#pragma unroll 1
            for (int j = 0; j < DIM; ++j)
                s.st(OFF_QH + j, A::nmsub(A::add(q_at(j), s.ld(OFF_D + j)), sigma_, s.ld(OFF_P + j)));
#pragma unroll
            for (int j = 0; j < NS; ++j) s.st(OFF_QH + DIM + j, A::nmsub(s.ld(OFF_D + DIM + j), rho_, s.ld(OFF_P + DIM + j)));
            // rhs = GhHhi * q_hat - bh                                                            :157-165
            spmv_csr<A, false>(s, OFF_RHS, OFF_QH, nrow_GhHhi, C->GhHhi_val, C->GhHhi_col, C->GhHhi_row);
#pragma unroll
            for (int j = 0; j < n; ++j) {
                s.st(OFF_RHS + j, A::sub(s.ld(OFF_RHS + j), s.ld(OFF_B0 + j)));
                s.st(OFF_RHS + NEQ + 1 + j, A::sub(s.ld(OFF_RHS + NEQ + 1 + j), s.ld(OFF_BP + j)));
            }
            s.st(OFF_RHS + NEQ - 1, A::sub(s.ld(OFF_RHS + NEQ - 1), s.ld(OFF_R)));
            // W rhs = rhs through L D L'                                                          :172-188
            ldl_solve_csc<A>(s, OFF_RHS, nrow_GhHhi, C->L_val, C->L_row, C->L_col, C->Dinv);
            // primal_hat = Hhi q_hat + HhiGh rhs                                                  :193-205
            spmv_csr<A, false>(s, OFF_PH, OFF_QH, nrow_Hhi, C->Hhi_val, C->Hhi_col, C->Hhi_row);
            spmv_csr<A, true>(s, OFF_PH, OFF_RHS, nrow_HhiGh, C->HhiGh_val, C->HhiGh_col, C->HhiGh_row);

            bool over = false;
            // z, lambda and their residuals                                                       :210-217, :247-249, :260-270
#pragma unroll 1
            for (int j = 0; j < DIM; ++j) {
                const real zh = s.ld(OFF_PH + j), lam = s.ld(OFF_D + j), zo = s.ld(OFF_P + j);
                real z = A::madd(zh, sigma_i_, lam);
                if (j < DIM - n - 1) z = clip(z, C->LB[j], C->UB[j]);
                s.st(OFF_P + j, z);
                s.st(OFF_D + j, A::madd(lam, sigma_, A::sub(zh, z)));
                over |= exceeds(A::sub(zo, z), (real)tol_d) || exceeds(A::sub(z, zh), (real)tol_p);
            }
            // s = proj_SOC(s_hat + mu/rho), mu                                                    :220-242, :252-254
            real sv[NS], sh[NS], mu[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                sh[j] = s.ld(OFF_PH + DIM + j);
                mu[j] = s.ld(OFF_D + DIM + j);
                sv[j] = A::madd(sh[j], rho_i_, mu[j]);
            }
            proj_soc<A, NS>(sv);
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const real so = s.ld(OFF_P + DIM + j);
                s.st(OFF_P + DIM + j, sv[j]);
                s.st(OFF_D + DIM + j, A::madd(mu[j], rho_, A::sub(sh[j], sv[j])));
                over |= exceeds(A::sub(so, sv[j]), (real)tol_d) || exceeds(A::sub(sv[j], sh[j]), (real)tol_p);
            }
            return !over;
        }

        __device__ void finish(long long inst, int k, int ef) {
#pragma unroll
            for (int j = 0; j < m; ++j) io.u[inst * m + j] = eng_u_out(C, (double)s.ld(OFF_P + j), j);   // u_opt = z[0..m)   :297-306
            io.k[inst] = k;
            io.e[inst] = ef;
            if (io.sol) {   // sol_<name>: z, s, z_hat, s_hat, lambda, mu (header_ellipMPC_ADMM_soc_C.h)
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

#include "ellipMPC_ADMM_soc_mma.cuh"
#include "ellipMPC_ADMM_soc_band.cuh"

// Host-side traits: the scalar skeleton plus the two tensor-core engines (FAST arithmetic, no debug payload): the structured
// one (ellipMPC_ADMM_soc_band.cuh) when the generated constants have its structure, else the dense map (ellipMPC_ADMM_soc_mma.cuh;
// SPCIES_CUDA_SOC_ENGINE=dense in the environment forces it)
struct Traits : PolicyTraits<Solver> {
    typedef PolicyTraits<Solver> Base;
    static bool &band_ok() {
        static bool ok = false;
        return ok;
    }
    static size_t blob_bytes() {
        if (HAS_BAND) return BAND_OFFSET + BAND_BYTES;
        return HAS_MMA ? MMA_OFFSET + SMALL_BYTES + FRAG_BYTES : Base::blob_bytes();
    }
    static void fill_blob(void *dst) {
        memset(dst, 0, blob_bytes());
        Base::fill_blob(dst);
        if constexpr (HAS_MMA) {
            MmaSmall *S = new MmaSmall;
            fill_mma_tables(spcies_h_consts, *S, reinterpret_cast<double2 *>((char *)dst + MMA_OFFSET + SMALL_BYTES));
            memcpy((char *)dst + MMA_OFFSET, S, sizeof *S);
            delete S;
        }
        if constexpr (HAS_BAND) {
            BandTables *B = new BandTables;
            band_ok() = fill_band_tables(spcies_h_consts, *B);
            memcpy((char *)dst + BAND_OFFSET, B, sizeof *B);
            delete B;
        }
    }
    static bool use_mma(int arith, const BatchIO &io) {
        if constexpr (!HAS_MMA && !HAS_BAND) return false;
        return arith != SPCIES_CUDA_ARITH_EXACT && io.sol == nullptr && io.engine != SPCIES_CUDA_ENGINE_SCALAR && io.LB == nullptr &&
               (HAS_MMA || use_band());
    }
    static bool use_band() {
        if constexpr (!HAS_BAND) return false;
        const char *e = getenv("SPCIES_CUDA_SOC_ENGINE");
        const bool dense = e != nullptr && strcmp(e, "dense") == 0 && HAS_MMA;
        return band_ok() && !dense;
    }
    static bool uses_scratch(int arith, const BatchIO &io) { return !use_mma(arith, io); }
    static void engine_shape(int arith, const BatchIO &io, int &block, size_t &smem, int &ipb) {
        ipb = block;
        if (use_mma(arith, io)) {
            block = use_band() ? BAND_BLOCK : MMA_BLOCK;
            smem = use_band() ? BAND_SMEM : MMA_SMEM;
            ipb = use_band() ? BAND_IPB : MMA_IPB;
        }
    }
    static cudaError_t launch(int arith, bool varb, int grid, int block, size_t smem, cudaStream_t s, const BatchIO &io,
                              const void *dc, void *scratch) {
        if (io.engine == SPCIES_CUDA_ENGINE_MMA && !use_mma(arith, io)) return cudaErrorNotSupported;
        if (use_mma(arith, io)) {
            if constexpr (HAS_BAND) {
                if (use_band()) {
                    cudaError_t e = cudaFuncSetAttribute(soc_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BAND_SMEM);
                    if (e != cudaSuccess) return e;
                    soc_band_kernel<<<grid, BAND_BLOCK, BAND_SMEM, s>>>(io, (const unsigned char *)dc);
                    return cudaGetLastError();
                }
            }
            if constexpr (HAS_MMA) {
                cudaError_t e = cudaFuncSetAttribute(soc_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MMA_SMEM);
                if (e != cudaSuccess) return e;
                soc_mma_kernel<<<grid, MMA_BLOCK, MMA_SMEM, s>>>(io, (const unsigned char *)dc);
                return cudaGetLastError();
            }
        }
        return Base::launch(arith, varb, grid, block, smem, s, io, dc, scratch);
    }
    static cudaError_t attributes(int arith, bool varb, cudaFuncAttributes *a) {
        if (arith != SPCIES_CUDA_ARITH_EXACT) {
            if constexpr (HAS_BAND) {
                if (use_band()) return cudaFuncGetAttributes(a, soc_band_kernel);
            }
            if constexpr (HAS_MMA) return cudaFuncGetAttributes(a, soc_mma_kernel);
        }
        return Base::attributes(arith, varb, a);
    }
};

}  // namespace soc
}  // namespace spcies

#define SPCIES_TRAITS ::spcies::soc::Traits
#include "spcies_entry.cuh"
