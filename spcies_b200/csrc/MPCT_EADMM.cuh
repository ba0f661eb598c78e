// MPCT_EADMM.cuh -- batched three-block extended-ADMM solver for the MPC-for-tracking formulation (artificial
// reference), hand-written for sm_100a.
//
// Per instance it performs exactly the arithmetic of formulations/+MPCT/code_MPCT_EADMM_C.c:85-459 (IS_DIAG path):
//   P1  z1_l = clip(H1i o (rho_l (z3_l + z2) [+ rho_0 x0] + lambda ...))                      :97-117
//   P2  q2 (a sequential accumulation over the horizon), z2 = W2 q2                            :123-149
//   P3  q3_l = rho_l (z2 - z1_l) + lambda_{l+1};  rhs = -G3 H3^-1 q3;  banded-Cholesky solve;
//       z3_l = -H3i o (q3_l + G3' mu)                                                          :157-366
//   residual res = A1 z1 + A2 z2 + A3 z3 - b, lambda += rho o res                              :371-402
//   exit on |z2_prev - z2|, |res|, |z3_prev - z3| <= tol                                       :408-457
// Fusions (none changes an operand order, so Arith<EXACT> stays bit-identical):
//   * z1_N is computed first so that the q2 accumulation can run in the same sweep as the z1 update;
//   * q3 is never stored (the reference keeps it in z3): it is recomputed where needed, which also keeps the previous
//     z3 alive for the |z3_prev - z3| test without a z3_prev array;
//   * z3, the residual, the lambda update and the exit tests are folded into the backward-substitution sweep.
// Persistent state per instance: z1, z3 [(N+1) nm], lambda [(N+3) nm], z2 [nm], mu [N n], x0, xr, ur.  At N = 50 that
// is 12.5 KB, so the skeleton places it in the L2-resident global scratch instead of shared memory.
#pragma once
#include "spcies_kernel.cuh"
#include "spcies_mma.cuh"

#if !defined(IS_DIAG) || IS_DIAG != 1
#error "MPCT_EADMM.cuh implements the IS_DIAG == 1 path (diagonal Q and R) of code_MPCT_EADMM_C.c"
#endif

namespace spcies {
namespace eadmm {

constexpr int n = nn_, m = mm_, nm = nm_, N = NN_;

struct Solver {
    typedef SPCIES_REAL real;
    static constexpr int OFF_Z1 = 0;                        // z1[N+1][nm]
    static constexpr int OFF_Z3 = OFF_Z1 + (N + 1) * nm;    // z3[N+1][nm]
    static constexpr int OFF_LAM = OFF_Z3 + (N + 1) * nm;   // lambda[N+3][nm]
    static constexpr int OFF_Z2 = OFF_LAM + (N + 3) * nm;   // z2[nm]
    static constexpr int OFF_MU = OFF_Z2 + nm;              // mu[N][n]
    static constexpr int OFF_X0 = OFF_MU + N * n;           // x0[n]
    static constexpr int OFF_XR = OFF_X0 + n;               // xr[n]
    static constexpr int OFF_UR = OFF_XR + n;               // ur[m]
    static constexpr int STATE = OFF_UR + m;
    static constexpr int STATE_VARB = STATE;
    static constexpr bool HAS_VARB = false;

    template <class A, bool VARB, class ST> struct Ctx {
        const spcies_consts *C;
        ST s;
        const BatchIO &io;
        __device__ Ctx(const spcies_consts *C_, ST s_, const BatchIO &io_) : C(C_), s(s_), io(io_) {}

        __device__ void init(long long inst) {
#pragma unroll
            for (int i = 0; i < n; ++i) {
                s.st(OFF_X0 + i, (real)eng_x(C, io.x0, inst, n, i));
                s.st(OFF_XR + i, (real)eng_x(C, io.xr, inst, n, i));
            }
#pragma unroll
            for (int i = 0; i < m; ++i) s.st(OFF_UR + i, (real)eng_u(C, io.ur, inst, m, i));
#pragma unroll 4
            for (int e = 0; e < OFF_MU; ++e) s.st(e, real(0));   // z1 = z3 = lambda = z2 = 0
        }

        __device__ __forceinline__ void fwd_block(real (&mu)[n], const real (&mprev)[n], int l, bool first) const {
            if (!first) {
#pragma unroll
                for (int i = 0; i < n; ++i)
#pragma unroll
                    for (int j = 0; j < n; ++j) mu[j] = A::nmsub(mu[j], C->Alpha[l - 1][i][j], mprev[i]);
            }
#pragma unroll
            for (int j = 0; j < n; ++j) {
#pragma unroll
                for (int i = 0; i < j; ++i) mu[j] = A::nmsub(mu[j], C->Beta[l][i][j], mu[i]);
                mu[j] = A::mul(C->Beta[l][j][j], mu[j]);
            }
        }
        __device__ __forceinline__ void bwd_block(real (&mu)[n], const real (&mnext)[n], int l, bool last) const {
#pragma unroll
            for (int j = n - 1; j >= 0; --j) {
                if (!last) {
#pragma unroll
                    for (int i = n - 1; i >= 0; --i) mu[j] = A::nmsub(mu[j], C->Alpha[l][j][i], mnext[i]);
                }
#pragma unroll
                for (int i = n - 1; i >= j + 1; --i) mu[j] = A::nmsub(mu[j], C->Beta[l][j][i], mu[i]);
                mu[j] = A::mul(C->Beta[l][j][j], mu[j]);
            }
        }
        // q3_l = rho_l (z2 - z1_l) + lambda_{l+1}                                                  :157-172
        __device__ __forceinline__ void q3_block(real (&q3)[nm], const real (&z2)[nm], int l) const {
#pragma unroll
            for (int j = 0; j < nm; ++j)
                q3[j] = A::madd(s.ld(OFF_LAM + (l + 1) * nm + j), C->rho[l][j], A::sub(z2[j], s.ld(OFF_Z1 + l * nm + j)));
        }
        // res_{l+1} = z2 + z3_l - z1_l;  lambda_{l+1} += rho_l res;  |res|, |z3_prev - z3| tests    :377-381, :390-394, :421-449
        __device__ __forceinline__ bool close_stage(const real (&z3n)[nm], const real (&z2)[nm], int l) const {
            bool over = false;
#pragma unroll
            for (int j = 0; j < nm; ++j) {
                const real res = A::sub(A::add(z2[j], z3n[j]), s.ld(OFF_Z1 + l * nm + j));
                over |= exceeds(res, (real)tol);
                over |= exceeds(A::sub(s.ld(OFF_Z3 + l * nm + j), z3n[j]), (real)tol);
                s.st(OFF_Z3 + l * nm + j, z3n[j]);
                s.st(OFF_LAM + (l + 1) * nm + j, A::madd(s.ld(OFF_LAM + (l + 1) * nm + j), C->rho[l][j], res));
            }
            return over;
        }

        __device__ bool iterate(int /*k*/) {
            real z2[nm], q2[nm], z1N[nm];
            bool over = false;
#pragma unroll
            for (int j = 0; j < nm; ++j) z2[j] = s.ld(OFF_Z2 + j);

            // ---------- P1 (last block first) and the head of q2                                   :112-117, :123-136
#pragma unroll
            for (int j = 0; j < nm; ++j) {
                const real z3N = s.ld(OFF_Z3 + N * nm + j), lA = s.ld(OFF_LAM + (N + 1) * nm + j),
                           lB = s.ld(OFF_LAM + (N + 2) * nm + j);
                const real rs = A::add(C->rho[N][j], C->rho_s[j]);
                real v = A::add(A::add(A::madd(A::mul(C->rho[N][j], z3N), rs, z2[j]), lA), lB);
                v = clip(A::mul(v, C->H1i[N][j]), C->LB_s[j], C->UB_s[j]);
                z1N[j] = v;
                s.st(OFF_Z1 + N * nm + j, v);
                q2[j] = A::add(A::add(A::nmsub(A::mul(C->rho[N][j], z3N), rs, v), lA), lB);
            }
#pragma unroll
            for (int j = 0; j < n; ++j)
#pragma unroll
                for (int i = 0; i < n; ++i) q2[j] = A::madd(q2[j], C->T[j][i], s.ld(OFF_XR + i));
#pragma unroll
            for (int j = 0; j < m; ++j)
#pragma unroll
                for (int i = 0; i < m; ++i) q2[n + j] = A::madd(q2[n + j], C->S[j][i], s.ld(OFF_UR + i));
            // ---------- P1 for l = 0..N-1 fused with the q2 accumulation                           :97-110, :137-141
#pragma unroll 1
            for (int l = 0; l < N; ++l) {
#pragma unroll
                for (int j = 0; j < nm; ++j) {
                    const real z3l = s.ld(OFF_Z3 + l * nm + j), lam1 = s.ld(OFF_LAM + (l + 1) * nm + j);
                    real v;
                    if (l == 0) {
                        const real x0j = (j < n) ? s.ld(OFF_X0 + j) : real(0);   // x0 is zero-padded to nm (:35)
                        v = A::sub(A::add(A::madd(A::mul(C->rho[0][j], A::add(z3l, z2[j])), C->rho_0[j], x0j), lam1),
                                   s.ld(OFF_LAM + j));
                        v = clip(A::mul(v, C->H1i[0][j]), C->LB_0[j], C->UB_0[j]);
                    } else {
                        v = A::add(A::mul(C->rho[l][j], A::add(z3l, z2[j])), lam1);
                        v = clip(A::mul(v, C->H1i[l][j]), C->LB[j], C->UB[j]);
                    }
                    s.st(OFF_Z1 + l * nm + j, v);
                    q2[j] = A::add(A::madd(q2[j], C->rho[l][j], A::sub(z3l, v)), lam1);   // (q2 + rho*(z3 - z1)) + lambda
                }
            }
            // ---------- P2: z2 = W2 q2                                                              :145-149
            {
                real z2n[nm];
#pragma unroll
                for (int j = 0; j < nm; ++j) {
                    real a = real(0);
#pragma unroll
                    for (int i = 0; i < nm; ++i) a = A::madd(a, C->W2[j][i], q2[i]);
                    z2n[j] = a;
                    over |= exceeds(A::sub(z2[j], a), (real)tol);                                  // :412-418
                }
#pragma unroll
                for (int j = 0; j < nm; ++j) {
                    z2[j] = z2n[j];
                    s.st(OFF_Z2 + j, z2n[j]);
                }
            }
            // ---------- P3 forward: rhs_l = H3i_{l+1} q3_{l+1} - AB (H3i_l q3_l), forward substitution      :176-184, :221-251
            real qa[nm], qb[nm], mu[n], mprev[n], mnext[n];
            q3_block(qa, z2, 0);
#pragma unroll 1
            for (int l = 0; l < N; ++l) {
                q3_block(qb, z2, l + 1);
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    real r = A::mul(C->H3i[l + 1][j], qb[j]);
#pragma unroll
                    for (int i = 0; i < nm; ++i) r = A::nmsub(r, cprod<A>(C->AB[j][i], C->H3i[l][i]), qa[i]);
                    mu[j] = r;
                }
                fwd_block(mu, mprev, l, l == 0);
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    if (l < N - 1) s.st(OFF_MU + l * n + j, mu[j]);
                    mprev[j] = mu[j];
                }
#pragma unroll
                for (int j = 0; j < nm; ++j) qa[j] = qb[j];
            }
            // ---------- P3 backward + z3 + residual + lambda + exit tests                        :254-287, :291-320, :371-449
            bwd_block(mu, mnext, N - 1, true);
            {   // z3_N = -H3i_N o (q3_N - [mu_{N-1}; 0])                                              :313-320
                real z3n[nm];
#pragma unroll
                for (int j = 0; j < nm; ++j) {
                    real v = qa[j];   // q3_N (left in qa by the forward sweep)
                    if (j < n) v = A::sub(v, mu[j]);
                    z3n[j] = A::mul(-C->H3i[N][j], v);
                }
                over |= close_stage(z3n, z2, N);
            }
#pragma unroll
            for (int j = 0; j < n; ++j) mnext[j] = mu[j];
#pragma unroll 1
            for (int l = N - 2; l >= 0; --l) {
#pragma unroll
                for (int j = 0; j < n; ++j) mu[j] = s.ld(OFF_MU + l * n + j);
                bwd_block(mu, mnext, l, false);
                // z3_{l+1} = -H3i_{l+1} o (q3_{l+1} - [mu_l; 0] + [A B]' mu_{l+1})                    :299-310
                real z3n[nm];
                q3_block(qb, z2, l + 1);
#pragma unroll
                for (int j = 0; j < nm; ++j) {
                    real v = qb[j];
                    if (j < n) v = A::sub(v, mu[j]);
#pragma unroll
                    for (int i = 0; i < n; ++i) v = A::madd(v, C->AB[i][j], mnext[i]);
                    z3n[j] = A::mul(-C->H3i[l + 1][j], v);
                }
                over |= close_stage(z3n, z2, l + 1);
#pragma unroll
                for (int j = 0; j < n; ++j) mnext[j] = mu[j];
            }
            {   // z3_0 = -H3i_0 o (q3_0 + [A B]' mu_0)                                                 :291-296
                real z3n[nm];
                q3_block(qb, z2, 0);
#pragma unroll
                for (int j = 0; j < nm; ++j) {
                    real v = qb[j];
#pragma unroll
                    for (int i = 0; i < n; ++i) v = A::madd(v, C->AB[i][j], mnext[i]);
                    z3n[j] = A::mul(-C->H3i[0][j], v);
                }
                over |= close_stage(z3n, z2, 0);
            }
            // res_0 = z1_0[0:n] - x0, lambda_0;  res_{N+2} = z2 - z1_N, lambda_{N+2}                   :371-402
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const real res = A::sub(s.ld(OFF_Z1 + j), s.ld(OFF_X0 + j));
                over |= exceeds(res, (real)tol);
                s.st(OFF_LAM + j, A::madd(s.ld(OFF_LAM + j), C->rho_0[j], res));
            }
#pragma unroll
            for (int j = 0; j < nm; ++j) {
                const real res = A::sub(z2[j], z1N[j]);
                over |= exceeds(res, (real)tol);
                s.st(OFF_LAM + (N + 2) * nm + j, A::madd(s.ld(OFF_LAM + (N + 2) * nm + j), C->rho_s[j], res));
            }
            return !over;
        }

        __device__ void finish(long long inst, int k, int ef) {
#pragma unroll
            for (int j = 0; j < m; ++j) io.u[inst * m + j] = eng_u_out(C, (double)s.ld(OFF_Z1 + n + j), j);   // u_opt = z1[0][n..]  :470-478
            io.k[inst] = k;
            io.e[inst] = ef;
            if (io.sol) {   // sol_<name>: z1, z2, z3, lambda (header_MPCT_EADMM_C.h); lambda written in full [l][j] order
                double *o = io.sol + inst * (long long)(sizeof(SPCIES_SOL_T) / sizeof(double));
                constexpr int L1 = (N + 1) * nm;
                for (int e = 0; e < L1; ++e) {
                    o[e] = (double)s.ld(OFF_Z1 + e);
                    o[L1 + nm + e] = (double)s.ld(OFF_Z3 + e);
                }
                for (int e = 0; e < nm; ++e) o[L1 + e] = (double)s.ld(OFF_Z2 + e);
                for (int e = 0; e < (N + 3) * nm; ++e) o[2 * L1 + nm + e] = (double)s.ld(OFF_LAM + e);
                for (int e = 2 * L1 + nm + (N + 3) * nm; e < (int)(sizeof(SPCIES_SOL_T) / sizeof(double)); ++e) o[e] = 0.0;
            }
        }
    };
};

#include "MPCT_EADMM_mma.cuh"

// Host-side traits: the scalar skeleton plus the tensor-core engine (FAST arithmetic, no debug payload)
struct Traits : PolicyTraits<Solver> {
    typedef PolicyTraits<Solver> Base;
    static size_t blob_bytes() { return HAS_MMA ? MMA_OFFSET + SMALL_BYTES + FRAG_BYTES : Base::blob_bytes(); }
    static void fill_blob(void *dst) {
        memset(dst, 0, blob_bytes());
        Base::fill_blob(dst);
        if constexpr (HAS_MMA) {
            MmaSmall *S = new MmaSmall;
            MmaFrag *F = new MmaFrag;
            fill_mma_tables(spcies_h_consts, *S, *F);
            memcpy((char *)dst + MMA_OFFSET, S, sizeof *S);
            memcpy((char *)dst + MMA_OFFSET + SMALL_BYTES, F, sizeof *F);
            delete S;
            delete F;
        }
    }
    static bool use_mma(int arith, const BatchIO &io) {
        if constexpr (!HAS_MMA) return false;
        return arith != SPCIES_CUDA_ARITH_EXACT && io.sol == nullptr && io.engine != SPCIES_CUDA_ENGINE_SCALAR && io.LB == nullptr;
    }
    static bool uses_scratch(int arith, const BatchIO &io) { return !use_mma(arith, io); }
    // the engine streams mu' (forward sweep -> backward sweep) through a per-warp global buffer
    static size_t engine_scratch_bytes(int arith, const BatchIO &io, int grid) {
        return use_mma(arith, io) ? (size_t)grid * MMA_WARPS * MUP_BYTES_PER_WARP : 0;
    }
    static void engine_shape(int arith, const BatchIO &io, int &block, size_t &smem, int &ipb) {
        ipb = block;
        if (use_mma(arith, io)) {
            block = MMA_BLOCK;
            smem = MMA_SMEM;
            ipb = MMA_IPB;
        }
    }
    static cudaError_t launch(int arith, bool varb, int grid, int block, size_t smem, cudaStream_t s, const BatchIO &io,
                              const void *dc, void *scratch) {
        if (io.engine == SPCIES_CUDA_ENGINE_MMA && !use_mma(arith, io)) return cudaErrorNotSupported;
        if constexpr (HAS_MMA) {
            if (use_mma(arith, io)) {
                cudaError_t e = cudaFuncSetAttribute(eadmm_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MMA_SMEM);
                if (e != cudaSuccess) return e;
                eadmm_mma_kernel<<<grid, MMA_BLOCK, MMA_SMEM, s>>>(io, (const unsigned char *)dc, (double2 *)scratch);
                return cudaGetLastError();
            }
        }
        return Base::launch(arith, varb, grid, block, smem, s, io, dc, scratch);
    }
    static cudaError_t attributes(int arith, bool varb, cudaFuncAttributes *a) {
        if constexpr (HAS_MMA) {
            if (arith != SPCIES_CUDA_ARITH_EXACT) return cudaFuncGetAttributes(a, eadmm_mma_kernel);
        }
        return Base::attributes(arith, varb, a);
    }
};

}  // namespace eadmm
}  // namespace spcies

#define SPCIES_TRAITS ::spcies::eadmm::Traits
#include "spcies_entry.cuh"
