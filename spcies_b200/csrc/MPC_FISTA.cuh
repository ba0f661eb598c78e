// MPC_FISTA.cuh -- batched FISTA solver for the laxMPC (SPCIES_TERMINAL == 1) and equMPC
// (SPCIES_TERMINAL == 0) formulations, hand-written for sm_100a.
//
// Replaces, per instance, exactly the arithmetic of the reference templates
//   formulations/+laxMPC/code_laxMPC_FISTA_C.c:275-456  (driver), :471-539 (compute_z_lambda),
//                                              :546-574  (compute_residual_vector), :577-651 (solve_W_matrix_form)
//   formulations/+equMPC/code_equMPC_FISTA_C.c            (same, without the terminal block)
//
// Mapping
//   * one thread = one MPC instance; one persistent CTA per SM; lanes pull instances from a global
//     queue as they finish (iteration counts are heavy-tailed).
//   * iterates y, lambda (and the solve workspace mu) live in shared memory, [element][thread];
//     everything with a stage-local lifetime (z_l, r_l, the running mu_{l-1}) lives in registers.
//   * the shared problem constants (AB, Alpha, Beta, QRi, bounds...) are staged once per CTA into shared
//     memory with one bulk async copy; AB is then held in registers for the whole kernel.
//   * one FISTA iteration = two fused sweeps over the horizon
//        pass A (l = 0..N-1):  z_l(y) -> r_l -> exit test -> forward substitution  mu_l
//        pass B (l = N-1..0):  backward substitution d_lambda_l -> lambda_l = y_l + d_lambda_l -> y_l update
//     so z and the residual are never materialised (the reference stores z[N-1][nm], d_lambda[N][n]).
//     The warm-up step of the reference (:300-320) is the same two passes with the exit test off.
//   * operation order inside every accumulation is the reference's, so with Arith<EXACT> the
//     iterates are bit-identical to gcc -O3; Arith<FAST> only fuses a*b+c and hoists 1/t.
//
// The including .cu (emitted by platforms/cuda_code.py) defines SPCIES_REAL, SPCIES_TERMINAL, SPCIES_SOL_T,
// SPCIES_FUNC, the reference #defines (nn_, mm_, nm_, NN_, k_max, tol [, VAR_BOUNDS]) and
// `struct spcies_consts` + `spcies_h_consts` holding LB, UB, AB, Alpha, Beta, Q, R, QRi [, T, Ti].
#pragma once
#include "spcies_host.cuh"

// Kernel variant (compile-time, chosen by the generator):
//   SPCIES_MU_REGS = 0   stage loops rolled; the forward-substitution result mu[N][n] goes through shared memory and
//                        [A B] is pinned in registers.  Smallest code, 200 doubles of shared memory per instance (N=10).
//   SPCIES_MU_REGS = 1   stage loops fully unrolled; mu[N][n] stays in registers between the two passes and [A B] is
//                        read from shared memory (warp-broadcast loads).  140 doubles of shared memory per instance,
//                        i.e. 6 instead of 4 resident warps per SM at N=10, and ptxas can overlap the triangular solve of
//                        stage l with the z / residual products of stage l+1.
// Measured on B200 (profiles/r1_fista_variants.md): variant 1 is 1.66x SLOWER than variant 0 at N = 10 -- its 88 KB of
// straight-line code misses the instruction cache (stall reason no_instruction = 2.2 per issue).  Default: 0.
#ifndef SPCIES_MU_REGS
#define SPCIES_MU_REGS 0
#endif
#ifndef SPCIES_UNROLL_STAGES
#define SPCIES_UNROLL_STAGES (SPCIES_MU_REGS ? NN_ : 1)
#endif

// Stage boundary of the unrolled variant: keeps the compiler from hoisting the (alias-free) shared-memory loads of
// all ten stages to the top of the pass, which would spill hundreds of registers.
#if SPCIES_MU_REGS
#define SPCIES_STAGE_FENCE() asm volatile("" ::: "memory")
#else
#define SPCIES_STAGE_FENCE() do { } while (0)
#endif

namespace spcies {
namespace fista {

typedef SPCIES_REAL real;
constexpr int n = nn_, m = mm_, nm = nm_, N = NN_;
constexpr bool TERMINAL = (SPCIES_TERMINAL != 0);
constexpr int UNROLL_STAGES = SPCIES_UNROLL_STAGES;
constexpr bool MU_REGS = (SPCIES_MU_REGS != 0);
static_assert(!MU_REGS || UNROLL_STAGES >= NN_, "mu in registers needs fully unrolled stage loops");

// shared-memory state elements per instance
constexpr int OFF_Y = 0;                 // y[N][n]      linearisation point
constexpr int OFF_LAM = OFF_Y + N * n;   // lambda[N][n]
constexpr int OFF_MU = OFF_LAM + N * n;  // mu[N][n]     W-solve workspace (forward result; only if !MU_REGS)
constexpr int OFF_B = OFF_MU + (MU_REGS ? 0 : N * n);   // b[n] = -A x0
constexpr int OFF_Q = OFF_B + n;         // q[nm] = [Q xr; R ur]   (Q, R stored negated)
constexpr int OFF_QT = OFF_Q + nm;       // qT[n] = T xr (lax)  |  xr (equ)
constexpr int STATE_FIXED = OFF_QT + n;
constexpr int OFF_LB = STATE_FIXED;      // per-instance bounds (VARB kernels only)
constexpr int OFF_UB = OFF_LB + nm;
constexpr int STATE_VARB = OFF_UB + nm;

constexpr size_t CONSTS_BYTES = (sizeof(spcies_consts) + 15) / 16 * 16;
constexpr size_t SMEM_MAX = 227 * 1024;

constexpr int block_for(int elems) {
    int t = (int)((SMEM_MAX - CONSTS_BYTES - 64) / ((size_t)elems * sizeof(real))) / 32 * 32;
    return t > 1024 ? 1024 : t;
}
constexpr int BLOCK_FIXED = block_for(STATE_FIXED);
constexpr int BLOCK_VARB = block_for(STATE_VARB);
static_assert(BLOCK_VARB >= 32, "per-instance state does not fit shared memory with one thread per instance");

template <bool VARB> struct Bounds {
    // stage bounds: u_0 (j in [0,m)), stage l (j in [0,nm)), terminal x_N (j in [0,n))
    const spcies_consts *C;
    const real *st;  // state base (VARB)
    int stride;
    __device__ __forceinline__ real lb0(int j) const {
#ifdef VAR_BOUNDS
        return VARB ? st[(OFF_LB + n + j) * stride] : C->LB0[j];
#else
        return VARB ? st[(OFF_LB + n + j) * stride] : C->LB[n + j];
#endif
    }
    __device__ __forceinline__ real ub0(int j) const {
#ifdef VAR_BOUNDS
        return VARB ? st[(OFF_UB + n + j) * stride] : C->UB0[j];
#else
        return VARB ? st[(OFF_UB + n + j) * stride] : C->UB[n + j];
#endif
    }
    __device__ __forceinline__ real lb(int l, int j) const {
#ifdef VAR_BOUNDS
        return VARB ? st[(OFF_LB + j) * stride] : C->LB[l][j];
#else
        return VARB ? st[(OFF_LB + j) * stride] : C->LB[j];
#endif
    }
    __device__ __forceinline__ real ub(int l, int j) const {
#ifdef VAR_BOUNDS
        return VARB ? st[(OFF_UB + j) * stride] : C->UB[l][j];
#else
        return VARB ? st[(OFF_UB + j) * stride] : C->UB[j];
#endif
    }
#if SPCIES_TERMINAL
    __device__ __forceinline__ real lbN(int j) const {
#ifdef VAR_BOUNDS
        return VARB ? st[(OFF_LB + j) * stride] : C->LBN[j];
#else
        return VARB ? st[(OFF_LB + j) * stride] : C->LB[j];
#endif
    }
    __device__ __forceinline__ real ubN(int j) const {
#ifdef VAR_BOUNDS
        return VARB ? st[(OFF_UB + j) * stride] : C->UBN[j];
#else
        return VARB ? st[(OFF_UB + j) * stride] : C->UB[j];
#endif
    }
#endif
};

// z_l = clip(QRi o (q - [A B]' y_{l+1} + [y_l; 0]))            code_laxMPC_FISTA_C.c:494-519
template <class A, bool VARB>
__device__ __forceinline__ void z_stage(real (&z)[nm], const real (&AB)[n][nm], const real (&yl)[n],
                                        const real (&yn)[n], const real (&q)[nm], const spcies_consts *C,
                                        const Bounds<VARB> &bd, int l) {
#pragma unroll
    for (int j = 0; j < nm; ++j) z[j] = q[j];
#pragma unroll
    for (int i = 0; i < n; ++i) {
        SPCIES_STAGE_FENCE();   // one row of [A B] in flight at a time
#pragma unroll
        for (int j = 0; j < nm; ++j) z[j] = A::nmsub(z[j], AB[i][j], yn[i]);
    }
#pragma unroll
    for (int j = 0; j < n; ++j) z[j] = A::add(z[j], yl[j]);
#pragma unroll
    for (int j = 0; j < nm; ++j) z[j] = clip(A::mul(z[j], C->QRi[j]), bd.lb(l, j), bd.ub(l, j));
}

// forward substitution of one block row                        code_laxMPC_FISTA_C.c:585-614
template <class A>
__device__ __forceinline__ void fwd_block(real (&mu)[n], const real (&mprev)[n], const spcies_consts *C, int l,
                                          bool first) {
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

// backward substitution of one block row                       code_laxMPC_FISTA_C.c:619-648
template <class A>
__device__ __forceinline__ void bwd_block(real (&mu)[n], const real (&mnext)[n], const spcies_consts *C, int l,
                                          bool last) {
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

template <bool EXACT, bool VARB, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1) fista_kernel(const BatchIO io, const spcies_consts *__restrict__ g_consts) {
    typedef Arith<real, EXACT> A;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    spcies_consts *C = reinterpret_cast<spcies_consts *>(smem_raw);
    stage_constants(C, g_consts, (uint32_t)CONSTS_BYTES, &mbar);

    real *stbase = reinterpret_cast<real *>(smem_raw + CONSTS_BYTES) + threadIdx.x;
    auto LD = [&](int e) -> real { return stbase[e * BLOCK]; };
    auto ST = [&](int e, real v) { stbase[e * BLOCK] = v; };
    Bounds<VARB> bd{C, stbase, BLOCK};

    // [A B] is reused by every stage of every iteration: pinned in registers (MU_REGS = 0) or read through
    // warp-broadcast shared-memory loads (MU_REGS = 1, where the registers hold mu instead)
#if SPCIES_MU_REGS
    const real(&AB)[n][nm] = C->AB;
#else
    real AB[n][nm];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < nm; ++j) AB[i][j] = C->AB[i][j];
#endif

    const real tol_ = (real)tol;
    WorkQueue wq{io.queue, io.B};
    unsigned long long stat_k = 0;
    unsigned int stat_nc = 0;

    long long inst = -1;
    int k = 0;
    bool warm = false;
    real t = real(1);

    for (;;) {
        if (inst < 0) {
            inst = wq.next();
            if (inst < 0) break;
            // ---- per-instance set-up                                   code_laxMPC_FISTA_C.c:94-100, 275-289
            real x0[n], xr[n], ur[m];
#pragma unroll
            for (int i = 0; i < n; ++i) {
                x0[i] = (real)io.x0[inst * n + i];
                xr[i] = (real)io.xr[inst * n + i];
            }
#pragma unroll
            for (int i = 0; i < m; ++i) ur[i] = (real)io.ur[inst * m + i];
#pragma unroll
            for (int j = 0; j < n; ++j) {
                real b = real(0);
#pragma unroll
                for (int i = 0; i < n; ++i) b = A::nmsub(b, AB[j][i], x0[i]);
                ST(OFF_B + j, b);
                ST(OFF_Q + j, A::mul(C->Q[j], xr[j]));
#if SPCIES_TERMINAL
                ST(OFF_QT + j, A::mul(C->T[j], xr[j]));
#else
                ST(OFF_QT + j, xr[j]);
#endif
            }
#pragma unroll
            for (int j = 0; j < m; ++j) ST(OFF_Q + n + j, A::mul(C->R[j], ur[j]));
            if (VARB) {
#pragma unroll
                for (int j = 0; j < nm; ++j) {
                    ST(OFF_LB + j, (real)io.LB[inst * nm + j]);
                    ST(OFF_UB + j, (real)io.UB[inst * nm + j]);
                }
            }
#pragma unroll 4
            for (int e = 0; e < 2 * N * n; ++e) ST(OFF_Y + e, real(0));   // y = lambda = 0
            k = 0;
            t = real(1);
            warm = true;
        }

        // ================= pass A: z(y) -> residual -> exit test -> forward substitution =================
        real q[nm], yl[n], yn[n], zp[nm], zc[nm], mu[n], mprev[n], u0[m];
#if SPCIES_MU_REGS
        real mus[N - 1][n];   // forward-substitution result of stages 0..N-2 (stage N-1 stays in `mu`)
#endif
        bool over = false;
#pragma unroll
        for (int j = 0; j < nm; ++j) q[j] = LD(OFF_Q + j);
#pragma unroll
        for (int j = 0; j < n; ++j) {
            yl[j] = LD(OFF_Y + j);
            yn[j] = LD(OFF_Y + n + j);
        }
        // stage 0: z_0 (first m decision variables), z[0], r_0            :474-491, :549-554
#pragma unroll
        for (int j = 0; j < m; ++j) {
            real v = q[n + j];
#pragma unroll
            for (int i = 0; i < n; ++i) v = A::nmsub(v, AB[i][n + j], yl[i]);
            u0[j] = clip(A::mul(v, C->QRi[n + j]), bd.lb0(j), bd.ub0(j));
        }
        z_stage<A, VARB>(zc, AB, yl, yn, q, C, bd, 0);
#pragma unroll
        for (int j = 0; j < n; ++j) {
            real r = A::add(LD(OFF_B + j), zc[j]);
#pragma unroll
            for (int i = 0; i < m; ++i) r = A::nmsub(r, AB[j][n + i], u0[i]);
            over |= exceeds(r, tol_);
            mu[j] = r;
        }
        fwd_block<A>(mu, mprev, C, 0, true);
#pragma unroll
        for (int j = 0; j < n; ++j) {
#if SPCIES_MU_REGS
            mus[0][j] = mu[j];
#else
            ST(OFF_MU + j, mu[j]);
#endif
            mprev[j] = mu[j];
            yl[j] = yn[j];
        }
#pragma unroll
        for (int j = 0; j < nm; ++j) zp[j] = zc[j];

        // stages 1 .. N-2                                                  :494-519, :557-564, :593-603
#pragma unroll UNROLL_STAGES
        for (int l = 1; l < N - 1; ++l) {
            SPCIES_STAGE_FENCE();
#pragma unroll
            for (int j = 0; j < n; ++j) yn[j] = LD(OFF_Y + (l + 1) * n + j);
            z_stage<A, VARB>(zc, AB, yl, yn, q, C, bd, l);
#pragma unroll
            for (int j = 0; j < n; ++j) {
                SPCIES_STAGE_FENCE();
                real r = zc[j];
#pragma unroll
                for (int i = 0; i < nm; ++i) r = A::nmsub(r, AB[j][i], zp[i]);
                over |= exceeds(r, tol_);
                mu[j] = r;
            }
            fwd_block<A>(mu, mprev, C, l, false);
#pragma unroll
            for (int j = 0; j < n; ++j) {
#if SPCIES_MU_REGS
                mus[l][j] = mu[j];
#else
                ST(OFF_MU + l * n + j, mu[j]);
#endif
                mprev[j] = mu[j];
                yl[j] = yn[j];
            }
#pragma unroll
            for (int j = 0; j < nm; ++j) zp[j] = zc[j];
        }

        // stage N-1: terminal block                                        :522-537, :567-572, :606-614
#pragma unroll
        for (int j = 0; j < n; ++j) {
#if SPCIES_TERMINAL
            real zN = A::add(LD(OFF_QT + j), yl[j]);
            zN = clip(A::mul(zN, C->Ti[j]), bd.lbN(j), bd.ubN(j));
            real r = zN;
#else
            real r = LD(OFF_QT + j);   // xr                            code_equMPC_FISTA_C.c:549
#endif
#pragma unroll
            for (int i = 0; i < nm; ++i) r = A::nmsub(r, AB[j][i], zp[i]);
            over |= exceeds(r, tol_);
            mu[j] = r;
        }
        fwd_block<A>(mu, mprev, C, N - 1, false);

        // ================= exit condition                                  :337-361 =================
        if (!warm) {
            k += 1;
            int ef = 0;
            if (!over) ef = 1;
            else if (k >= k_max) ef = -1;
            if (ef != 0) {
#pragma unroll
                for (int j = 0; j < m; ++j) io.u[inst * m + j] = (double)u0[j];
                io.k[inst] = k;
                io.e[inst] = ef;
                stat_k += (unsigned long long)k;
                stat_nc += (ef < 0);
                if (io.sol) {
                    // debug payload of sol_<name>: z (recomputed from y exactly as above) and lambda = y  :413-447
                    double *s = io.sol + inst * (long long)(sizeof(SPCIES_SOL_T) / sizeof(double));
                    real a[n], c[n], zz[nm];
#pragma unroll
                    for (int j = 0; j < m; ++j) s[j] = (double)u0[j];
#pragma unroll
                    for (int j = 0; j < n; ++j) a[j] = LD(OFF_Y + j);
                    for (int l = 0; l < N - 1; ++l) {
#pragma unroll
                        for (int j = 0; j < n; ++j) c[j] = LD(OFF_Y + (l + 1) * n + j);
                        z_stage<A, VARB>(zz, AB, a, c, q, C, bd, l);
#pragma unroll
                        for (int j = 0; j < nm; ++j) s[m + l * nm + j] = (double)zz[j];
#pragma unroll
                        for (int j = 0; j < n; ++j) a[j] = c[j];
                    }
                    constexpr int ZLEN = TERMINAL ? N * nm : N * nm - n;
#if SPCIES_TERMINAL
#pragma unroll
                    for (int j = 0; j < n; ++j) {
                        real zN = A::add(LD(OFF_QT + j), a[j]);
                        s[m + (N - 1) * nm + j] = (double)clip(A::mul(zN, C->Ti[j]), bd.lbN(j), bd.ubN(j));
                    }
#endif
                    for (int e = 0; e < N * n; ++e) s[ZLEN + e] = (double)LD(OFF_Y + e);
                    for (int e = ZLEN + N * n; e < (int)(sizeof(SPCIES_SOL_T) / sizeof(double)); ++e) s[e] = 0.0;
                }
                inst = -1;
                continue;
            }
        }

        // ================= pass B: backward substitution, lambda and y updates   :368-385, :619-648 =================
        const real t1 = t;
        if (!warm) t = A::mul(real(0.5), A::add(real(1), A::sqrt(A::add(real(1), A::mul(A::mul(real(4), t1), t1)))));
        const real coef = A::sub(t1, real(1));
        const real beta = EXACT ? real(0) : A::div(coef, t);
        real mnext[n];
        bwd_block<A>(mu, mnext, C, N - 1, true);
        {
            const int l = N - 1;
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const real lam1 = LD(OFF_LAM + l * n + j);
                const real lam = A::add(LD(OFF_Y + l * n + j), mu[j]);
                const real d = A::sub(lam, lam1);
                const real ynew = EXACT ? A::add(lam, A::div(A::mul(coef, d), t)) : A::madd(lam, beta, d);
                ST(OFF_LAM + l * n + j, lam);
                ST(OFF_Y + l * n + j, ynew);
                mnext[j] = mu[j];
            }
        }
#pragma unroll UNROLL_STAGES
        for (int l = N - 2; l >= 0; --l) {
            SPCIES_STAGE_FENCE();
#pragma unroll
#if SPCIES_MU_REGS
            for (int j = 0; j < n; ++j) mu[j] = mus[l][j];
#else
            for (int j = 0; j < n; ++j) mu[j] = LD(OFF_MU + l * n + j);
#endif
            bwd_block<A>(mu, mnext, C, l, false);
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const real lam1 = LD(OFF_LAM + l * n + j);
                const real lam = A::add(LD(OFF_Y + l * n + j), mu[j]);
                const real d = A::sub(lam, lam1);
                const real ynew = EXACT ? A::add(lam, A::div(A::mul(coef, d), t)) : A::madd(lam, beta, d);
                ST(OFF_LAM + l * n + j, lam);
                ST(OFF_Y + l * n + j, ynew);
                mnext[j] = mu[j];
            }
        }
        warm = false;
    }
    flush_stats(io.queue, stat_k, stat_nc);
}

struct Traits {
    static constexpr int NN = n, MM = m, NMM = nm;
    static constexpr bool HAS_R = false;
    static constexpr bool HAS_VARB = true;
    static constexpr int SOL_DOUBLES = (int)(sizeof(SPCIES_SOL_T) / sizeof(double));
    typedef spcies_consts Consts;
    static const Consts &host_consts() { return spcies_h_consts; }
    static int default_block(bool varb) { return varb ? BLOCK_VARB : BLOCK_FIXED; }
    static size_t smem_bytes(int block, bool varb) {
        return CONSTS_BYTES + (size_t)(varb ? STATE_VARB : STATE_FIXED) * block * sizeof(real);
    }
    template <bool EXACT, bool VARB, int BLOCK>
    static cudaError_t launch_t(int grid, size_t smem, cudaStream_t s, const BatchIO &io, const void *dc) {
        auto kern = fista_kernel<EXACT, VARB, BLOCK>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, BLOCK, smem, s>>>(io, (const spcies_consts *)dc);
        return cudaGetLastError();
    }
    static size_t scratch_bytes(int, int, bool) { return 0; }
    static cudaError_t launch(int arith, bool varb, int grid, int block, size_t smem, cudaStream_t s, const BatchIO &io,
                              const void *dc, void *) {
        const bool ex = arith == SPCIES_CUDA_ARITH_EXACT;
        if (varb) {
            if (block != BLOCK_VARB) return cudaErrorInvalidConfiguration;
            return ex ? launch_t<true, true, BLOCK_VARB>(grid, smem, s, io, dc)
                      : launch_t<false, true, BLOCK_VARB>(grid, smem, s, io, dc);
        }
        if (block != BLOCK_FIXED) return cudaErrorInvalidConfiguration;
        return ex ? launch_t<true, false, BLOCK_FIXED>(grid, smem, s, io, dc)
                  : launch_t<false, false, BLOCK_FIXED>(grid, smem, s, io, dc);
    }
    static cudaError_t attributes(int arith, bool varb, cudaFuncAttributes *a) {
        const bool ex = arith == SPCIES_CUDA_ARITH_EXACT;
        if (varb)
            return ex ? cudaFuncGetAttributes(a, fista_kernel<true, true, BLOCK_VARB>)
                      : cudaFuncGetAttributes(a, fista_kernel<false, true, BLOCK_VARB>);
        return ex ? cudaFuncGetAttributes(a, fista_kernel<true, false, BLOCK_FIXED>)
                  : cudaFuncGetAttributes(a, fista_kernel<false, false, BLOCK_FIXED>);
    }
};

}  // namespace fista
}  // namespace spcies

#define SPCIES_TRAITS ::spcies::fista::Traits
#include "spcies_entry.cuh"
