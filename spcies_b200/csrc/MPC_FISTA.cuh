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
//   * the shared problem constants (AB, Alpha, Beta, QRi, bounds, the derived FAST-mode blocks) are staged once
//     per CTA into shared memory with one bulk async copy; AB is then held in registers for the whole kernel.
//   * one FISTA iteration = two fused sweeps over the horizon
//        pass A (l = 0..N-1):  z_l(y) -> r_l -> exit test -> forward step -> w_l
//        pass B (l = N-1..0):  backward step d_lambda_l -> lambda_l = y_l + d_lambda_l -> y_l update
//     so z and the residual are never materialised (the reference stores z[N-1][nm], d_lambda[N][n]).
//     The warm-up step of the reference (:300-320) is the same two passes with the exit test off.
//   * per-instance iterates: y[N][n] and the vectors b, q, qT [, LB, UB] in shared memory, [element][thread];
//     lambda[N][n] and the solve workspace w[N][n] either
//        TM = true    in Tensor Memory, as private per-lane rows (spcies_tmem.cuh): 640 B instead of 1600 B of shared
//                     memory per instance at N = 10, i.e. 8 instead of 4 resident warps per SM -- the throughput
//                     configuration (single launch, or first launch of two);
//        TM = false   in shared memory as well: lowest latency per warp -- the configuration of the second (tail)
//                     launch, and the fallback when the iterates do not fit TMEM.
//     Everything with a stage-local lifetime (z_l, r_l, mu_{l-1}) lives in registers.
//
// tcgen05.ld / st are warp-collective, so the iteration has no divergent control flow around the row accesses: a lane
// that has finished its instance (or found the queue empty) stays in the warp with `live = false` and computes on
// stale data until the next refill point; results are only written for live lanes.  lambda needs no zero-
// initialisation: the warm-up pass (`warm`) selects lambda = 0 instead of the loaded value.
//
// Arithmetic
//   EXACT  the reference's block forward / backward substitutions, operation for operation: bit-identical to the
//          reference C compiled by gcc -O3 (no contraction).
//   FAST   FMA contraction, 1/t hoisted, and the same linear solve W mu = r written with explicit inverses of the
//          n x n triangular diagonal blocks and pre-multiplied coupling blocks (FistaDerived, computed once on the host
//          in extended precision from the generated Alpha / Beta):
//              forward   mu_l = Linv_l r_l - F_l mu_{l-1},        F_l = Linv_l Alpha_{l-1}^T
//              backward  dl_l = Uinv_l mu_l - G_l dl_{l+1},        G_l = Uinv_l Alpha_l
//          Same FMA count (57 + 57 per stage at n = 6), but every product is a dense mat-vec with n independent
//          accumulators: the 11-deep dependent chain of a triangular substitution (~90 cycles of DFMA latency per
//          block, twice per stage) and the diagonal DMULs disappear, and w_l = Uinv_l mu_l is computed in pass A, where
//          it overlaps with the z / residual products, so pass B is one n x n mat-vec per stage.  Results differ from
//          the reference by rounding only (gate: u_opt <= 1e-9 relative, e_flag identical, |dk| <= 1; measured 8e-15).
//
// Tail handling (io.phase, spcies_common.cuh)
//   Iteration counts are heavy-tailed (C2: mean 32, 0.6 % of the instances run to k_max = 1000 and hold 19 % of the
//   work).  When the queue runs dry every warp still holds a few of those and keeps iterating at the full-occupancy
//   per-warp rate.  With io.phase = 1 a lane whose instance is still running io.grace iterations after the queue ran
//   dry parks it (k, t, y, lambda -> global memory) and the launch ends; a second launch (io.phase = 2, TM = false,
//   one warp per scheduler) resumes the parked instances at the single-warp rate.  Parking copies the iterates
//   verbatim, so results do not depend on it (bit-identical in EXACT mode).
//
// The including .cu (emitted by platforms/cuda_code.py) defines SPCIES_REAL, SPCIES_TERMINAL, SPCIES_SOL_T,
// SPCIES_FUNC, the reference #defines (nn_, mm_, nm_, NN_, k_max, tol [, VAR_BOUNDS]) and
// `struct spcies_consts` + `spcies_h_consts` holding LB, UB, AB, Alpha, Beta, Q, R, QRi [, T, Ti].
#pragma once
#include "spcies_host.cuh"
#include "spcies_tmem.cuh"
#include "spcies_mma.cuh"
#include "spcies_dense_mma.cuh"

// Compile-time switches (set by the generator / tools/variants.py)
#ifndef SPCIES_FISTA_TMEM
#define SPCIES_FISTA_TMEM 1          // 0: never use Tensor Memory (all iterates in shared memory)
#endif
#ifndef SPCIES_FISTA_MAXBLOCK
#define SPCIES_FISTA_MAXBLOCK 256    // threads per CTA of the throughput configuration (255 registers per thread at 256)
#endif
#ifndef SPCIES_FISTA_UNROLL
#define SPCIES_FISTA_UNROLL 2        // unroll factor of the stage loops (2: ping-pong registers instead of moves, +4 %)
#endif
#ifndef SPCIES_FISTA_BLOCK2
#define SPCIES_FISTA_BLOCK2 128      // threads per CTA of the tail launch (one warp per scheduler)
#endif
#ifndef SPCIES_FISTA_TAIL_TMEM
#define SPCIES_FISTA_TAIL_TMEM 0     // 1: the tail launch keeps lambda / w in Tensor Memory as well (fewer shared-memory wavefronts)
#endif
#ifndef SPCIES_FISTA_CBANK
#define SPCIES_FISTA_CBANK 0         // 1: QRi / LB / UB operands through the constant bank instead of shared memory
#endif

namespace spcies {
namespace fista {

typedef SPCIES_REAL real;
constexpr int n = nn_, m = mm_, nm = nm_, N = NN_;
constexpr bool TERMINAL = (SPCIES_TERMINAL != 0);
constexpr int FISTA_UNROLL = SPCIES_FISTA_UNROLL;

#if SPCIES_FISTA_CBANK
__constant__ spcies_consts c_consts;
#endif

// ---------------------------------------------------------------------------------------------------------------------
// Derived constants of the FAST arithmetic
// ---------------------------------------------------------------------------------------------------------------------
struct alignas(16) FistaDerived {
    real Linv[N][n][n];   // (U_l^T)^-1 (lower triangular); U_l = upper factor block: U[i][j] = Beta[l][i][j], U[j][j] = 1 / Beta[l][j][j]
    real F[N][n][n];      // Linv_l * Alpha_{l-1}^T, l >= 1
    real Uinv[N][n][n];   // U_l^-1 (upper triangular)
    real G[N][n][n];      // Uinv_l * Alpha_l, l <= N-2
};

static inline void compute_derived(const spcies_consts &C, FistaDerived &D) {
    typedef long double ld;
    memset(&D, 0, sizeof D);
    for (int l = 0; l < N; ++l) {
        ld U[n][n] = {}, Ui[n][n] = {};
        for (int i = 0; i < n; ++i)
            for (int j = i; j < n; ++j) U[i][j] = (i == j) ? (ld)1 / (ld)C.Beta[l][j][j] : (ld)C.Beta[l][i][j];
        for (int c = 0; c < n; ++c)            // U * Ui[:, c] = e_c by back substitution
            for (int i = n - 1; i >= 0; --i) {
                ld v = (i == c) ? (ld)1 : (ld)0;
                for (int j = i + 1; j < n; ++j) v -= U[i][j] * Ui[j][c];
                Ui[i][c] = v / U[i][i];
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                D.Uinv[l][i][j] = (j >= i) ? (real)Ui[i][j] : real(0);
                D.Linv[l][i][j] = (j <= i) ? (real)Ui[j][i] : real(0);
            }
        if (l >= 1)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    ld v = 0;
                    for (int k = 0; k <= j; ++k) v += Ui[k][j] * (ld)C.Alpha[l - 1][i][k];   // Linv[j][k] * Alpha^T[k][i]
                    D.F[l][j][i] = (real)v;
                }
        if (l <= N - 2)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    ld v = 0;
                    for (int k = j; k < n; ++k) v += Ui[j][k] * (ld)C.Alpha[l][k][i];
                    D.G[l][j][i] = (real)v;
                }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Sizes and layouts
// ---------------------------------------------------------------------------------------------------------------------
constexpr size_t CONSTS_BYTES = (sizeof(spcies_consts) + 15) / 16 * 16;
constexpr size_t DERIVED_BYTES = (sizeof(FistaDerived) + 15) / 16 * 16;
constexpr size_t BLOB_BYTES = CONSTS_BYTES + DERIVED_BYTES;   // one bulk copy global -> shared per CTA
constexpr size_t SMEM_MAX = 227 * 1024;

// shared-memory state elements per instance ([element][thread])
constexpr int OFF_Y = 0;                 // y[N][n]      linearisation point
constexpr int OFF_B = OFF_Y + N * n;     // b[n] = -A x0
constexpr int OFF_Q = OFF_B + n;         // q[nm] = [Q xr; R ur]   (Q, R stored negated)
constexpr int OFF_QT = OFF_Q + nm;       // qT[n] = T xr (lax)  |  xr (equ)
constexpr int OFF_LB = OFF_QT + n;       // per-instance bounds (VARB kernels only)
constexpr int OFF_UB = OFF_LB + nm;
__host__ __device__ constexpr int state_common(bool varb) { return varb ? OFF_UB + nm : OFF_LB; }
// lambda[N][n] and w[N][n] follow when they are not in TMEM
__host__ __device__ constexpr int state_elems(bool varb, bool tm) { return state_common(varb) + (tm ? 0 : 2 * N * n); }

constexpr int WPR = (int)(sizeof(real) / 4);                 // 32-bit TMEM columns per real
constexpr int ROW_LAM = 0, ROW_W = 1;                        // the two row arrays
constexpr int TM_COLS = 2 * N * n * WPR;                     // TMEM columns per thread

constexpr int fit_threads(int elems) {
    int t = (int)((SMEM_MAX - BLOB_BYTES - 64) / ((size_t)elems * sizeof(real))) / 32 * 32;
    return t > 1024 ? 1024 : t;
}
constexpr int block_tm(bool varb) {
    int t = fit_threads(state_elems(varb, true));
    if (t > SPCIES_FISTA_MAXBLOCK) t = SPCIES_FISTA_MAXBLOCK;
    if (t > 128 && 2 * TM_COLS > 512) t = 128;               // two warps per TMEM lane quadrant need 2 x TM_COLS columns
    return t;
}
constexpr int block_sm(bool varb) {
    int t = fit_threads(state_elems(varb, false));
    return t > 256 ? 256 : t;
}
constexpr bool USE_TMEM = (SPCIES_FISTA_TMEM != 0) && TM_COLS <= 512 && block_tm(true) > block_sm(true);
static_assert(block_sm(true) >= 32 || USE_TMEM, "per-instance state does not fit on chip with one thread per instance");
// throughput configuration (single launch, or first launch of two)
constexpr int BLOCK1_FIXED = USE_TMEM ? block_tm(false) : block_sm(false);
constexpr int BLOCK1_VARB = USE_TMEM ? block_tm(true) : block_sm(true);
// tail configuration (second launch): all iterates in shared memory when that fits, at most BLOCK2 threads
constexpr bool TAIL_TMEM = USE_TMEM && (block_sm(true) < 32 || SPCIES_FISTA_TAIL_TMEM != 0);
constexpr int block_tail(bool varb) {
    int t = TAIL_TMEM ? block_tm(varb) : block_sm(varb);
    return t > SPCIES_FISTA_BLOCK2 ? SPCIES_FISTA_BLOCK2 : t;
}
constexpr int BLOCK2_FIXED = block_tail(false);
constexpr int BLOCK2_VARB = block_tail(true);
constexpr int PARK_DOUBLES = 3 + 2 * N * n;                  // inst, k, t, y[N][n], lambda[N][n]

// ---------------------------------------------------------------------------------------------------------------------
// Building blocks
// ---------------------------------------------------------------------------------------------------------------------
#if SPCIES_FISTA_CBANK
#define SPCIES_CB(member) c_consts.member
#else
#define SPCIES_CB(member) C->member
#endif
template <bool VARB> struct Bounds {
    // stage bounds: u_0 (j in [0,m)), stage l (j in [0,nm)), terminal x_N (j in [0,n))
    const spcies_consts *C;
    const real *st;  // state base (VARB)
    int stride;
#ifdef VAR_BOUNDS
    __device__ __forceinline__ real lb0(int j) const { return VARB ? st[(OFF_LB + n + j) * stride] : SPCIES_CB(LB0[j]); }
    __device__ __forceinline__ real ub0(int j) const { return VARB ? st[(OFF_UB + n + j) * stride] : SPCIES_CB(UB0[j]); }
    __device__ __forceinline__ real lb(int l, int j) const { return VARB ? st[(OFF_LB + j) * stride] : C->LB[l][j]; }
    __device__ __forceinline__ real ub(int l, int j) const { return VARB ? st[(OFF_UB + j) * stride] : C->UB[l][j]; }
#if SPCIES_TERMINAL
    __device__ __forceinline__ real lbN(int j) const { return VARB ? st[(OFF_LB + j) * stride] : SPCIES_CB(LBN[j]); }
    __device__ __forceinline__ real ubN(int j) const { return VARB ? st[(OFF_UB + j) * stride] : SPCIES_CB(UBN[j]); }
#endif
#else
    __device__ __forceinline__ real lb0(int j) const { return VARB ? st[(OFF_LB + n + j) * stride] : SPCIES_CB(LB[n + j]); }
    __device__ __forceinline__ real ub0(int j) const { return VARB ? st[(OFF_UB + n + j) * stride] : SPCIES_CB(UB[n + j]); }
    __device__ __forceinline__ real lb(int, int j) const { return VARB ? st[(OFF_LB + j) * stride] : SPCIES_CB(LB[j]); }
    __device__ __forceinline__ real ub(int, int j) const { return VARB ? st[(OFF_UB + j) * stride] : SPCIES_CB(UB[j]); }
#if SPCIES_TERMINAL
    __device__ __forceinline__ real lbN(int j) const { return VARB ? st[(OFF_LB + j) * stride] : SPCIES_CB(LB[j]); }
    __device__ __forceinline__ real ubN(int j) const { return VARB ? st[(OFF_UB + j) * stride] : SPCIES_CB(UB[j]); }
#endif
#endif
    __device__ __forceinline__ real qri(int j) const { return SPCIES_CB(QRi[j]); }
};

// z_l = clip(QRi o (q - [A B]' y_{l+1} + [y_l; 0]))            code_laxMPC_FISTA_C.c:494-519
template <class A, bool VARB>
__device__ __forceinline__ void z_stage(real (&z)[nm], const real (&AB)[n][nm], const real (&yl)[n],
                                        const real (&yn)[n], const real (&q)[nm], const Bounds<VARB> &bd, int l) {
#pragma unroll
    for (int j = 0; j < nm; ++j) z[j] = q[j];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < nm; ++j) z[j] = A::nmsub(z[j], AB[i][j], yn[i]);
#pragma unroll
    for (int j = 0; j < n; ++j) z[j] = A::add(z[j], yl[j]);
#pragma unroll
    for (int j = 0; j < nm; ++j) z[j] = clip(A::mul(z[j], bd.qri(j)), bd.lb(l, j), bd.ub(l, j));
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

// Forward step of stage l.  In: mu = r_l (residual block).  Out: mu (forward-substitution result, input of the next
// stage) and w (what pass B needs from this stage: EXACT w = mu; FAST w = Uinv_l mu).
template <class A, bool EXACT>
__device__ __forceinline__ void fwd_stage(real (&mu)[n], real (&w)[n], const real (&mprev)[n], const spcies_consts *C,
                                          const FistaDerived *D, int l, bool first) {
    if constexpr (EXACT) {
        fwd_block<A>(mu, mprev, C, l, first);
#pragma unroll
        for (int j = 0; j < n; ++j) w[j] = mu[j];
    } else {
        real r[n];
#pragma unroll
        for (int j = 0; j < n; ++j) r[j] = mu[j];
#pragma unroll
        for (int j = 0; j < n; ++j) {
            real s = D->Linv[l][j][0] * r[0];
#pragma unroll
            for (int i = 1; i <= j; ++i) s = fma(D->Linv[l][j][i], r[i], s);
            mu[j] = s;
        }
        if (!first) {
#pragma unroll
            for (int j = 0; j < n; ++j)
#pragma unroll
                for (int i = 0; i < n; ++i) mu[j] = fma(-D->F[l][j][i], mprev[i], mu[j]);
        }
#pragma unroll
        for (int j = 0; j < n; ++j) {
            real s = D->Uinv[l][j][j] * mu[j];
#pragma unroll
            for (int i = j + 1; i < n; ++i) s = fma(D->Uinv[l][j][i], mu[i], s);
            w[j] = s;
        }
    }
}

// Backward step of stage l.  In: w (from pass A), dnext = d_lambda_{l+1}.  Out: w overwritten with d_lambda_l.
template <class A, bool EXACT>
__device__ __forceinline__ void bwd_stage(real (&w)[n], const real (&dnext)[n], const spcies_consts *C, const FistaDerived *D,
                                          int l, bool last) {
    if constexpr (EXACT) {
        bwd_block<A>(w, dnext, C, l, last);
    } else {
        if (!last) {
#pragma unroll
            for (int j = 0; j < n; ++j)
#pragma unroll
                for (int i = 0; i < n; ++i) w[j] = fma(-D->G[l][j][i], dnext[i], w[j]);
        }
    }
}

// lambda[N][n] and w[N][n]: private rows in TMEM (TM) or [element][thread] in shared memory.  Loads are split into
// issue / wait / get so that the TMEM flavour can prefetch; the shared-memory flavour loads at get().
template <bool TM, int BLOCK> struct Rows {
    typedef tmem::Vec<real, n> TV;
    uint32_t trow;    // TM: address of column 0 of this thread's row
    real *sbase;      // !TM: &lambda[0][0] of this thread
    struct Pending {
        TV v;
        const real *p;
    };
    __device__ __forceinline__ void issue(Pending &pd, int arr, int l) const {
        if constexpr (TM) pd.v.issue(trow + (uint32_t)((arr * N + l) * (n * WPR)));
        else pd.p = sbase + (size_t)((arr * N + l) * n) * BLOCK;
    }
    __device__ __forceinline__ void wait() const {
        if constexpr (TM) tmem::wait_ld();
    }
    __device__ __forceinline__ void get(Pending &pd, real (&out)[n]) const {
        if constexpr (TM) pd.v.get(out);
        else {
#pragma unroll
            for (int j = 0; j < n; ++j) out[j] = pd.p[j * BLOCK];
        }
    }
    __device__ __forceinline__ void store(int arr, int l, const real (&v)[n]) const {
        if constexpr (TM) TV::store(trow + (uint32_t)((arr * N + l) * (n * WPR)), v);
        else {
            real *p = sbase + (size_t)((arr * N + l) * n) * BLOCK;
#pragma unroll
            for (int j = 0; j < n; ++j) p[j * BLOCK] = v[j];
        }
    }
    __device__ __forceinline__ void wait_st() const {
        if constexpr (TM) tmem::wait_st();
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------------------------------
template <bool EXACT, bool VARB, int BLOCK, bool TM>
__global__ void __launch_bounds__(BLOCK, 1) fista_kernel(const BatchIO io, const spcies_consts *__restrict__ g_consts) {
    typedef Arith<real, EXACT> A;
    typedef Rows<TM, BLOCK> RowsT;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    spcies_consts *C = reinterpret_cast<spcies_consts *>(smem_raw);
    const FistaDerived *D = reinterpret_cast<const FistaDerived *>(smem_raw + CONSTS_BYTES);
    stage_constants(C, g_consts, (uint32_t)BLOB_BYTES, &mbar);

    real *stbase = reinterpret_cast<real *>(smem_raw + BLOB_BYTES) + threadIdx.x;
    auto LD = [&](int e) -> real { return stbase[e * BLOCK]; };
    auto ST = [&](int e, real v) { stbase[e * BLOCK] = v; };
    Bounds<VARB> bd{C, stbase, BLOCK};
    RowsT rows;
    uint32_t tbase = 0;
    if constexpr (TM) {
        tbase = tmem::alloc_all(&tmem_slot);
        rows.trow = tmem::my_row(tbase, BLOCK > 128 ? 256 : 512);
        rows.sbase = nullptr;
    } else {
        rows.trow = 0;
        constexpr int COMMON = state_common(VARB);
        rows.sbase = stbase + (size_t)COMMON * BLOCK;
    }

    // [A B] is reused by every stage of every iteration: pinned in registers
    real AB[n][nm];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < nm; ++j) AB[i][j] = C->AB[i][j];

    const real tol_ = (real)tol;
    const bool resume = io.phase == 2;
    const long long B = resume ? (long long)io.queue[6] : io.B;
    const WorkQueue wq{resume ? io.queue + 7 : io.queue, B, resume ? nullptr : io.ready};
    const WorkQueue marks{io.queue, B, nullptr};
    marks.mark_start();
    unsigned long long stat_k = 0;
    unsigned int stat_nc = 0;

    long long inst = -1, rslot = 0;
    int k = 0;
    int grace_left = io.grace;
    bool live = false, drained = false, warm = false, fresh = false;
    real t = real(1);

    for (;;) {
        // ---- refill: a lane without an instance pulls the next one                code_laxMPC_FISTA_C.c:94-100, 275-289
        if (!live && !drained) {
            const long long slot = wq.next();
            if (slot < 0) {
                drained = true;
                marks.mark_drained();
            } else {
                const double *pk = io.park + slot;
                inst = resume ? __double_as_longlong(pk[0]) : slot;
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
                if (resume) {
                    // parked instance: k, t, y from its record (lambda follows below, through a collective row store)
                    k = (int)__double_as_longlong(pk[1 * io.park_cap]);
                    t = (real)pk[2 * io.park_cap];
#pragma unroll 4
                    for (int e = 0; e < N * n; ++e) ST(OFF_Y + e, (real)pk[(3 + e) * io.park_cap]);
                    warm = false;
                } else {
#pragma unroll 4
                    for (int e = 0; e < N * n; ++e) ST(OFF_Y + e, real(0));   // y = 0 (lambda = 0 through `warm`)
                    k = 0;
                    t = real(1);
                    warm = true;
                }
                live = true;
                fresh = resume;
                rslot = slot;
            }
        }
        __syncwarp();
        if (resume && __any_sync(FULL, fresh)) {
            // lambda of the lanes that just resumed.  Row stores are warp-collective in TMEM: the other lanes write back
            // the rows they hold (loaded first).
            rows.wait_st();
#pragma unroll 1
            for (int l = 0; l < N; ++l) {
                typename RowsT::Pending cur;
                rows.issue(cur, ROW_LAM, l);
                rows.wait();
                real v[n];
                rows.get(cur, v);
                if (fresh) {
#pragma unroll
                    for (int j = 0; j < n; ++j) v[j] = (real)io.park[(3 + N * n + l * n + j) * io.park_cap + rslot];
                }
                rows.store(ROW_LAM, l, v);
            }
            rows.wait_st();
            fresh = false;
        }
        if (!__any_sync(FULL, live)) break;
        // phase 1: has the queue run dry (for anybody)?  One broadcast load per warp and iteration.
        const bool dry = (io.phase == 1) && (*(volatile unsigned long long *)io.queue >= (unsigned long long)B);

        // ================= pass A: z(y) -> residual -> exit test -> forward step =================
        real q[nm], yl[n], yn[n], zp[nm], zc[nm], mu[n], w[n], mprev[n], u0[m];
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
            u0[j] = clip(A::mul(v, bd.qri(n + j)), bd.lb0(j), bd.ub0(j));
        }
        z_stage<A, VARB>(zc, AB, yl, yn, q, bd, 0);
#pragma unroll
        for (int j = 0; j < n; ++j) {
            real r = A::add(LD(OFF_B + j), zc[j]);
#pragma unroll
            for (int i = 0; i < m; ++i) r = A::nmsub(r, AB[j][n + i], u0[i]);
            over |= exceeds(r, tol_);
            mu[j] = r;
        }
        fwd_stage<A, EXACT>(mu, w, mprev, C, D, 0, true);
        rows.store(ROW_W, 0, w);
#pragma unroll
        for (int j = 0; j < n; ++j) {
            mprev[j] = mu[j];
            yl[j] = yn[j];
        }
#pragma unroll
        for (int j = 0; j < nm; ++j) zp[j] = zc[j];

        // stages 1 .. N-2                                                  :494-519, :557-564, :593-603
#pragma unroll FISTA_UNROLL
        for (int l = 1; l < N - 1; ++l) {
#pragma unroll
            for (int j = 0; j < n; ++j) yn[j] = LD(OFF_Y + (l + 1) * n + j);
            z_stage<A, VARB>(zc, AB, yl, yn, q, bd, l);
#pragma unroll
            for (int j = 0; j < n; ++j) {
                real r = zc[j];
#pragma unroll
                for (int i = 0; i < nm; ++i) r = A::nmsub(r, AB[j][i], zp[i]);
                over |= exceeds(r, tol_);
                mu[j] = r;
            }
            fwd_stage<A, EXACT>(mu, w, mprev, C, D, l, false);
            rows.store(ROW_W, l, w);
#pragma unroll
            for (int j = 0; j < n; ++j) {
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
        fwd_stage<A, EXACT>(mu, w, mprev, C, D, N - 1, false);

        // ================= exit condition                                  :337-361 =================
        if (live && !warm) {
            k += 1;
            int ef = 0;
            if (!over) ef = 1;
            else if (k >= k_max) ef = -1;
            if (ef != 0) {
#pragma unroll
                for (int j = 0; j < m; ++j) io.u[inst * m + j] = eng_u_out(C, (double)u0[j], j);
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
                        z_stage<A, VARB>(zz, AB, a, c, q, bd, l);
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
                live = false;
            }
        }
        __syncwarp();

        // ================= pass B: backward step, lambda and y updates          :368-385, :619-648 =================
        const real t1 = t;
        if (!warm) t = A::mul(real(0.5), A::add(real(1), A::sqrt(A::add(real(1), A::mul(A::mul(real(4), t1), t1)))));
        const real coef = A::sub(t1, real(1));
        const real beta = EXACT ? real(0) : A::div(coef, t);
        real dnext[n], lam1[n], wn[n];
        typename RowsT::Pending p_lam, p_w;
        rows.wait_st();                                   // w rows of pass A, lambda rows of the previous pass B
        rows.issue(p_lam, ROW_LAM, N - 1);
        if (N > 1) rows.issue(p_w, ROW_W, N - 2);
        auto update = [&](int l) {                         // lambda_l = y_l + d_lambda_l;  y_l <- lambda_l + beta (lambda_l - lambda1_l)
            rows.wait();
            rows.get(p_lam, lam1);
            if (l > 0) rows.get(p_w, wn);
            real lamv[n];
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const real l1 = warm ? real(0) : lam1[j];
                const real lam = A::add(LD(OFF_Y + l * n + j), w[j]);
                const real d = A::sub(lam, l1);
                const real ynew = EXACT ? A::add(lam, A::div(A::mul(coef, d), t)) : A::madd(lam, beta, d);
                lamv[j] = lam;
                ST(OFF_Y + l * n + j, ynew);
                dnext[j] = w[j];
                w[j] = wn[j];
            }
            rows.store(ROW_LAM, l, lamv);
            if (l > 0) {
                rows.issue(p_lam, ROW_LAM, l - 1);
                if (l > 1) rows.issue(p_w, ROW_W, l - 2);
            }
        };
        bwd_stage<A, EXACT>(w, dnext, C, D, N - 1, true);
        update(N - 1);
#pragma unroll FISTA_UNROLL
        for (int l = N - 2; l >= 0; --l) {
            bwd_stage<A, EXACT>(w, dnext, C, D, l, false);
            update(l);
        }

        // ---- phase 1: park an instance that is still running `grace` iterations after the queue ran dry
        bool park_now = false;
        if (dry && live && !warm) {
            park_now = grace_left <= 0;
            grace_left -= 1;
        }
        if (__any_sync(FULL, park_now)) {
            long long pslot = 0;
            if (park_now) {
                pslot = (long long)atomicAdd(io.queue + 6, 1ULL);
                double *pk = io.park + pslot;
                pk[0] = __longlong_as_double(inst);
                pk[1 * io.park_cap] = __longlong_as_double((long long)k);
                pk[2 * io.park_cap] = (double)t;
#pragma unroll 4
                for (int e = 0; e < N * n; ++e) pk[(3 + e) * io.park_cap] = (double)LD(OFF_Y + e);
            }
            __syncwarp();
            rows.wait_st();
#pragma unroll 1
            for (int l = 0; l < N; ++l) {
                typename RowsT::Pending cur;
                rows.issue(cur, ROW_LAM, l);
                rows.wait();
                real v[n];
                rows.get(cur, v);
                if (park_now) {
#pragma unroll
                    for (int j = 0; j < n; ++j) io.park[(3 + N * n + l * n + j) * io.park_cap + pslot] = (double)v[j];
                }
            }
            if (park_now) live = false;
            __syncwarp();
        }
        warm = false;
    }
    flush_stats(io.queue, stat_k, stat_nc);
    marks.mark_end();
    if constexpr (TM) tmem::free_all(tbase);
}

#include "MPC_FISTA_mma.cuh"
#include "MPC_FISTA_single.cuh"
#include "MPC_FISTA_dense.cuh"

struct Traits {
    static constexpr int NN = n, MM = m, NMM = nm;
    static constexpr bool HAS_R = false;
    static constexpr int NREF = 1;
    static constexpr bool TV = false;
    static constexpr int extra_width(int) { return 0; }
    static constexpr bool HAS_VARB = true;
    static constexpr bool HAS_PARK = true;                   // phase-1 / phase-2 launches (park & resume the slow tail)
    static constexpr int PARK_DOUBLES = fista::PARK_DOUBLES;
    static constexpr int K_MAX = k_max;
    static constexpr int SOL_DOUBLES = (int)(sizeof(SPCIES_SOL_T) / sizeof(double));
    typedef spcies_consts Consts;
    static const Consts &host_consts() { return spcies_h_consts; }
    // device constant blob: the generated constants + the derived FAST-mode blocks
    typedef dense::Plan<DenseEngine> DP;
    static size_t blob_bytes() {
        if (HAS_DENSE) return DP::BLOB_BYTES;
        return (HAS_MMA ? TOTAL_BLOB_BYTES : BLOB_BYTES) + (HAS_SINGLE ? SINGLE_BYTES : 0);
    }
    static void fill_blob(void *dst) {
        memset(dst, 0, blob_bytes());
        memcpy(dst, &spcies_h_consts, sizeof spcies_h_consts);
        FistaDerived *D = new FistaDerived;
        compute_derived(spcies_h_consts, *D);
        memcpy((char *)dst + CONSTS_BYTES, D, sizeof *D);
        if constexpr (HAS_MMA) {
            MmaTables *T = new MmaTables;
            fill_mma_tables(spcies_h_consts, *D, *T);
            memcpy((char *)dst + MMA_OFFSET, T, sizeof *T);
            delete T;
        }
        delete D;
        if constexpr (HAS_SINGLE || HAS_DENSE) {
            PrimalForm *PF = new PrimalForm(spcies_h_consts);
            if constexpr (HAS_SINGLE) {
                SingleTables *S = new SingleTables;
                fill_single_tables(*PF, *S);
                memcpy((char *)dst + SINGLE_OFFSET, S, sizeof *S);
                delete S;
            }
            if constexpr (HAS_DENSE) {
                DenseEngine::Small *S = new DenseEngine::Small;
                memset(S, 0, sizeof *S);
                long double *F = new long double[(size_t)DenseEngine::NO * 8 * DP::NIN * 8]();
                DenseEngine::fill(*PF, *S, F);
                dense::fill_fragments<DenseEngine>(F, reinterpret_cast<double2 *>((char *)dst + DP::OFF_FRAG));
                memcpy((char *)dst + DP::OFF_SMALL, S, sizeof *S);
                delete[] F;
                delete S;
            }
            delete PF;
        }
    }
    // latency engine (one CTA per instance) for small host-buffer calls: FAST arithmetic, no debug payload
    static bool single_engine(int arith, const BatchIO &io) {
        if constexpr (!HAS_SINGLE) return false;
        return arith != SPCIES_CUDA_ARITH_EXACT && io.sol == nullptr &&
               (io.engine == SPCIES_CUDA_ENGINE_AUTO || io.engine == SPCIES_CUDA_ENGINE_SINGLE);
    }
    static cudaError_t launch_single(bool varb, int grid, cudaStream_t s, const BatchIO &io, const void *dc, int &block, size_t &smem,
                                     const double *hx0, const double *hxr, const double *hur) {
        if constexpr (HAS_SINGLE) {
            block = SG_BLOCK;
            smem = SINGLE_SMEM;
            SingleArgs a;
            a.count = 0;
            if (hx0 != nullptr && grid <= SG_NARG) {        // the inputs of a few instances travel with the launch
                a.count = grid;
                memcpy(a.x0, hx0, (size_t)grid * n * sizeof(double));
                memcpy(a.xr, hxr, (size_t)grid * n * sizeof(double));
                memcpy(a.ur, hur, (size_t)grid * m * sizeof(double));
            }
            if (varb) fista_single_kernel<true><<<grid, SG_BLOCK, SINGLE_SMEM, s>>>(io, (const unsigned char *)dc, a);
            else fista_single_kernel<false><<<grid, SG_BLOCK, SINGLE_SMEM, s>>>(io, (const unsigned char *)dc, a);
            return cudaGetLastError();
        }
        return cudaErrorNotSupported;
    }
    // the lingering server of the single-instance symbol (MPC_FISTA_single.cuh); mb = device address of the mapped mailbox
    static constexpr bool HAS_SERVER = HAS_SINGLE;
    static cudaError_t launch_server(cudaStream_t s, const void *dc, void *mb, unsigned int seq0, unsigned long long linger_ns, int &block,
                                     size_t &smem) {
        if constexpr (HAS_SINGLE) {
            block = SG_BLOCK;
            smem = SINGLE_SMEM;
            fista_server_kernel<<<1, SG_BLOCK, SINGLE_SMEM, s>>>((const unsigned char *)dc, (Mailbox *)mb, seq0, linger_ns);
            return cudaGetLastError();
        }
        return cudaErrorNotSupported;
    }
    // which engine runs a call: the tensor-core kernel for FAST arithmetic without debug payload, unless the caller
    // asked for the scalar one (spcies_batch_opts.engine)
    static bool use_mma(int arith, const BatchIO &io) {
        if constexpr (!HAS_MMA) return false;
        return arith != SPCIES_CUDA_ARITH_EXACT && io.sol == nullptr && io.engine != SPCIES_CUDA_ENGINE_SCALAR;
    }
    // systems the banded engine does not take (nn_ + mm_ > 8, N > 12): the generic dense tensor-core engine (MPC_FISTA_dense.cuh)
    static bool use_dense(int arith, const BatchIO &io) {
        if constexpr (!HAS_DENSE) return false;
        return arith != SPCIES_CUDA_ARITH_EXACT && io.sol == nullptr && io.LB == nullptr &&
               (io.engine == SPCIES_CUDA_ENGINE_MMA || (io.engine == SPCIES_CUDA_ENGINE_AUTO && DENSE_PREFERRED));
    }
    static bool park_engine(int arith, const BatchIO &io) { return !use_dense(arith, io); }   // park & resume: the FISTA kernels of this file
    static bool caps_engine(int arith, const BatchIO &io) { return use_mma(arith, io); }   // iteration-cap rounds
    static bool cl_engine(int arith, const BatchIO &io) { return use_mma(arith, io) && io.LB == nullptr; }   // closed loop inside the kernel
    static void engine_shape(int arith, const BatchIO &io, int &block, size_t &smem, int &ipb) {
        ipb = block;
        if (use_mma(arith, io)) {
            block = (io.phase != 2 && io.B >= MMA_BULK_MIN) ? MMA_BLOCK_BULK : MMA_BLOCK;
            smem = MMA_BYTES;
            ipb = block / 4;
        } else if (use_dense(arith, io)) {
            block = DP::BLOCK;
            smem = DP::SMEM;
            ipb = DP::IPB;
        }
    }
    static cudaError_t init_device_symbols() {
#if SPCIES_FISTA_CBANK
        return cudaMemcpyToSymbol(c_consts, &spcies_h_consts, sizeof spcies_h_consts);
#else
        return cudaSuccess;
#endif
    }
    static int default_block(bool varb) { return varb ? BLOCK1_VARB : BLOCK1_FIXED; }
    static int resume_block(bool varb) { return varb ? BLOCK2_VARB : BLOCK2_FIXED; }
    // which storage a block size runs with: the throughput block uses TMEM when available, the tail block shared memory
    static constexpr bool tm_for(int block, bool varb) {
        return block == (varb ? BLOCK1_VARB : BLOCK1_FIXED) ? USE_TMEM : TAIL_TMEM;
    }
    static size_t smem_bytes(int block, bool varb) {
        return BLOB_BYTES + (size_t)state_elems(varb, tm_for(block, varb)) * block * sizeof(real);
    }
    template <bool EXACT, bool VARB, int BLOCK, bool TM>
    static cudaError_t launch_t(int grid, size_t smem, cudaStream_t s, const BatchIO &io, const void *dc) {
        auto kern = fista_kernel<EXACT, VARB, BLOCK, TM>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, BLOCK, smem, s>>>(io, (const spcies_consts *)dc);
        return cudaGetLastError();
    }
    template <bool EXACT, bool VARB>
    static cudaError_t launch_b(int grid, int block, size_t smem, cudaStream_t s, const BatchIO &io, const void *dc) {
        constexpr int B1 = VARB ? BLOCK1_VARB : BLOCK1_FIXED;
        constexpr int B2 = VARB ? BLOCK2_VARB : BLOCK2_FIXED;
        if (block == B1) return launch_t<EXACT, VARB, B1, USE_TMEM>(grid, smem, s, io, dc);
        if constexpr (B2 != B1) {
            if (block == B2) return launch_t<EXACT, VARB, B2, TAIL_TMEM>(grid, smem, s, io, dc);
        }
        return cudaErrorInvalidConfiguration;
    }
    static size_t scratch_bytes(int, int, bool) { return 0; }
    static bool uses_scratch(int, const BatchIO &) { return false; }
    static size_t engine_scratch_bytes(int, const BatchIO &, int) { return 0; }
    static cudaError_t launch(int arith, bool varb, int grid, int block, size_t smem, cudaStream_t s, const BatchIO &io,
                              const void *dc, void *) {
        const bool ex = arith == SPCIES_CUDA_ARITH_EXACT;
        if (io.engine == SPCIES_CUDA_ENGINE_MMA && !use_mma(arith, io) && !use_dense(arith, io)) return cudaErrorNotSupported;
        if (io.engine == SPCIES_CUDA_ENGINE_SINGLE) return cudaErrorNotSupported;      // only through launch_single (small host-buffer calls)
        if constexpr (HAS_DENSE) {
            if (use_dense(arith, io)) {
                if (io.cl_steps > 0) return cudaErrorNotSupported;
                cudaError_t e = cudaFuncSetAttribute(dense::dense_mma_kernel<DenseEngine>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DP::SMEM);
                if (e != cudaSuccess) return e;
                dense::dense_mma_kernel<DenseEngine><<<grid, DP::BLOCK, DP::SMEM, s>>>(io, (const unsigned char *)dc);
                return cudaGetLastError();
            }
        }
        if constexpr (HAS_MMA) {
            if (use_mma(arith, io)) {
                const bool bulk = block == MMA_BLOCK_BULK && MMA_BLOCK_BULK != MMA_BLOCK;
                auto kern = bulk ? (varb ? fista_mma_kernel<true, MMA_BLOCK_BULK, false> : fista_mma_kernel<false, MMA_BLOCK_BULK, false>)
                                 : (varb ? fista_mma_kernel<true, MMA_BLOCK, false> : fista_mma_kernel<false, MMA_BLOCK, false>);
                if (io.cl_steps > 0) {      // closed loop: the instances stay on chip across the sampling times (no per-instance bounds)
                    if (varb) return cudaErrorNotSupported;
                    kern = bulk ? fista_mma_kernel<false, MMA_BLOCK_BULK, true> : fista_mma_kernel<false, MMA_BLOCK, true>;
                }
                if (block != MMA_BLOCK_BULK && block != MMA_BLOCK) return cudaErrorInvalidConfiguration;
                const size_t sm = MMA_BYTES;
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                if (e != cudaSuccess) return e;
                kern<<<grid, block, sm, s>>>(io, (const unsigned char *)dc);
                return cudaGetLastError();
            }
        }
        if (varb) return ex ? launch_b<true, true>(grid, block, smem, s, io, dc) : launch_b<false, true>(grid, block, smem, s, io, dc);
        return ex ? launch_b<true, false>(grid, block, smem, s, io, dc) : launch_b<false, false>(grid, block, smem, s, io, dc);
    }
    static cudaError_t attributes(int arith, bool varb, cudaFuncAttributes *a) {
        const bool ex = arith == SPCIES_CUDA_ARITH_EXACT;
        if constexpr (HAS_MMA) {
            if (!ex) return varb ? cudaFuncGetAttributes(a, fista_mma_kernel<true, MMA_BLOCK, false>) : cudaFuncGetAttributes(a, fista_mma_kernel<false, MMA_BLOCK, false>);
        }
        if constexpr (HAS_DENSE) {
            if (!ex && !varb) return cudaFuncGetAttributes(a, dense::dense_mma_kernel<DenseEngine>);
        }
        if (varb)
            return ex ? cudaFuncGetAttributes(a, fista_kernel<true, true, BLOCK1_VARB, USE_TMEM>)
                      : cudaFuncGetAttributes(a, fista_kernel<false, true, BLOCK1_VARB, USE_TMEM>);
        return ex ? cudaFuncGetAttributes(a, fista_kernel<true, false, BLOCK1_FIXED, USE_TMEM>)
                  : cudaFuncGetAttributes(a, fista_kernel<false, false, BLOCK1_FIXED, USE_TMEM>);
    }
};

}  // namespace fista
}  // namespace spcies

#define SPCIES_TRAITS ::spcies::fista::Traits
#include "spcies_entry.cuh"
