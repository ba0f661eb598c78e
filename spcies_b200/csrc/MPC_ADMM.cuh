// MPC_ADMM.cuh -- batched ADMM solver for the equMPC (SPCIES_TERMINAL == 0), laxMPC (== 1) and ellipMPC
// (== 2, terminal ellipsoid in the P^(1/2) metric) formulations, hand-written for sm_100a.
//
// Per instance it performs exactly the arithmetic of the reference templates
//   formulations/+equMPC/code_equMPC_ADMM_C.c:291-553
//   formulations/+laxMPC/code_laxMPC_ADMM_C.c:308-633      (adds the terminal block z_N, v_N, lambda_N, dense Hi_N)
//   formulations/+ellipMPC/code_ellipMPC_ADMM_C.c:108-447  (terminal radial projection onto (v-c)'P(v-c) <= r^2)
//
// One ADMM iteration is two fused sweeps over the horizon (no z, v1 or q_hat arrays are materialised;
// the reference keeps z[N-1][nm], v1[N-1][nm] and reuses z for q_hat):
//   pass A (l = 0..N-1): q_hat_l = q + lambda_l - rho v_l  ->  r.h.s. of the W system  ->  forward substitution mu_l
//   pass B (l = N-1..0): backward substitution mu_l  ->  z_l = -Hi (q_hat_l + G' mu)  ->  v_l = clip(z_l + lambda_l/rho)
//                        ->  lambda_l += rho (z_l - v_l)  ->  residuals |v_old - v|, |z - v|
// Every accumulation keeps the reference's operand order, so Arith<EXACT> reproduces gcc -O3 bit for bit.
//
// Persistent state per instance: v, lambda (decision-vector sized), mu[N][n], b, q, qT|xr  (+ LB, UB when the
// bounds are per instance).  The emitted .cu provides `spcies_consts` with the members named in
// cons_equMPC_ADMM_C.m / cons_laxMPC_ADMM_C.m / cons_ellipMPC_ADMM_C.m.
#pragma once
#include "spcies_kernel.cuh"
#include "spcies_mma.cuh"
#include "spcies_dense_mma.cuh"

namespace spcies {
namespace admm {

constexpr int n = nn_, m = mm_, nm = nm_, N = NN_;
constexpr int TERM = SPCIES_TERMINAL;   // 0 equ, 1 lax, 2 ellip
constexpr bool HAS_TN = (TERM != 0);    // terminal block z_N / v_N / lambda_N present

struct Solver {
    typedef SPCIES_REAL real;
    static constexpr int ZLEN = m + (N - 1) * nm + (HAS_TN ? n : 0);
    static constexpr int OFF_V = 0;                 // v_0[m], v[N-1][nm], (v_N[n])
    static constexpr int OFF_LAM = OFF_V + ZLEN;    // same layout
    static constexpr int OFF_MU = OFF_LAM + ZLEN;   // mu[N][n]
    static constexpr int OFF_B = OFF_MU + N * n;    // b[n]
    static constexpr int OFF_Q = OFF_B + n;         // q[nm]
    static constexpr int OFF_QT = OFF_Q + nm;       // qT[n] (lax, ellip) | xr[n] (equ)
    static constexpr int STATE = OFF_QT + n;
    static constexpr int OFF_LB = STATE;
    static constexpr int OFF_UB = OFF_LB + nm;
    static constexpr int STATE_VARB = OFF_UB + nm;
    static constexpr bool HAS_VARB = (TERM != 2);

    template <class A, bool VARB, class ST> struct Ctx {
        const spcies_consts *C;
        ST s;
        const BatchIO &io;
        double *sol;   // debug payload of the instance being solved (nullptr when not requested)
        __device__ Ctx(const spcies_consts *C_, ST s_, const BatchIO &io_) : C(C_), s(s_), io(io_), sol(nullptr) {}

        // ---- penalty parameter (scalar #define or per-element arrays, cons_equMPC_ADMM_C.m:113-120) ----
        __device__ __forceinline__ real rho0(int j) const {
#ifdef SCALAR_RHO
            return (real)rho;
#else
            return C->rho_0[j];
#endif
        }
        __device__ __forceinline__ real rhoi0(int j) const {
#ifdef SCALAR_RHO
            return (real)rho_i;
#else
            return C->rho_i_0[j];
#endif
        }
        __device__ __forceinline__ real rhoL(int l, int j) const {
#ifdef SCALAR_RHO
            return (real)rho;
#else
            return C->rho[l][j];
#endif
        }
        __device__ __forceinline__ real rhoiL(int l, int j) const {
#ifdef SCALAR_RHO
            return (real)rho_i;
#else
            return C->rho_i[l][j];
#endif
        }
#if SPCIES_TERMINAL != 0
        __device__ __forceinline__ real rhoN(int j) const {
#ifdef SCALAR_RHO
            return (real)rho;
#else
            return C->rho_N[j];
#endif
        }
        __device__ __forceinline__ real rhoiN(int j) const {
#ifdef SCALAR_RHO
            return (real)rho_i;
#else
            return C->rho_i_N[j];
#endif
        }
#endif
        // ---- bounds ----
        __device__ __forceinline__ real lb0(int j) const {
#if SPCIES_TERMINAL == 2
            return C->LBu0[j];
#elif defined(VAR_BOUNDS)
            return VARB ? s.ld(OFF_LB + n + j) : C->LB0[j];
#else
            return VARB ? s.ld(OFF_LB + n + j) : C->LB[n + j];
#endif
        }
        __device__ __forceinline__ real ub0(int j) const {
#if SPCIES_TERMINAL == 2
            return C->UBu0[j];
#elif defined(VAR_BOUNDS)
            return VARB ? s.ld(OFF_UB + n + j) : C->UB0[j];
#else
            return VARB ? s.ld(OFF_UB + n + j) : C->UB[n + j];
#endif
        }
        __device__ __forceinline__ real lbL(int l, int j) const {
#if SPCIES_TERMINAL == 2
            return C->LBz[l][j];
#elif defined(VAR_BOUNDS)
            return VARB ? s.ld(OFF_LB + j) : C->LB[l][j];
#else
            return VARB ? s.ld(OFF_LB + j) : C->LB[j];
#endif
        }
        __device__ __forceinline__ real ubL(int l, int j) const {
#if SPCIES_TERMINAL == 2
            return C->UBz[l][j];
#elif defined(VAR_BOUNDS)
            return VARB ? s.ld(OFF_UB + j) : C->UB[l][j];
#else
            return VARB ? s.ld(OFF_UB + j) : C->UB[j];
#endif
        }
#if SPCIES_TERMINAL == 1
        __device__ __forceinline__ real lbN(int j) const {
#ifdef VAR_BOUNDS
            return VARB ? s.ld(OFF_LB + j) : C->LBN[j];
#else
            return VARB ? s.ld(OFF_LB + j) : C->LB[j];
#endif
        }
        __device__ __forceinline__ real ubN(int j) const {
#ifdef VAR_BOUNDS
            return VARB ? s.ld(OFF_UB + j) : C->UBN[j];
#else
            return VARB ? s.ld(OFF_UB + j) : C->UB[j];
#endif
        }
#endif

        // ---- set-up: b = -A x0, q = [Q xr; R ur], qT = T xr, iterates = 0      code_equMPC_ADMM_C.c:268-283 ----
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
                real b = real(0);
#pragma unroll
                for (int i = 0; i < n; ++i) b = A::nmsub(b, C->AB[j][i], x0[i]);
                s.st(OFF_B + j, b);
                s.st(OFF_Q + j, A::mul(C->Q[j], xr[j]));
#if SPCIES_TERMINAL == 0
                s.st(OFF_QT + j, xr[j]);
#else
                real qT = real(0);                                  // code_laxMPC_ADMM_C.c:292-295 (T dense, negated)
#pragma unroll
                for (int i = 0; i < n; ++i) qT = A::madd(qT, C->T[j][i], xr[i]);
                s.st(OFF_QT + j, qT);
#endif
            }
#pragma unroll
            for (int j = 0; j < m; ++j) s.st(OFF_Q + n + j, A::mul(C->R[j], ur[j]));
            if (VARB) {
#pragma unroll
                for (int j = 0; j < nm; ++j) {
                    s.st(OFF_LB + j, (real)io.LB[inst * nm + j]);
                    s.st(OFF_UB + j, (real)io.UB[inst * nm + j]);
                }
            }
#pragma unroll 4
            for (int e = 0; e < 2 * ZLEN; ++e) s.st(OFF_V + e, real(0));
            sol = io.sol ? io.sol + inst * (long long)(sizeof(SPCIES_SOL_T) / sizeof(double)) : nullptr;
        }

        // forward / backward substitution of one block row, code_equMPC_ADMM_C.c:356-422 (same as solve_W_matrix_form)
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
        // q_hat of stage block l: q + lambda - rho*v                               code_equMPC_ADMM_C.c:317-326
        __device__ __forceinline__ void qhat_block(real (&z)[nm], real (&v)[nm], real (&lam)[nm], const real (&q)[nm], int l) const {
#pragma unroll
            for (int j = 0; j < nm; ++j) {
                v[j] = s.ld(OFF_V + m + l * nm + j);
                lam[j] = s.ld(OFF_LAM + m + l * nm + j);
                z[j] = A::nmsub(A::add(q[j], lam[j]), rhoL(l, j), v[j]);
            }
        }
        // |v_old - v| > tol || |z - v| > tol                                        code_equMPC_ADMM_C.c:497-502
        __device__ __forceinline__ bool res_exceeds(real v_old, real v_new, real z) const {
            return exceeds(A::sub(v_old, v_new), (real)tol) || exceeds(A::sub(z, v_new), (real)tol);
        }

        __device__ bool iterate(int /*k*/) {
            real q[nm], zp[nm], zc[nm], vv[nm], ll[nm], mu[n], mprev[n], mnext[n], z0h[m];
#if SPCIES_TERMINAL != 0
            real zN[n], vN[n], lN[n];
#endif
            bool over = false;
#pragma unroll
            for (int j = 0; j < nm; ++j) q[j] = s.ld(OFF_Q + j);

            // ================= pass A: q_hat -> r.h.s. -> forward substitution =================
            // stage 0                                                               :304-313, :332-338, :356-362
#pragma unroll
            for (int j = 0; j < m; ++j)
                z0h[j] = A::nmsub(A::add(q[n + j], s.ld(OFF_LAM + j)), rho0(j), s.ld(OFF_V + j));
            qhat_block(zc, vv, ll, q, 0);
#pragma unroll
            for (int j = 0; j < n; ++j) {
                real r = A::madd(-s.ld(OFF_B + j), C->Hi[0][j], zc[j]);                  // Hi[0][j]*z[0][j] - b[j]
#pragma unroll
                for (int i = 0; i < m; ++i) r = A::nmsub(r, cprod<A>(C->AB[j][n + i], C->Hi_0[i]), z0h[i]);
                mu[j] = r;
            }
            fwd_block(mu, mprev, 0, true);
#pragma unroll
            for (int j = 0; j < n; ++j) {
                s.st(OFF_MU + j, mu[j]);
                mprev[j] = mu[j];
            }
#pragma unroll
            for (int j = 0; j < nm; ++j) zp[j] = zc[j];
            // stages 1 .. N-2                                                        :341-348, :365-376
#pragma unroll 1
            for (int l = 1; l < N - 1; ++l) {
                qhat_block(zc, vv, ll, q, l);
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    real r = A::mul(C->Hi[l][j], zc[j]);
#pragma unroll
                    for (int i = 0; i < nm; ++i) r = A::nmsub(r, cprod<A>(C->AB[j][i], C->Hi[l - 1][i]), zp[i]);
                    mu[j] = r;
                }
                fwd_block(mu, mprev, l, false);
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    s.st(OFF_MU + l * n + j, mu[j]);
                    mprev[j] = mu[j];
                }
#pragma unroll
                for (int j = 0; j < nm; ++j) zp[j] = zc[j];
            }
            // stage N-1: terminal block                                              :351-353 | lax :373-381
#if SPCIES_TERMINAL != 0
#pragma unroll
            for (int j = 0; j < n; ++j) {
                vN[j] = s.ld(OFF_V + m + (N - 1) * nm + j);
                lN[j] = s.ld(OFF_LAM + m + (N - 1) * nm + j);
            }
#pragma unroll
            for (int j = 0; j < n; ++j) {
#if SPCIES_TERMINAL == 1
                zN[j] = A::nmsub(A::add(s.ld(OFF_QT + j), lN[j]), rhoN(j), vN[j]);
#else
                real a = s.ld(OFF_QT + j);                                              // code_ellipMPC_ADMM_C.c:146-155
#pragma unroll
                for (int i = 0; i < n; ++i)
                    a = A::nmsub(A::madd(a, C->P_half[j][i], lN[i]), A::mul(C->P[j][i], rhoN(i)), vN[i]);
                zN[j] = a;
#endif
            }
#endif
#pragma unroll
            for (int j = 0; j < n; ++j) {
                real r = real(0);
#if SPCIES_TERMINAL != 0
#pragma unroll
                for (int i = 0; i < n; ++i) r = A::madd(r, C->Hi_N[j][i], zN[i]);
#endif
#pragma unroll
                for (int i = 0; i < nm; ++i) r = A::nmsub(r, cprod<A>(C->AB[j][i], C->Hi[N - 2][i]), zp[i]);
#if SPCIES_TERMINAL == 0
                r = A::sub(r, s.ld(OFF_QT + j));                                         // - xr[j]
#endif
                mu[j] = r;
            }
            fwd_block(mu, mprev, N - 1, false);

            // ================= pass B: backward substitution -> z -> v -> lambda -> residuals =================
            bwd_block(mu, mnext, N - 1, true);
#if SPCIES_TERMINAL != 0
            {   // terminal block z_N, v_N, lambda_N                        code_laxMPC_ADMM_C.c:476-485, :523-538, :561-569
                real aux[n], zNn[n], vNn[n];
#pragma unroll
                for (int j = 0; j < n; ++j) aux[j] = A::sub(zN[j], mu[j]);
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    real a = real(0);
#pragma unroll
                    for (int i = 0; i < n; ++i) a = A::nmsub(a, C->Hi_N[j][i], aux[i]);
                    zNn[j] = a;
                }
#if SPCIES_TERMINAL == 1
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    vNn[j] = clip(A::madd(zNn[j], rhoiN(j), lN[j]), lbN(j), ubN(j));
                    s.st(OFF_LAM + m + (N - 1) * nm + j, A::madd(lN[j], rhoN(j), A::sub(zNn[j], vNn[j])));
                }
#else
                // radial projection onto the ellipsoid in the P metric          code_ellipMPC_ADMM_C.c:319-351, :375-387
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    real a = zNn[j];
#pragma unroll
                    for (int i = 0; i < n; ++i) a = A::madd(a, A::mul(C->Pinv_half[j][i], rhoiN(i)), lN[i]);
                    vNn[j] = a;
                }
                real vPv = real(0);
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    real a = real(0);
#pragma unroll
                    for (int i = 0; i < n; ++i) a = A::madd(a, C->P[j][i], A::sub(vNn[i], C->c[i]));
                    aux[j] = a;
                }
#pragma unroll
                for (int j = 0; j < n; ++j) vPv = A::madd(vPv, A::sub(vNn[j], C->c[j]), aux[j]);
                if (vPv > A::mul(C->r, C->r)) {
                    vPv = A::div(C->r, A::sqrt(vPv));
#pragma unroll
                    for (int j = 0; j < n; ++j) vNn[j] = A::madd(C->c[j], vPv, A::sub(vNn[j], C->c[j]));
                }
#pragma unroll
                for (int j = 0; j < n; ++j) aux[j] = A::mul(rhoN(j), A::sub(zNn[j], vNn[j]));
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    real a = lN[j];
#pragma unroll
                    for (int i = 0; i < n; ++i) a = A::madd(a, C->P_half[j][i], aux[i]);
                    s.st(OFF_LAM + m + (N - 1) * nm + j, a);
                }
#endif
#pragma unroll
                for (int j = 0; j < n; ++j) {
                    over |= res_exceeds(vN[j], vNn[j], zNn[j]);
                    s.st(OFF_V + m + (N - 1) * nm + j, vNn[j]);
                    if (sol) sol[m + (N - 1) * nm + j] = (double)zNn[j];
                }
            }
#endif
#pragma unroll
            for (int j = 0; j < n; ++j) mnext[j] = mu[j];

#pragma unroll 1
            for (int l = N - 2; l >= 0; --l) {
#pragma unroll
                for (int j = 0; j < n; ++j) mu[j] = s.ld(OFF_MU + l * n + j);
                bwd_block(mu, mnext, l, false);
                // z[l] = -Hi[l] o (q_hat[l] - [mu_l; 0] + [A B]' mu_{l+1})            :431-440
                qhat_block(zc, vv, ll, q, l);
#pragma unroll
                for (int j = 0; j < n; ++j) zc[j] = A::sub(zc[j], mu[j]);
#pragma unroll
                for (int j = 0; j < nm; ++j) {
#pragma unroll
                    for (int i = 0; i < n; ++i) zc[j] = A::madd(zc[j], C->AB[i][j], mnext[i]);
                    zc[j] = A::mul(-C->Hi[l][j], zc[j]);
                }
                // v, lambda, residuals                                               :458-472, :486-494, :509-524
#pragma unroll
                for (int j = 0; j < nm; ++j) {
                    const real vn = clip(A::madd(zc[j], rhoiL(l, j), ll[j]), lbL(l, j), ubL(l, j));
                    over |= res_exceeds(vv[j], vn, zc[j]);
                    s.st(OFF_V + m + l * nm + j, vn);
                    s.st(OFF_LAM + m + l * nm + j, A::madd(ll[j], rhoL(l, j), A::sub(zc[j], vn)));
                    if (sol) sol[m + l * nm + j] = (double)zc[j];
                }
#pragma unroll
                for (int j = 0; j < n; ++j) mnext[j] = mu[j];
            }
            // first m decision variables                                              :424-429, :447-456, :476-483
#pragma unroll
            for (int j = 0; j < m; ++j) {
                const real v0 = s.ld(OFF_V + j), l0 = s.ld(OFF_LAM + j);
                real z = A::nmsub(A::add(q[n + j], l0), rho0(j), v0);
#pragma unroll
                for (int i = 0; i < n; ++i) z = A::madd(z, C->AB[i][n + j], mnext[i]);
                z = A::mul(-C->Hi_0[j], z);
                const real vn = clip(A::madd(z, rhoi0(j), l0), lb0(j), ub0(j));
                over |= res_exceeds(v0, vn, z);
                s.st(OFF_V + j, vn);
                s.st(OFF_LAM + j, A::madd(l0, rho0(j), A::sub(z, vn)));
                if (sol) sol[j] = (double)z;
            }
            return !over;
        }

        // u_opt = v_0 (the clipped copy, :557-566); debug payload z (written during the last pass B), v, lambda
        __device__ void finish(long long inst, int k, int ef) {
#pragma unroll
            for (int j = 0; j < m; ++j) io.u[inst * m + j] = eng_u_out(C, (double)s.ld(OFF_V + j), j);
            io.k[inst] = k;
            io.e[inst] = ef;
            if (sol) {
                for (int e = 0; e < ZLEN; ++e) {
                    sol[ZLEN + e] = (double)s.ld(OFF_V + e);
                    sol[2 * ZLEN + e] = (double)s.ld(OFF_LAM + e);
                }
                for (int e = 3 * ZLEN; e < (int)(sizeof(SPCIES_SOL_T) / sizeof(double)); ++e) sol[e] = 0.0;
            }
        }
    };
};

#include "MPC_ADMM_mma.cuh"
#include "MPC_ADMM_dense.cuh"

// Host-side traits: the scalar skeleton plus the tensor-core engines (FAST arithmetic, no debug payload): the banded one
// (MPC_ADMM_mma.cuh) for the systems it takes, else the generic dense one (MPC_ADMM_dense.cuh: equMPC / laxMPC, any size,
// vector rho, VAR_BOUNDS)
struct Traits : PolicyTraits<Solver> {
    typedef PolicyTraits<Solver> Base;
    static bool &mma_ok() {
        static bool ok = false;   // set by fill_blob(): the generated constants are uniform over the horizon
        return ok;
    }
#if SPCIES_TERMINAL != 2
    typedef dense::Plan<DenseEngine> DP;
#endif
    static size_t blob_bytes() {
#if SPCIES_TERMINAL != 2
        if (HAS_DENSE) return DP::BLOB_BYTES;
#endif
#if SPCIES_ADMM_MMA_ELIGIBLE
        if (HAS_MMA) return MMA_OFFSET + MMA_BYTES;
#endif
        return Base::blob_bytes();
    }
    static void fill_blob(void *dst) {
        memset(dst, 0, blob_bytes());
        Base::fill_blob(dst);
#if SPCIES_ADMM_MMA_ELIGIBLE
        if constexpr (HAS_MMA) {
            MmaTables *T = new MmaTables;
            mma_ok() = fill_mma_tables(spcies_h_consts, *T);
            memcpy((char *)dst + MMA_OFFSET, T, sizeof *T);
            delete T;
        }
#endif
#if SPCIES_TERMINAL != 2
        if constexpr (HAS_DENSE) {
            DenseEngine::Small *S = new DenseEngine::Small;
            memset(S, 0, sizeof *S);
            long double *F = new long double[(size_t)DenseEngine::NO * 8 * DP::NIN * 8]();
            DenseEngine::fill(spcies_h_consts, *S, F);
            dense::fill_fragments<DenseEngine>(F, reinterpret_cast<double2 *>((char *)dst + DP::OFF_FRAG));
            memcpy((char *)dst + DP::OFF_SMALL, S, sizeof *S);
            delete[] F;
            delete S;
        }
#endif
    }
    static bool use_mma(int arith, const BatchIO &io) {
        if constexpr (!HAS_MMA) return false;
        return mma_ok() && arith != SPCIES_CUDA_ARITH_EXACT && io.sol == nullptr && io.engine != SPCIES_CUDA_ENGINE_SCALAR;
    }
    static bool use_dense(int arith, const BatchIO &io) {
        if constexpr (!HAS_DENSE) return false;
        return arith != SPCIES_CUDA_ARITH_EXACT && io.sol == nullptr && io.LB == nullptr &&
               (io.engine == SPCIES_CUDA_ENGINE_MMA || (io.engine == SPCIES_CUDA_ENGINE_AUTO && DENSE_PREFERRED));
    }
    static bool uses_scratch(int arith, const BatchIO &io) { return !use_mma(arith, io) && !use_dense(arith, io); }
    static void engine_shape(int arith, const BatchIO &io, int &block, size_t &smem, int &ipb) {
        ipb = block;
#if SPCIES_ADMM_MMA_ELIGIBLE
        if (use_mma(arith, io)) {
            block = MMA_BLOCK;
            smem = MMA_SMEM;
            ipb = MMA_IPB;
        }
#endif
#if SPCIES_TERMINAL != 2
        if (use_dense(arith, io)) {
            block = DP::BLOCK;
            smem = DP::SMEM;
            ipb = DP::IPB;
        }
#endif
    }
    static cudaError_t launch(int arith, bool varb, int grid, int block, size_t smem, cudaStream_t s, const BatchIO &io,
                              const void *dc, void *scratch) {
        if (io.engine == SPCIES_CUDA_ENGINE_MMA && !use_mma(arith, io) && !use_dense(arith, io)) return cudaErrorNotSupported;
#if SPCIES_ADMM_MMA_ELIGIBLE
        if constexpr (HAS_MMA) {
            if (use_mma(arith, io)) {
                auto kern = varb ? admm_mma_kernel<true> : admm_mma_kernel<false>;
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MMA_SMEM);
                if (e != cudaSuccess) return e;
                kern<<<grid, MMA_BLOCK, MMA_SMEM, s>>>(io, (const unsigned char *)dc);
                return cudaGetLastError();
            }
        }
#endif
#if SPCIES_TERMINAL != 2
        if constexpr (HAS_DENSE) {
            if (use_dense(arith, io)) {
                cudaError_t e = cudaFuncSetAttribute(dense::dense_mma_kernel<DenseEngine>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DP::SMEM);
                if (e != cudaSuccess) return e;
                dense::dense_mma_kernel<DenseEngine><<<grid, DP::BLOCK, DP::SMEM, s>>>(io, (const unsigned char *)dc);
                return cudaGetLastError();
            }
        }
#endif
        BatchIO io2 = io;
        io2.engine = SPCIES_CUDA_ENGINE_AUTO;
        return Base::launch(arith, varb, grid, block, smem, s, io2, dc, scratch);
    }
    static cudaError_t attributes(int arith, bool varb, cudaFuncAttributes *a) {
#if SPCIES_ADMM_MMA_ELIGIBLE
        if constexpr (HAS_MMA) {
            if (mma_ok() && arith != SPCIES_CUDA_ARITH_EXACT)
                return varb ? cudaFuncGetAttributes(a, admm_mma_kernel<true>) : cudaFuncGetAttributes(a, admm_mma_kernel<false>);
        }
#endif
#if SPCIES_TERMINAL != 2
        if constexpr (HAS_DENSE) {
            if (arith != SPCIES_CUDA_ARITH_EXACT && !varb) return cudaFuncGetAttributes(a, dense::dense_mma_kernel<DenseEngine>);
        }
#endif
        return Base::attributes(arith, varb, a);
    }
};

}  // namespace admm
}  // namespace spcies

#define SPCIES_TRAITS ::spcies::admm::Traits
#include "spcies_entry.cuh"
