// MPC_ADMM_dense.cuh -- policy of the equMPC / laxMPC ADMM solvers for the generic dense tensor-core engine (spcies_dense_mma.cuh);
// included by MPC_ADMM.cuh, inside spcies::admm, after MPC_ADMM_mma.cuh.
//
// The banded engine (MPC_ADMM_mma.cuh) takes 5 <= nn_ <= 6, nn_ + mm_ <= 8, a scalar penalty and stage-independent bounds.  Every
// other equMPC / laxMPC ADMM solver (other system sizes, `rho` arrays, VAR_BOUNDS) used to fall back to the one-thread-per-instance
// kernel; this policy runs it on the tensor cores.  The z update of code_equMPC_ADMM_C.c:297-445 (code_laxMPC_ADMM_C.c:318-487),
//     q_hat = q + lambda - rho v;   rhs = -G H^-1 q_hat - b;   W mu = rhs (Alpha / Beta);   z = -H^-1 (q_hat + G' mu),
// is linear in w = lambda - rho v and in c = (x0, xr, ur): z = F [w ; c].  F is formed on the host in extended precision by pushing
// unit vectors through the generated constants (Hi, Hi_0, Hi_N, AB, Alpha, Beta, Q, R, T) in the reference's own order.
// Iterates per group: v, lambda.  Per-element rho and bounds come from a staged table (vector rho, VAR_BOUNDS).
#pragma once
// (spcies_dense_mma.cuh is included by the parent header, outside its namespace)

#ifndef SPCIES_ADMM_DENSE
#define SPCIES_ADMM_DENSE 1
#endif

#if SPCIES_TERMINAL != 2
struct DenseEngine {
    static constexpr int ZLEN = Solver::ZLEN;
    static constexpr int ZT = (ZLEN + 7) / 8;
    static constexpr int NO = ZT, NW = ZT, NC = (2 * n + m + 7) / 8;
    static constexpr int NSTATE = 2 * ZT;            // v at [0, ZT), lambda at [ZT, 2 ZT)
    static constexpr int NB = 4, TEAM = 1;
    static constexpr bool OK = SPCIES_ADMM_DENSE != 0;
    struct alignas(16) Small {
        double LB[ZT][8], UB[ZT][8], pen[ZT][8], pen_i[ZT][8];      // (`rho` / `rho_i` are #defines of the scalar-penalty solvers)
    };
    struct Lane {};
    __device__ static __forceinline__ void lane_reset(Lane &) {}

    // element e of z / v / lambda: u_0[j] | stage l (x_{l+1}, u_{l+1})[j] | x_N[j]
    static inline void split(int e, int &kind, int &l, int &j) {
        if (e < m) {
            kind = 0; l = 0; j = e;
        } else if (e < m + (N - 1) * nm) {
            kind = 1; l = (e - m) / nm; j = (e - m) % nm;
        } else {
            kind = 2; l = 0; j = e - m - (N - 1) * nm;
        }
    }
    static inline void fill(const spcies_consts &C, Small &S, long double *F) {
        typedef long double ld;
        for (int t = 0; t < ZT; ++t)
            for (int c = 0; c < 8; ++c) {
                const int e = t * 8 + c;
                double lb = 0.0, ub = 0.0, rh = 0.0, ri = 0.0;
                if (e < ZLEN) {
                    int kind, l, j;
                    split(e, kind, l, j);
#ifdef SCALAR_RHO
                    rh = (double)rho;
                    ri = (double)rho_i;
#else
                    rh = kind == 0 ? (double)C.rho_0[j] : (kind == 1 ? (double)C.rho[l][j] : 0.0);
                    ri = kind == 0 ? (double)C.rho_i_0[j] : (kind == 1 ? (double)C.rho_i[l][j] : 0.0);
#if SPCIES_TERMINAL == 1
                    if (kind == 2) {
                        rh = (double)C.rho_N[j];
                        ri = (double)C.rho_i_N[j];
                    }
#endif
#endif
#ifdef VAR_BOUNDS
                    lb = kind == 0 ? (double)C.LB0[j] : (kind == 1 ? (double)C.LB[l][j] : 0.0);
                    ub = kind == 0 ? (double)C.UB0[j] : (kind == 1 ? (double)C.UB[l][j] : 0.0);
#if SPCIES_TERMINAL == 1
                    if (kind == 2) {
                        lb = (double)C.LBN[j];
                        ub = (double)C.UBN[j];
                    }
#endif
#else
                    lb = kind == 0 ? (double)C.LB[n + j] : (double)C.LB[j];
                    ub = kind == 0 ? (double)C.UB[n + j] : (double)C.UB[j];
#endif
                }
                S.LB[t][c] = lb;
                S.UB[t][c] = ub;
                S.pen[t][c] = rh;
                S.pen_i[t][c] = ri;
            }
        // the reference's z update on (q_hat, b, xr)                                    code_equMPC_ADMM_C.c:321-445
        auto apply = [&](const ld *qh, const ld *b, const ld *xr, ld *zout) {
            ld z0[m], z[N - 1][nm], mu[N][n];
#if SPCIES_TERMINAL == 1
            ld zN[n];
            for (int j = 0; j < n; ++j) zN[j] = qh[m + (N - 1) * nm + j];
#endif
            for (int j = 0; j < m; ++j) z0[j] = qh[j];
            for (int l = 0; l < N - 1; ++l)
                for (int j = 0; j < nm; ++j) z[l][j] = qh[m + l * nm + j];
            for (int j = 0; j < n; ++j) {
                mu[0][j] = (ld)C.Hi[0][j] * z[0][j] - b[j];
                for (int i = 0; i < m; ++i) mu[0][j] -= (ld)cprod_host(C.AB[j][i + n], C.Hi_0[i]) * z0[i];
            }
            for (int l = 1; l < N - 1; ++l)
                for (int j = 0; j < n; ++j) {
                    mu[l][j] = (ld)C.Hi[l][j] * z[l][j];
                    for (int i = 0; i < nm; ++i) mu[l][j] -= (ld)cprod_host(C.AB[j][i], C.Hi[l - 1][i]) * z[l - 1][i];
                }
            for (int j = 0; j < n; ++j) {
                mu[N - 1][j] = 0;
#if SPCIES_TERMINAL == 1
                for (int i = 0; i < n; ++i) mu[N - 1][j] += (ld)C.Hi_N[j][i] * zN[i];
#endif
                for (int i = 0; i < nm; ++i) mu[N - 1][j] -= (ld)cprod_host(C.AB[j][i], C.Hi[N - 2][i]) * z[N - 2][i];
#if SPCIES_TERMINAL == 0
                mu[N - 1][j] -= xr[j];
#endif
            }
            for (int l = 0; l < N; ++l)                                                   // forward substitution
                for (int j = 0; j < n; ++j) {
                    if (l > 0)
                        for (int i = 0; i < n; ++i) mu[l][j] -= (ld)C.Alpha[l - 1][i][j] * mu[l - 1][i];
                    for (int i = 0; i < j; ++i) mu[l][j] -= (ld)C.Beta[l][i][j] * mu[l][i];
                    mu[l][j] *= (ld)C.Beta[l][j][j];
                }
            for (int l = N - 1; l >= 0; --l)                                              // backward substitution
                for (int j = n - 1; j >= 0; --j) {
                    if (l < N - 1)
                        for (int i = 0; i < n; ++i) mu[l][j] -= (ld)C.Alpha[l][j][i] * mu[l + 1][i];
                    for (int i = n - 1; i > j; --i) mu[l][j] -= (ld)C.Beta[l][j][i] * mu[l][i];
                    mu[l][j] *= (ld)C.Beta[l][j][j];
                }
            for (int j = 0; j < m; ++j) {
                ld a = z0[j];
                for (int i = 0; i < n; ++i) a += (ld)C.AB[i][j + n] * mu[0][i];
                zout[j] = -(ld)C.Hi_0[j] * a;
            }
            for (int l = 0; l < N - 1; ++l)
                for (int j = 0; j < nm; ++j) {
                    ld a = z[l][j] - (j < n ? mu[l][j] : (ld)0);
                    for (int i = 0; i < n; ++i) a += (ld)C.AB[i][j] * mu[l + 1][i];
                    zout[m + l * nm + j] = -(ld)C.Hi[l][j] * a;
                }
#if SPCIES_TERMINAL == 1
            for (int j = 0; j < n; ++j) {
                ld a = 0;
                for (int i = 0; i < n; ++i) a -= (ld)C.Hi_N[j][i] * (zN[i] - mu[N - 1][i]);
                zout[m + (N - 1) * nm + j] = a;
            }
#endif
            (void)xr;
        };
        constexpr int NINC = (NW + NC) * 8;
        ld *qh = new ld[ZLEN], *z = new ld[ZLEN], b[n], xr[n];
        for (int col = 0; col < NINC; ++col) {
            for (int i = 0; i < ZLEN; ++i) qh[i] = 0;
            for (int i = 0; i < n; ++i) b[i] = xr[i] = 0;
            bool used = false;
            if (col < NW * 8) {
                if (col < ZLEN) {
                    qh[col] = 1;
                    used = true;
                }
            } else {
                const int e = col - NW * 8;                   // c = (x0 [n], xr [n], ur [m])
                if (e < n) {                                  // b = -A x0                                    :268-274
                    for (int j = 0; j < n; ++j) b[j] = -(ld)C.AB[j][e];
                    used = true;
                } else if (e < 2 * n) {                       // q[j] = Q[j] xr[j] in every stage, qT = T xr | the last r.h.s.   :276-279
                    const int i = e - n;
                    for (int l = 0; l < N - 1; ++l) qh[m + l * nm + i] = (ld)C.Q[i];
#if SPCIES_TERMINAL == 1
                    for (int j = 0; j < n; ++j) qh[m + (N - 1) * nm + j] = (ld)C.T[j][i];
#else
                    xr[i] = 1;
#endif
                    used = true;
                } else if (e < 2 * n + m) {                   // q[n + j] = R[j] ur[j]                        :280-282
                    const int i = e - 2 * n;
                    qh[i] = (ld)C.R[i];
                    for (int l = 0; l < N - 1; ++l) qh[m + l * nm + n + i] = (ld)C.R[i];
                    used = true;
                }
            }
            if (!used) continue;
            apply(qh, b, xr, z);
            for (int i = 0; i < ZLEN; ++i) F[(size_t)i * NINC + col] = z[i];
        }
        delete[] qh;
        delete[] z;
    }

    __device__ static __forceinline__ void init(Lane &, const spcies_consts *C, const Small *, const BatchIO &io, long long inst,
                                                double2 *st, double2 *cin, int t4, int /*rank*/) {
#pragma unroll 4
        for (int t = 0; t < NSTATE; ++t) st[t * 32] = make_double2(0.0, 0.0);
#pragma unroll
        for (int t = 0; t < NC; ++t) {
            double v[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int e = t * 8 + 2 * t4 + i;
                v[i] = e < n ? eng_x(C, io.x0, inst, n, e)
                             : (e < 2 * n ? eng_x(C, io.xr, inst, n, e - n) : (e < 2 * n + m ? eng_u(C, io.ur, inst, m, e - 2 * n) : 0.0));
            }
            cin[t * 32] = make_double2(v[0], v[1]);
        }
    }
    __device__ static __forceinline__ double2 make_w(Lane &, const spcies_consts *, const Small *S, int t, const double2 *st, int t4) {
        const double2 v = st[t * 32], lam = st[(ZT + t) * 32], rh = reinterpret_cast<const double2 *>(S->pen[t])[t4];
        return make_double2(fma(-rh.x, v.x, lam.x), fma(-rh.y, v.y, lam.y));                            // :300-318 without q
    }
    __device__ static __forceinline__ void update(Lane &, const spcies_consts *, const Small *S, int t0, const double (&acc)[NB][2],
                                                  double2 *st, int t4, bool &over) {
        const double tl = (double)tol;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int t = t0 + b;
            if (t >= ZT) break;
            const double2 vo = st[t * 32], lam = st[(ZT + t) * 32];
            const double2 lo = reinterpret_cast<const double2 *>(S->LB[t])[t4], hi = reinterpret_cast<const double2 *>(S->UB[t])[t4];
            const double2 rh = reinterpret_cast<const double2 *>(S->pen[t])[t4], ri = reinterpret_cast<const double2 *>(S->pen_i[t])[t4];
            const double z0 = acc[b][0], z1 = acc[b][1];
            const double v0 = clip(fma(ri.x, lam.x, z0), lo.x, hi.x), v1 = clip(fma(ri.y, lam.y, z1), lo.y, hi.y);   // :447-487
            st[t * 32] = make_double2(v0, v1);
            st[(ZT + t) * 32] = make_double2(fma(rh.x, z0 - v0, lam.x), fma(rh.y, z1 - v1, lam.y));               // :489-510
            over = over || (fabs(vo.x - v0) > tl) || (fabs(z0 - v0) > tl) || (fabs(vo.y - v1) > tl) || (fabs(z1 - v1) > tl);   // :497-524
        }
    }
    __device__ static __forceinline__ void finish(Lane &, const spcies_consts *C, const BatchIO &io, long long inst, const double2 *st,
                                                  int t4) {
#pragma unroll
        for (int j = 0; j < m; ++j)                                                                     // u_opt = v_0   :557-566
            if ((j % 8) / 2 == t4) {
                const double2 v = st[(j / 8) * 32];
                io.u[inst * m + j] = eng_u_out(C, (j & 1) ? v.y : v.x, j);
            }
    }
};
constexpr bool HAS_DENSE = !HAS_MMA && dense::Plan<DenseEngine>::HAS;
// AUTO picks it where it beats the one-thread-per-instance kernel: larger problems, and those whose iterates the scalar kernel
// can only keep in its global scratch.  Measured on 64 Ki instances (tools/fallback_times.py, profiles/r2_fallback_times.json):
// n = 8, m = 4, N = 8 (|z| = 88 / 96): 1.42x; n = 6, m = 2, N = 10 with a penalty vector (|z| = 74): 0.98x; n = 4, m = 2, N = 8 (|z| = 44): 0.66x
constexpr bool DENSE_PREFERRED = HAS_DENSE && (DenseEngine::ZLEN >= 80 || KernelPlan<Solver>::GSTATE_FIXED);
#else
constexpr bool HAS_DENSE = false;
constexpr bool DENSE_PREFERRED = false;
#endif
