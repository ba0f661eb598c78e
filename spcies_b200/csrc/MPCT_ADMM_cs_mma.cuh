// MPCT_ADMM_cs_mma.cuh -- tensor-core engine policy of the MPCT ADMM_cs solver for spcies_dense_mma.cuh (included by
// MPCT_ADMM_cs.cuh, inside spcies::mpct_cs).
//
// The z update of code_MPCT_ADMM_cs_C.c:101-165,
//     q_hat = q + lambda - rho v;  rhs = AHi q_hat - b;  W mu = rhs (CSC L D L');  z = Hi q_hat + HiA mu,
// is the linear map  z = M (lambda - rho v) + M q(xr, ur) - Mb x0  with  M = Hi + HiA W^-1 AHi  and  Mb = HiA W^-1 (first n
// columns): with w = lambda - rho v (DIM / 8 tiles) and c = (x0, xr, ur) (the raw inputs: Tz, Sz and the per-stage replication of
// q are folded into the columns of F) it is  z = F [w ; c].  F is formed on the host in extended precision by pushing unit
// vectors through the generated CSR / CSC-LDL constants (the same L, Dinv the reference solves with).
// Iterates per group: v, lambda.  Vector rho is supported (rho, 1 / rho by tile and column in the staged table).
#pragma once
// (spcies_dense_mma.cuh is included by the parent header, outside its namespace)

struct Engine {
    static constexpr int ZT = (DIM + 7) / 8;
    static constexpr int NO = ZT, NW = ZT, NC = (2 * n + m + 7) / 8;
    static constexpr int NSTATE = 2 * ZT;            // v at [0, ZT), lambda at [ZT, 2 ZT)
    static constexpr int NB = 4, TEAM = 1;
    static constexpr bool OK = nrow_HiA == DIM;
    struct alignas(16) Small {
        double LB[ZT][8], UB[ZT][8], rho[ZT][8], rho_i[ZT][8];
    };
    struct Lane {};
    __device__ static __forceinline__ void lane_reset(Lane &) {}

    static inline double rho_h(const spcies_consts &C, int j, bool inv) {
#ifdef SCALAR_RHO
        (void)j;
        return inv ? (double)C.rho_i : (double)C.rho;
#else
        return inv ? (double)C.rho_i[j] : (double)C.rho[j];
#endif
    }
    static inline void fill(const spcies_consts &C, Small &S, long double *F) {
        typedef long double ld;
        for (int t = 0; t < ZT; ++t)
            for (int c = 0; c < 8; ++c) {
                const int j = t * 8 + c;
                S.LB[t][c] = j < DIM ? (double)C.LB[j] : 0.0;
                S.UB[t][c] = j < DIM ? (double)C.UB[j] : 0.0;
                S.rho[t][c] = j < DIM ? rho_h(C, j, false) : 0.0;
                S.rho_i[t][c] = j < DIM ? rho_h(C, j, true) : 0.0;
            }
        // z = Hi q_hat + HiA W^-1 (AHi q_hat - b)                                        code_MPCT_ADMM_cs_C.c:111-165
        auto apply = [&](const ld *qh, const ld *b, ld *z) {
            ld mu[NR];
            for (int i = 0; i < NR; ++i) {
                ld a = 0;
                for (int j = C.AHi_row[i]; j < C.AHi_row[i + 1]; ++j) a += (ld)C.AHi_val[j] * qh[C.AHi_col[j]];
                mu[i] = a;
            }
            for (int j = 0; j < n; ++j) mu[j] -= b[j];
            for (int i = 0; i < NR; ++i)
                for (int j = C.L_col[i]; j < C.L_col[i + 1]; ++j) mu[C.L_row[j]] -= (ld)C.L_val[j] * mu[i];
            for (int i = 0; i < NR; ++i) mu[i] *= (ld)C.Dinv[i];
            for (int i = NR - 1; i >= 0; --i)
                for (int j = C.L_col[i]; j < C.L_col[i + 1]; ++j) mu[i] -= (ld)C.L_val[j] * mu[C.L_row[j]];
            for (int i = 0; i < DIM; ++i) {
                ld a = 0;
                for (int j = C.Hi_row[i]; j < C.Hi_row[i + 1]; ++j) a += (ld)C.Hi_val[j] * qh[C.Hi_col[j]];
                for (int j = C.HiA_row[i]; j < C.HiA_row[i + 1]; ++j) a += (ld)C.HiA_val[j] * mu[C.HiA_col[j]];
                z[i] = a;
            }
        };
        constexpr int NINC = (NW + NC) * 8;
        ld *qh = new ld[DIM], *z = new ld[DIM], b[n];
        for (int col = 0; col < NINC; ++col) {
            for (int i = 0; i < DIM; ++i) qh[i] = 0;
            for (int i = 0; i < n; ++i) b[i] = 0;
            bool used = false;
            if (col < NW * 8) {
                if (col < DIM) {
                    qh[col] = 1;
                    used = true;
                }
            } else {
                const int e = col - NW * 8;                   // c = (x0 [n], xr [n], ur [m])
                if (e < n) {
                    b[e] = 1;
                    used = true;
                } else if (e < 2 * n) {                       // q[n + j] += Tz[j][i] xr[i], every stage      :73-77
                    for (int l = 0; l < N; ++l)
                        for (int j = 0; j < n; ++j) qh[l * DNM + n + j] = (ld)C.Tz[j][e - n];
                    used = true;
                } else if (e < 2 * n + m) {                   // q[2n + m + j] += Sz[j][i] ur[i]                :78-82
                    for (int l = 0; l < N; ++l)
                        for (int j = 0; j < m; ++j) qh[l * DNM + 2 * n + m + j] = (ld)C.Sz[j][e - 2 * n];
                    used = true;
                }
            }
            if (!used) continue;
            apply(qh, b, z);
            for (int i = 0; i < DIM; ++i) F[(size_t)i * NINC + col] = z[i];
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
        const double2 v = st[t * 32], lam = st[(ZT + t) * 32], rho = reinterpret_cast<const double2 *>(S->rho[t])[t4];
        return make_double2(fma(-rho.x, v.x, lam.x), fma(-rho.y, v.y, lam.y));                       // :101-107 without q
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
            const double2 rho = reinterpret_cast<const double2 *>(S->rho[t])[t4], rhi = reinterpret_cast<const double2 *>(S->rho_i[t])[t4];
            const double z0 = acc[b][0], z1 = acc[b][1];
            const double v0 = clip(fma(rhi.x, lam.x, z0), lo.x, hi.x), v1 = clip(fma(rhi.y, lam.y, z1), lo.y, hi.y);   // :169-178
            st[t * 32] = make_double2(v0, v1);
            st[(ZT + t) * 32] = make_double2(fma(rho.x, z0 - v0, lam.x), fma(rho.y, z1 - v1, lam.y));               // :182-188
            over = over || (fabs(vo.x - v0) > tl) || (fabs(z0 - v0) > tl) || (fabs(vo.y - v1) > tl) || (fabs(z1 - v1) > tl);
        }
    }
    __device__ static __forceinline__ void finish(Lane &, const spcies_consts *C, const BatchIO &io, long long inst, const double2 *st,
                                                  int t4) {
#pragma unroll
        for (int j = 0; j < m; ++j) {                                                   // u_opt = v[2 n + j]   :228-239
            const int e = 2 * n + j;
            if ((e % 8) / 2 == t4) {
                const double2 v = st[(e / 8) * 32];
                io.u[inst * m + j] = eng_u_out(C, (e & 1) ? v.y : v.x, j);
            }
        }
    }
};
