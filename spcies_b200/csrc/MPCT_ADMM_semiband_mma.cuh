// MPCT_ADMM_semiband_mma.cuh -- tensor-core engine policy of the MPCT ADMM_semiband solver for spcies_dense_mma.cuh (included by
// MPCT_ADMM_semiband.cuh, inside spcies::mpct_sb).
//
// The z update of code_MPCT_ADMM_semiband_C.c:141-770 -- three Woodbury-structured solves (9a), (9b), (9c) -- is the solution of
// the equality-constrained QP
//     min 1/2 z'(H + rho I) z + p'z   s.t.  G z = b,       p = lambda - rho v + q(xr, ur),  b = (x0, 0, ..., 0),
// i.e. the linear map  [H + rho I, G'; G, 0] [z; mu] = [-p; b]  of (w = lambda - rho v, x0, xr, ur).  For a batch that shares the
// model the structure exploited per instance by the reference buys nothing: z = F [w ; c] is one GEMM.  F is formed on the host by
// Gaussian elimination of the KKT matrix in extended precision from the generated Q, R, S, T, A, B (the one-iteration answer of
// the instantiated reference template is reproduced to 3e-13, tests/test_oracle_golden.py).
// Iterates per group: v, lambda.
#pragma once
// (spcies_dense_mma.cuh is included by the parent header, outside its namespace)

struct Engine {
    static constexpr int ZT = (L + 7) / 8;
    static constexpr int NO = ZT, NW = ZT, NC = (2 * n + m + 7) / 8;
    static constexpr int NSTATE = 2 * ZT;            // v at [0, ZT), lambda at [ZT, 2 ZT)
    static constexpr int NB = 4, TEAM = 1;
    static constexpr bool OK = true;
    struct alignas(16) Small {
        double LB[ZT][8], UB[ZT][8];
    };
    struct Lane {};
    __device__ static __forceinline__ void lane_reset(Lane &) {}

    static inline void fill(const spcies_consts &C, Small &S, long double *F) {
        typedef long double ld;
        for (int t = 0; t < ZT; ++t)
            for (int c = 0; c < 8; ++c) {
                const int i = t * 8 + c, l = i / nm, e = i % nm;
                double lo = 0.0, hi = 0.0;
                if (i < L) {
                    if (l == 0 && e < n) {
                        lo = -spcies_inf_value;
                        hi = spcies_inf_value;
                    } else if (l < N) {
                        lo = (double)C.LB[e];
                        hi = (double)C.UB[e];
                    } else {
                        const double eps = e < n ? (double)eps_x : (double)eps_u;
                        lo = (double)C.LB[e] + eps;
                        hi = (double)C.UB[e] - eps;
                    }
                }
                S.LB[t][c] = lo;
                S.UB[t][c] = hi;
            }
        // KKT matrix (compute_MPCT_ADMM_semiband_ingredients.m:118-150: H = banded part + coupling with (x_s, u_s); G)
        constexpr int KD = L + LM;
        constexpr int NINC = (NW + NC) * 8;
        ld *K = new ld[(size_t)KD * KD]();
        ld *X = new ld[(size_t)KD * NINC]();
        auto Kat = [&](int r, int c) -> ld & { return K[(size_t)r * KD + c]; };
        auto QR = [&](int a, int b_) -> ld { return a < n ? (b_ < n ? (ld)C.Q[a][b_] : 0) : (b_ >= n ? (ld)C.R[a - n][b_ - n] : 0); };
        auto TS = [&](int a, int b_) -> ld { return a < n ? (b_ < n ? (ld)C.T[a][b_] : 0) : (b_ >= n ? (ld)C.S[a - n][b_ - n] : 0); };
        for (int l = 0; l < N; ++l)
            for (int a = 0; a < nm; ++a)
                for (int b_ = 0; b_ < nm; ++b_) {
                    Kat(l * nm + a, l * nm + b_) += QR(a, b_);
                    Kat(l * nm + a, N * nm + b_) += -QR(a, b_);
                    Kat(N * nm + a, l * nm + b_) += -QR(a, b_);
                    Kat(N * nm + a, N * nm + b_) += QR(a, b_);                 // N (Q, R) in total
                }
        for (int a = 0; a < nm; ++a)
            for (int b_ = 0; b_ < nm; ++b_) Kat(N * nm + a, N * nm + b_) += TS(a, b_);
        for (int i = 0; i < L; ++i) Kat(i, i) += (ld)rho;
        auto Gset = [&](int r, int c, ld v) {
            Kat(L + r, c) = v;
            Kat(c, L + r) = v;
        };
        for (int i = 0; i < n; ++i) Gset(i, i, 1);                                 // x_0 = x(t)
        for (int l = 0; l < N; ++l)                                                // x_{l+1} = A x_l + B u_l (x_N := x_s)
            for (int i = 0; i < n; ++i) {
                for (int j = 0; j < n; ++j) Gset((l + 1) * n + i, l * nm + j, (ld)C.A[i][j]);
                for (int j = 0; j < m; ++j) Gset((l + 1) * n + i, l * nm + n + j, (ld)C.B[i][j]);
                Kat(L + (l + 1) * n + i, (l + 1) * nm + i) += -1;
                Kat((l + 1) * nm + i, L + (l + 1) * n + i) += -1;
            }
        for (int i = 0; i < n; ++i) {                                              // (A - I) x_s + B u_s = 0
            for (int j = 0; j < n; ++j) {
                const ld v = (ld)C.A[i][j] - (i == j ? 1 : 0);
                Kat(L + (N + 1) * n + i, N * nm + j) += v;
                Kat(N * nm + j, L + (N + 1) * n + i) += v;
            }
            for (int j = 0; j < m; ++j) {
                Kat(L + (N + 1) * n + i, N * nm + n + j) += (ld)C.B[i][j];
                Kat(N * nm + n + j, L + (N + 1) * n + i) += (ld)C.B[i][j];
            }
        }
        // right-hand sides [-p; b]: w_j -> p = e_j;  x0_i -> b = e_i;  xr_i -> p = -T[:, i] on x_s;  ur_i -> p = -S[:, i] on u_s   (:91-109)
        for (int col = 0; col < NINC; ++col) {
            if (col < NW * 8) {
                if (col < L) X[(size_t)col * NINC + col] = -1;
            } else {
                const int e = col - NW * 8;
                if (e < n) X[(size_t)(L + e) * NINC + col] = 1;
                else if (e < 2 * n)
                    for (int j = 0; j < n; ++j) X[(size_t)(N * nm + j) * NINC + col] = (ld)C.T[j][e - n];
                else if (e < 2 * n + m)
                    for (int j = 0; j < m; ++j) X[(size_t)(N * nm + n + j) * NINC + col] = (ld)C.S[j][e - 2 * n];
            }
        }
        // Gaussian elimination with partial pivoting
        for (int c = 0; c < KD; ++c) {
            int piv = c;
            for (int r = c + 1; r < KD; ++r)
                if (fabsl(Kat(r, c)) > fabsl(Kat(piv, c))) piv = r;
            if (piv != c) {
                for (int j = 0; j < KD; ++j) {
                    const ld t = Kat(c, j);
                    Kat(c, j) = Kat(piv, j);
                    Kat(piv, j) = t;
                }
                for (int j = 0; j < NINC; ++j) {
                    const ld t = X[(size_t)c * NINC + j];
                    X[(size_t)c * NINC + j] = X[(size_t)piv * NINC + j];
                    X[(size_t)piv * NINC + j] = t;
                }
            }
            const ld d = 1 / Kat(c, c);
            for (int r = c + 1; r < KD; ++r) {
                const ld f = Kat(r, c) * d;
                if (f == 0) continue;
                for (int j = c; j < KD; ++j) Kat(r, j) -= f * Kat(c, j);
                for (int j = 0; j < NINC; ++j) X[(size_t)r * NINC + j] -= f * X[(size_t)c * NINC + j];
            }
        }
        for (int r = KD - 1; r >= 0; --r) {
            const ld d = 1 / Kat(r, r);
            for (int j = 0; j < NINC; ++j) {
                ld a = X[(size_t)r * NINC + j];
                for (int c = r + 1; c < KD; ++c) a -= Kat(r, c) * X[(size_t)c * NINC + j];
                X[(size_t)r * NINC + j] = a * d;
            }
        }
        for (int i = 0; i < L; ++i)
            for (int j = 0; j < NINC; ++j) F[(size_t)i * NINC + j] = X[(size_t)i * NINC + j];
        delete[] K;
        delete[] X;
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
    __device__ static __forceinline__ double2 make_w(Lane &, const spcies_consts *, const Small *, int t, const double2 *st, int) {
        const double2 v = st[t * 32], lam = st[(ZT + t) * 32];
        return make_double2(fma(-(double)rho, v.x, lam.x), fma(-(double)rho, v.y, lam.y));           // :141-149 without q
    }
    __device__ static __forceinline__ void update(Lane &, const spcies_consts *, const Small *S, int t0, const double (&acc)[NB][2],
                                                  double2 *st, int t4, bool &over) {
        const double tp = (double)tol_p, td = (double)tol_d;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int t = t0 + b;
            if (t >= ZT) break;
            const double2 vo = st[t * 32], lam = st[(ZT + t) * 32];
            const double2 lo = reinterpret_cast<const double2 *>(S->LB[t])[t4], hi = reinterpret_cast<const double2 *>(S->UB[t])[t4];
            const double z0 = acc[b][0], z1 = acc[b][1];
            const double v0 = clip(fma((double)rho_i, lam.x, z0), lo.x, hi.x), v1 = clip(fma((double)rho_i, lam.y, z1), lo.y, hi.y);   // :775-830
            st[t * 32] = make_double2(v0, v1);
            st[(ZT + t) * 32] = make_double2(fma((double)rho, z0 - v0, lam.x), fma((double)rho, z1 - v1, lam.y));                       // :1052-1062
            over = over || (fabs(v0 - vo.x) > td) || (fabs(z0 - v0) > tp) || (fabs(v1 - vo.y) > td) || (fabs(z1 - v1) > tp);
        }
    }
    __device__ static __forceinline__ void finish(Lane &, const spcies_consts *C, const BatchIO &io, long long inst, const double2 *st,
                                                  int t4) {
#pragma unroll
        for (int j = 0; j < m; ++j) {                                                   // u_opt = v[n + j]   :1140-1149
            const int e = n + j;
            if ((e % 8) / 2 == t4) {
                const double2 v = st[(e / 8) * 32];
                io.u[inst * m + j] = eng_u_out(C, (e & 1) ? v.y : v.x, j);
            }
        }
    }
};
