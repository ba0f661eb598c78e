// HMPC_ADMM_split_mma.cuh -- tensor-core engine policy of the HMPC (S)ADMM_split solver for spcies_dense_mma.cuh (included by
// HMPC_ADMM_split.cuh, inside spcies::hmpc).
//
// The hot loop of code_HMPC_ADMM_split_C.c:176-188 is the dense product primal_hat = M2 bh - M1 q_hat with M1 = (dim + n_s)^2 =
// 442^2 at N = 50: 195 k sequential FMA per instance and iteration for a thread that owns an instance, reading a 1.5 MB matrix.
// For a batch that shares M1 it is a GEMM.  With  w = (sigma z - lambda ; rho s - mu)  and  c = (x0, xr, ur)  (bh = -A x0 and the
// 2 n + m non-zeros of q -- -Te xr - QQ x0 at x_e, -QQ x0 at x_c, -Se ur at u_e, :99-129 -- folded into the columns of F):
//     primal_hat = F [w ; c],   F = [-M1 | -M2 A - M1(:, x_e) QQ - M1(:, x_c) QQ | -M1(:, x_e) Te | -M1(:, u_e) Se].
// Round 1 streamed the fragment table once per warp through per-warp cp.async rings (two warps per SM at N = 50, 0.29 of the
// FP64 ceiling); on the dense engine one producer warp streams it once per CTA with bulk copies and a group of 8 instances is
// served by a team of two warps, so all four SM schedulers issue.
//
// Tile order: 0,1,2 = the s part re-ordered as (y_e, y_s, y_c) x 8 triples (s[3 g + c] -> tile c, column g: a lane owns complete
// triples for the diamond-set projections, :255-258), 3.. = z (z_j -> tile 3 + j / 8, column j % 8).  Iterates per group: primal,
// dual.  Arithmetic: FAST (FMA, MMA accumulation order); EXACT, float and the debug payload use the scalar kernel.
#pragma once
// (spcies_dense_mma.cuh is included by the parent header, outside its namespace)

struct Engine {
    static constexpr int ZT = (DIM + 7) / 8;          // tiles of z
    static constexpr int NT = ZT + 3;
    static constexpr int NO = NT, NW = NT, NC = (2 * n + m + 7) / 8;
    static constexpr int NSTATE = 2 * NT;             // primal at [0, NT), dual at [NT, 2 NT)
    static constexpr int NB = 4;
    static constexpr int TEAM = (NSTATE + NW + NC + 1) * 512 > 60 * 1024 ? 2 : 1;
    static constexpr int NCLIP = DIM - 3 * n - 3 * m; // clipped entries of z (:233)
    static constexpr bool OK = NS == 3 * nm && nm <= 8;
    struct alignas(16) Small {
        double LB[ZT][8], UB[ZT][8];                  // bounds of z by tile / column (+-1e300 where z is not clipped)
        double LBy[8], UBy[8];
    };
    struct Lane {};
    __device__ static __forceinline__ void lane_reset(Lane &) {}

    // element of (z, s) held by (tile, column), or -1
    static inline int elem_at(int tile, int col) {
        if (tile < 3) return col < nm ? DIM + 3 * col + tile : -1;
        const int j = (tile - 3) * 8 + col;
        return j < DIM ? j : -1;
    }

    static inline void fill(const spcies_consts &C, Small &S, long double *F) {
        typedef long double ld;
        for (int t = 0; t < ZT; ++t)
            for (int c = 0; c < 8; ++c) {
                const int j = t * 8 + c;
                S.LB[t][c] = j < NCLIP ? (double)C.LB[j] : (j < DIM ? -1e300 : 0.0);
                S.UB[t][c] = j < NCLIP ? (double)C.UB[j] : (j < DIM ? 1e300 : 0.0);
            }
        for (int g = 0; g < 8; ++g) {
            S.LBy[g] = g < nm ? (double)C.LBy[g] : 0.0;
            S.UBy[g] = g < nm ? (double)C.UBy[g] : 0.0;
        }
        constexpr int NINC = (NW + NC) * 8;
#ifdef NON_SPARSE
        for (int ot = 0; ot < NO; ++ot)
            for (int o = 0; o < 8; ++o) {
                const int row = elem_at(ot, o);
                if (row < 0) continue;
                ld *fr = F + (size_t)(ot * 8 + o) * NINC;
                for (int it = 0; it < NW; ++it)
                    for (int c = 0; c < 8; ++c) {
                        const int col = elem_at(it, c);
                        if (col >= 0) fr[it * 8 + c] = -(ld)C.M1[row][col];
                    }
                for (int e = 0; e < 2 * n + m; ++e) {
                    ld a = 0;
                    if (e < n) {                      // x0
                        for (int j = 0; j < n; ++j)
                            a += -(ld)C.M2[row][j] * (ld)C.A[j][e] - ((ld)C.M1[row][Q0 + j] + (ld)C.M1[row][Q0 + 2 * n + j]) * (ld)C.QQ[j][e];
                    } else if (e < 2 * n) {           // xr
                        for (int j = 0; j < n; ++j) a += -(ld)C.M1[row][Q0 + j] * (ld)C.Te[j][e - n];
                    } else {                          // ur
                        for (int j = 0; j < m; ++j) a += -(ld)C.M1[row][Q0 + 3 * n + j] * (ld)C.Se[j][e - 2 * n];
                    }
                    fr[NW * 8 + e] = a;
                }
            }
#else
        // sparse = true: the same map, column by column, through the generated L D L' factors of the KKT matrix: the head of
        // K^-1 (q_hat; bh) for unit inputs (q_hat = w - q on the z part, bh[idx_x0] = -A x0)       code_HMPC_ADMM_split_C.c:161-209
        constexpr int NM = nrow_M;
        ld *rhs = new ld[NM];
        auto solve = [&]() {
            for (int i = 0; i < NM; ++i)
                for (int j = C.L_col[i]; j < C.L_col[i + 1]; ++j) rhs[C.L_row[j]] -= (ld)C.L_val[j] * rhs[i];
            for (int i = 0; i < NM; ++i) rhs[i] *= (ld)C.Dinv[i];
            for (int i = NM - 1; i >= 0; --i)
                for (int j = C.L_col[i]; j < C.L_col[i + 1]; ++j) rhs[i] -= (ld)C.L_val[j] * rhs[C.L_row[j]];
        };
        for (int col = 0; col < NINC; ++col) {
            for (int i = 0; i < NM; ++i) rhs[i] = 0;
            bool used = false;
            if (col < NW * 8) {
                const int e = elem_at(col / 8, col % 8);
                if (e >= 0) {
                    rhs[e] = 1;
                    used = true;
                }
            } else {
                const int e = col - NW * 8;
                if (e < n) {                          // x0: bh = -A x0;  q_hat -= q,  q_e = -QQ x0, q_c = -QQ x0
                    for (int j = 0; j < n; ++j) {
                        rhs[NP + C.idx_x0[j]] = -(ld)C.A[j][e];
                        rhs[Q0 + j] = (ld)C.QQ[j][e];
                        rhs[Q0 + 2 * n + j] = (ld)C.QQ[j][e];
                    }
                    used = true;
                } else if (e < 2 * n) {               // xr: q_e = -Te xr
                    for (int j = 0; j < n; ++j) rhs[Q0 + j] = (ld)C.Te[j][e - n];
                    used = true;
                } else if (e < 2 * n + m) {           // ur: q_ue = -Se ur
                    for (int j = 0; j < m; ++j) rhs[Q0 + 3 * n + j] = (ld)C.Se[j][e - 2 * n];
                    used = true;
                }
            }
            if (!used) continue;
            solve();
            for (int ot = 0; ot < NO; ++ot)
                for (int o = 0; o < 8; ++o) {
                    const int row = elem_at(ot, o);
                    if (row >= 0) F[(size_t)(ot * 8 + o) * NINC + col] = rhs[row];
                }
        }
        delete[] rhs;
#endif
    }

    __device__ static __forceinline__ void init(Lane &, const spcies_consts *C, const Small *, const BatchIO &io, long long inst,
                                                double2 *st, double2 *cin, int t4, int rank) {
#pragma unroll 4
        for (int t = rank; t < NSTATE; t += TEAM) st[t * 32] = make_double2(0.0, 0.0);
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
    __device__ static __forceinline__ double2 make_w(Lane &, const spcies_consts *C, const Small *, int t, const double2 *st, int) {
        const double pen = t < 3 ? (double)C->rho : (double)C->sigma;                                    // :149-154 without q
        const double2 p = st[t * 32], d = st[(NT + t) * 32];
        return make_double2(fma(pen, p.x, -d.x), fma(pen, p.y, -d.y));
    }
    __device__ static __forceinline__ void update(Lane &, const spcies_consts *C, const Small *S, int t0, const double (&acc)[NB][2],
                                                  double2 *st, int t4, bool &over) {
        typedef Arith<double, false> A;
        const double sigma_ = C->sigma, sigma_i_ = C->sigma_i, rho_ = C->rho, rho_i_ = C->rho_i;
        const double as = SYMMETRIC ? (double)SPCIES_ALPHA * sigma_ : sigma_;      // alpha_SADMM * sigma
        const double ar = SYMMETRIC ? (double)SPCIES_ALPHA * rho_ : rho_;
        const double told = (double)tol_d, tolp = (double)tol_p;
        // z tile: [SADMM dual step], z = clip(z_hat + lambda / sigma), dual step, exit tests        :215-219, :230-238, :288-334
        auto update_z = [&](int t, const double (&zh)[2]) {
            const double2 zo = st[t * 32], lam0 = st[(NT + t) * 32];
            const double2 lo = reinterpret_cast<const double2 *>(S->LB[t - 3])[t4], hi = reinterpret_cast<const double2 *>(S->UB[t - 3])[t4];
            const double zov[2] = {zo.x, zo.y}, lov[2] = {lo.x, lo.y}, hiv[2] = {hi.x, hi.y};
            double lam[2] = {lam0.x, lam0.y}, z[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (SYMMETRIC) lam[i] = fma(as, zh[i] - zov[i], lam[i]);
                z[i] = clip(fma(sigma_i_, lam[i], zh[i]), lov[i], hiv[i]);
                lam[i] = fma(as, zh[i] - z[i], lam[i]);
                over = over || (fabs(zov[i] - z[i]) > told) || (fabs(z[i] - zh[i]) > tolp);
            }
            st[t * 32] = make_double2(z[0], z[1]);
            st[(NT + t) * 32] = make_double2(lam[0], lam[1]);
        };
        int b0 = 0;
        if (t0 == 0) {
            // s, mu: diamond-set projection of each (y_e, y_s, y_c) triple                              :220-225, :241-258, :294-311
            const double2 so0 = st[0], so1 = st[32], so2 = st[64];
            const double2 mu0 = st[NT * 32], mu1 = st[(NT + 1) * 32], mu2 = st[(NT + 2) * 32];
            const double so[3][2] = {{so0.x, so0.y}, {so1.x, so1.y}, {so2.x, so2.y}};
            double mu[3][2] = {{mu0.x, mu0.y}, {mu1.x, mu1.y}, {mu2.x, mu2.y}}, sn[3][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int g = 2 * t4 + i;
                const bool trip = g < nm;
                double sv[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (SYMMETRIC) mu[c][i] = fma(ar, acc[c][i] - so[c][i], mu[c][i]);
                    sv[c] = fma(rho_i_, mu[c][i], acc[c][i]);
                }
                proj_soc3<A>(sv, 1.0, S->LBy[g]);
                proj_soc3<A>(sv, -1.0, S->UBy[g]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    sn[c][i] = trip ? sv[c] : 0.0;
                    mu[c][i] = trip ? fma(ar, acc[c][i] - sv[c], mu[c][i]) : 0.0;
                    over = over || (trip && ((fabs(so[c][i] - sv[c]) > told) || (fabs(sv[c] - acc[c][i]) > tolp)));
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                st[c * 32] = make_double2(sn[c][0], sn[c][1]);
                st[(NT + c) * 32] = make_double2(mu[c][0], mu[c][1]);
            }
            b0 = 3;
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b < b0) continue;
            const int t = t0 + b;
            if (t >= NT) break;
            update_z(t, acc[b]);
        }
    }
    __device__ static __forceinline__ void finish(Lane &, const spcies_consts *C, const BatchIO &io, long long inst, const double2 *st,
                                                  int t4) {
        const double2 z0 = st[3 * 32];                                                                  // u_opt = z[0..m)   (:359-368)
        if (2 * t4 < m) io.u[inst * m + 2 * t4] = eng_u_out(C, z0.x, 2 * t4);
        if (2 * t4 + 1 < m) io.u[inst * m + 2 * t4 + 1] = eng_u_out(C, z0.y, 2 * t4 + 1);
    }
};
