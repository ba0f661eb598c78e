// HMPC_ADMM_mma.cuh -- tensor-core engine policy of the non-split HMPC / ellipHMPC ADMM solver for spcies_dense_mma.cuh
// (included by HMPC_ADMM.cuh, inside spcies::hmpc_ns).
//
// One iteration of code_HMPC_ADMM_C.c:122-169 maps w = rho s + lambda (n_s entries) to Cz = C z with
//     z = M2 b + M1 (q + C' w),
// i.e. the iteration lives in the constraint space and  Cz = (C M1 C') w + C (M2 b + M1 q)  is one linear map of [w ; c], where
// c = (x0, xr, ur [, x_rs, x_rc, u_rs, u_rc]) are the raw inputs (A, QQ, Te, Se, Th, Sh are folded into the columns of F on the
// host, in extended precision).  One extra output tile carries z[0..8) = the rows of the same map without C: u_opt.
//
// Tile order of the outputs: 0,1,2 = the cone part re-ordered as (y_e, y_s, y_c) x 8 triples (a lane owns complete triples: the
// diamond-set projections need no shuffles), 3 = z[0..8), 4.. = the box part.  Inputs: the same without tile 3.
// Iterates per group: s, lambda (by s tile), and the u tile.
#pragma once
// (spcies_dense_mma.cuh is included by the parent header, outside its namespace)

struct Engine {
    static constexpr int BT = (NBOX + 7) / 8;              // box tiles
    static constexpr int NTS = 3 + BT;                     // tiles of s: cone (3), box (BT)
    static constexpr int NO = NTS + 1, NW = NTS;
    static constexpr int NCE = n + SPCIES_NREF * (n + m);  // entries of c
    static constexpr int NC = (NCE + 7) / 8;
    static constexpr int NSTATE = 2 * NTS + 1;             // s at [0, NTS), lambda at [NTS, 2 NTS), u tile at 2 NTS
    static constexpr int NB = 4;
    static constexpr int TEAM = (NSTATE + NW + NC + 1) * 512 > 60 * 1024 ? 2 : 1;
    static constexpr bool OK = !SOC && NCONE <= 8 && nrow_C == NS && nrow_Ct == DIM;
    struct alignas(16) Small {
        double LB[BT][8], UB[BT][8];                       // box bounds by box tile / column
        double LBy[8], UBy[8];
    };
    struct Lane {};
    __device__ static __forceinline__ void lane_reset(Lane &) {}

    // element of s held by (s tile, column), or -1
    static inline int s_elem(int tile, int col) {
        if (tile < 3) return col < NCONE ? NBOX + 3 * col + tile : -1;
        const int j = (tile - 3) * 8 + col;
        return j < NBOX ? j : -1;
    }

    static inline void fill(const spcies_consts &C, Small &S, long double *F) {
        typedef long double ld;
        for (int t = 0; t < BT; ++t)
            for (int c = 0; c < 8; ++c) {
                const int j = t * 8 + c;
                S.LB[t][c] = j < NBOX ? (double)C.LB[j] : 0.0;
                S.UB[t][c] = j < NBOX ? (double)C.UB[j] : 0.0;
            }
        for (int g = 0; g < 8; ++g) {
            S.LBy[g] = g < NCONE ? (double)C.LBy[g] : 0.0;
            S.UBy[g] = g < NCONE ? (double)C.UBy[g] : 0.0;
        }
        constexpr int NINC = (NW + NC) * 8;
        // z as a function of (w, c): columns of Z = [M1 C' | dz/dc]
        ld *Z = new ld[(size_t)DIM * NINC]();
        for (int it = 0; it < NW; ++it)
            for (int c = 0; c < 8; ++c) {
                const int e = s_elem(it, c);
                if (e < 0) continue;
                // C' e_e = row e of C
                for (int j = C.C_row[e]; j < C.C_row[e + 1]; ++j) {
                    const int col = C.C_col[j];
                    const ld cv = (ld)C.C_val[j];
                    for (int i = 0; i < DIM; ++i) Z[(size_t)i * NINC + it * 8 + c] += (ld)C.M1[i][col] * cv;
                }
            }
        for (int e = 0; e < NCE; ++e) {
            const int col = NW * 8 + e;
            ld q[NQ] = {}, b[n] = {};
            if (e < n) {                                                   // x0: b = -A x0, q_e -= QQ x0, q_c -= QQ x0
                for (int j = 0; j < n; ++j) {
                    b[j] = -(ld)C.A[j][e];
                    q[j] = -(ld)C.QQ[j][e];
                    q[2 * n + j] = -(ld)C.QQ[j][e];
                }
            } else if (e < 2 * n) {                                        // xr (x_re): q_e -= Te xr
                for (int j = 0; j < n; ++j) q[j] = -(ld)C.Te[j][e - n];
            } else if (e < 2 * n + m) {                                    // ur (u_re): q_ue -= Se ur
                for (int j = 0; j < m; ++j) q[3 * n + j] = -(ld)C.Se[j][e - 2 * n];
            }
#if SPCIES_NREF == 3
            else if (e < 3 * n + m) {                                      // x_rs: q_s -= Th x_rs
                for (int j = 0; j < n; ++j) q[n + j] = -(ld)C.Th[j][e - 2 * n - m];
            } else if (e < 4 * n + m) {                                    // x_rc: q_c -= Th x_rc
                for (int j = 0; j < n; ++j) q[2 * n + j] = -(ld)C.Th[j][e - 3 * n - m];
            } else if (e < 4 * n + 2 * m) {                                // u_rs
                for (int j = 0; j < m; ++j) q[3 * n + m + j] = -(ld)C.Sh[j][e - 4 * n - m];
            } else {                                                       // u_rc
                for (int j = 0; j < m; ++j) q[3 * n + 2 * m + j] = -(ld)C.Sh[j][e - 4 * n - 2 * m];
            }
#endif
            for (int i = 0; i < DIM; ++i) {
                ld a = 0;
                for (int j = 0; j < n; ++j) a += (ld)C.M2[i][j] * b[j];
                for (int j = 0; j < NQ; ++j) a += (ld)C.M1[i][Q0 + j] * q[j];
                Z[(size_t)i * NINC + col] = a;
            }
        }
        // outputs: Cz rows for the s tiles, z[0..8) for tile 3
        for (int ot = 0; ot < NO; ++ot)
            for (int o = 0; o < 8; ++o) {
                ld *row = F + (size_t)(ot * 8 + o) * NINC;
                if (ot == 3) {
                    if (o < DIM)
                        for (int c = 0; c < NINC; ++c) row[c] = Z[(size_t)o * NINC + c];
                    continue;
                }
                const int e = s_elem(ot < 3 ? ot : ot - 1, o);
                if (e < 0) continue;
                for (int j = C.C_row[e]; j < C.C_row[e + 1]; ++j) {
                    const ld cv = (ld)C.C_val[j];
                    const ld *zr = Z + (size_t)C.C_col[j] * NINC;
                    for (int c = 0; c < NINC; ++c) row[c] += cv * zr[c];
                }
            }
        delete[] Z;
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
                double x = 0.0;
                if (e < n) x = eng_x(C, io.x0, inst, n, e);
                else if (e < 2 * n) x = eng_x(C, io.xr, inst, n, e - n);
                else if (e < 2 * n + m) x = eng_u(C, io.ur, inst, m, e - 2 * n);
#if SPCIES_NREF == 3
                else if (e < 3 * n + m) x = io.ex[0][inst * n + e - 2 * n - m];
                else if (e < 4 * n + m) x = io.ex[1][inst * n + e - 3 * n - m];
                else if (e < 4 * n + 2 * m) x = io.ex[2][inst * m + e - 4 * n - m];
                else if (e < 4 * n + 3 * m) x = io.ex[3][inst * m + e - 4 * n - 2 * m];
#endif
                v[i] = x;
            }
            cin[t * 32] = make_double2(v[0], v[1]);
        }
    }
    __device__ static __forceinline__ double2 make_w(Lane &, const spcies_consts *C, const Small *, int t, const double2 *st, int) {
        const double rho_ = C->rho;
        const double2 s = st[t * 32], lam = st[(NTS + t) * 32];
        return make_double2(fma(rho_, s.x, lam.x), fma(rho_, s.y, lam.y));                              // :125-131
    }
    __device__ static __forceinline__ void update(Lane &, const spcies_consts *C, const Small *S, int t0, const double (&acc)[NB][2],
                                                  double2 *st, int t4, bool &over) {
        typedef Arith<double, false> A;
        const double rho_ = C->rho, rho_i_ = C->rho_i;
        const double ar = SYMMETRIC ? (double)SPCIES_ALPHA * rho_ : rho_;
        const double tp = (double)tol_p, td = (double)tol_d;
        if (t0 == 0) {
            // cone part: tiles 0, 1, 2 = (y_e, y_s, y_c) of the triples in columns 2 t4, 2 t4 + 1; tile 3 = z[0..8)
            const double2 so0 = st[0], so1 = st[32], so2 = st[64];
            const double2 l0 = st[NTS * 32], l1 = st[(NTS + 1) * 32], l2 = st[(NTS + 2) * 32];
            const double so[3][2] = {{so0.x, so0.y}, {so1.x, so1.y}, {so2.x, so2.y}};
            double lam[3][2] = {{l0.x, l0.y}, {l1.x, l1.y}, {l2.x, l2.y}}, sn[3][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int g = 2 * t4 + i;
                double sv[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (SYMMETRIC) lam[c][i] = fma(ar, acc[c][i] + so[c][i], lam[c][i]);              // :173-180
                    sv[c] = -acc[c][i] - rho_i_ * lam[c][i];                                         // :184-186
                }
                proj_soc3<A>(sv, 1.0, S->LBy[g]);                                                    // :202-205
                proj_soc3<A>(sv, -1.0, S->UBy[g]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double cz2 = acc[c][i] + sv[c];                                            // :209-211
                    sn[c][i] = sv[c];
                    lam[c][i] = fma(ar, cz2, lam[c][i]);                                             // :215-229
                    over = over || (g < NCONE && ((fabs(cz2) > tp) || (fabs(sv[c] - so[c][i]) > td)));
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                st[c * 32] = make_double2(sn[c][0], sn[c][1]);
                st[(NTS + c) * 32] = make_double2(lam[c][0], lam[c][1]);
            }
            st[(2 * NTS) * 32] = make_double2(acc[3][0], acc[3][1]);                                 // z[0..8): u_opt at exit
            return;
        }
        // box part
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int ts = t0 + b - 1;                  // s tile (output tile t0 + b, the u tile sits at output index 3)
            if (ts >= NTS) break;
            const double2 so = st[ts * 32], l = st[(NTS + ts) * 32];
            const double2 lo = reinterpret_cast<const double2 *>(S->LB[ts - 3])[t4], hi = reinterpret_cast<const double2 *>(S->UB[ts - 3])[t4];
            const double sov[2] = {so.x, so.y}, lov[2] = {lo.x, lo.y}, hiv[2] = {hi.x, hi.y};
            double lam[2] = {l.x, l.y}, sn[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double cz = acc[b][i];
                if (SYMMETRIC) lam[i] = fma(ar, cz + sov[i], lam[i]);
                sn[i] = clip(-cz - rho_i_ * lam[i], lov[i], hiv[i]);                                 // :184-194
                const double cz2 = cz + sn[i];
                lam[i] = fma(ar, cz2, lam[i]);
                over = over || (fabs(cz2) > tp) || (fabs(sn[i] - sov[i]) > td);
            }
            st[ts * 32] = make_double2(sn[0], sn[1]);
            st[(NTS + ts) * 32] = make_double2(lam[0], lam[1]);
        }
    }
    __device__ static __forceinline__ void finish(Lane &, const spcies_consts *C, const BatchIO &io, long long inst, const double2 *st,
                                                  int t4) {
        const double2 z0 = st[(2 * NTS) * 32];                                                      // u_opt = z[0..m)   :271-282
        if (2 * t4 < m) io.u[inst * m + 2 * t4] = eng_u_out(C, z0.x, 2 * t4);
        if (2 * t4 + 1 < m) io.u[inst * m + 2 * t4 + 1] = eng_u_out(C, z0.y, 2 * t4 + 1);
    }
};
