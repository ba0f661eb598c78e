// ellipMPC_ADMM_soc_band.cuh -- structured tensor-core (DMMA) engine of the ellipMPC ADMM_soc solver (included by
// ellipMPC_ADMM_soc.cuh, inside spcies::soc, after the dense engine ellipMPC_ADMM_soc_mma.cuh).
//
// The dense engine applies the whole map primal_hat = M1 q_hat - Mb bh (88 x 88 at N = 10: 336 DMMA per warp-iteration).  The map
// has the structure of the other MPC solvers (compute_ellipMPC_ADMM_soc_ingredients.m:60-147):
//   * z = (u_0, x_1, u_1, ..., x_{N-1}, u_{N-1}, x_N, t), s = (s_0, s_1..s_n); Hh is diagonal but for the block of x_N (T + sigma I);
//   * Gh = dynamics rows (the equMPC ones), the row t = r, the row s_0 - t = 0 and the rows s_{1..n} - P^(1/2) x_N = -P^(-1/2) P xr.
//     The two rows with t decouple from the rest of W = Gh Hh^-1 Gh' and pin t_hat = s_hat_0 = r; the rest of W is block
//     tridiagonal with N + 1 blocks of n rows (the N dynamics blocks + the cone rows, coupled through x_N).
// So one iteration is the equMPC ADMM engine (MPC_ADMM_mma.cuh: merged block recurrences, blocks as FP64 MMA B fragments) with a
// dense terminal block and one more block in the recurrences:
//   pass A  q_hat_b = q + lambda_b - sigma z_b,  s_b = Hi o q_hat_b (s_0: state columns = -x0),  s_N = Hi_N (qT - sigma z_N)
//           r_b = s_{b+1} - [A B] s_b  (b < N),   r_N = P^(1/2) s_N + s_s - mu_s / rho + PhiP xr
//           nu'_b = Linv_b r_b - F_b nu'_{b-1}                                                  b = 0..N
//   pass B  nu_b = Uinv_b nu'_b - G_b nu_{b+1}
//           s_hat = s_s - (mu_s + nu_N) / rho,  s = proj_SOC(s_hat + mu / rho),  mu += rho (s_hat - s)     (:220-254)
//           z_N = z_hat_N = -Hi_N (q_hat_N - nu_{N-1} - P^(1/2)' nu_N)     (x_N is not clipped: lambda_N stays exactly 0, :210-217)
//           z_hat_b = -Hi o (q_hat_b - [nu_{b-1}; 0] + [A B]' nu_b),  z_b = clip(z_hat_b + lambda_b / sigma),  lambda_b += sigma (z_hat_b - z_b)
//           exit test |primal_prev - primal| > tol_d or |primal - primal_hat| > tol_p              (:260-281)
// 10 N + 12 = 112 DMMA per warp-iteration at N = 10.  t (z_hat_t = z_t = r, lambda_t = 0) is not stored.
// 16 warps per SM (512 threads at 128 registers; measured on C4, 1 Mi instances: 7.69 / 8.17 / 8.81 / 9.12 M solves/s at 256 / 320 /
// 384 / 512 threads -- the recurrences are latency bound, more resident warps hide them), 11.5 KB of iterates per warp.
//
// The tables are derived on the host from the generated CSR / CSC constants (Gh = GhHhi Hh, W = GhHhi Gh', block Cholesky in
// extended precision); if the structure above is not found (checked entry by entry) the dense engine keeps the solver.
// Arithmetic: FAST; EXACT mode, float and the debug payload use the scalar kernel.
#pragma once

#ifndef SPCIES_SOC_BAND
#define SPCIES_SOC_BAND 1
#endif
#ifndef SPCIES_SOC_BAND_BLOCK
#define SPCIES_SOC_BAND_BLOCK 512
#endif

constexpr int BAND_BLOCK = SPCIES_SOC_BAND_BLOCK;
constexpr int BAND_IPB = BAND_BLOCK / 4;
constexpr int NBK = N + 1;                                   // blocks of the recurrences: N dynamics blocks + the cone rows
constexpr bool BAND_SHAPE_OK = mma::MmaLayout<n, m>::OK && N >= 3 && NS == n + 1 && DIM == N * nm + 1 && NEQ == N * n + 1;
constexpr int BAND_NST = 2 * N + 3;                          // z_b, lambda_b (b < N), z_N, s, mu
constexpr size_t BAND_STATE_PER_WARP = (size_t)BAND_NST * 32 * sizeof(double2);

struct alignas(16) BandTables {
    // by column (spcies_mma.cuh: MmaLayout); matrices row-major [output column][input column]
    double NAB[64];               // -[A B]          (out x, in z)
    double ABt[64];               //  [A B]'         (out z, in x)
    double HiN[64], NHiN[64];     // (T + sigma I)^-1 and its negative (x columns)
    double Ph[64], NPht[64];      // P^(1/2) and -P^(1/2)'
    double2 FWa[NBK][32], BWa[NBK][32];
    double FWb[NBK][32], BWb[NBK][32];
    double Hi[8], Qs[8];          // 1 / (Q + sigma), 1 / (R + sigma) and the (negated) diagonal of Q, R by column
    double LB[N][8], UB[N][8];    // bounds by block and column
    double Tm[8][8], PhiP[8][8];  // T (negated) and PhiP by [component][component]
    int xat[8], uat[8];
};
constexpr size_t BAND_BYTES = (sizeof(BandTables) + 15) / 16 * 16;
constexpr size_t BAND_OFFSET = HAS_MMA ? MMA_OFFSET + SMALL_BYTES + FRAG_BYTES : MMA_OFFSET;   // blob: consts | dense tables | band tables
constexpr size_t BAND_SMEM = BAND_BYTES + (BAND_BLOCK / 32) * BAND_STATE_PER_WARP;
constexpr bool HAS_BAND = SPCIES_SOC_BAND != 0 && BAND_SHAPE_OK && sizeof(SPCIES_REAL) == 8 && BAND_SMEM <= 227 * 1024 - 64 &&
                          nrow_GhHhi == NR && nrow_HhiGh == NP && nrow_Hhi == NP;

#if SPCIES_SOC_BAND
// false when the generated constants do not have the structure the engine relies on (then the dense engine is used)
static inline bool fill_band_tables(const spcies_consts &C, BandTables &T) {
    if constexpr (!BAND_SHAPE_OK) {
        return false;
    } else {
        typedef long double ld;
        typedef mma::MmaLayout<n, m> L;
        memset(&T, 0, sizeof T);
        const ld EPS = 1e-13L;
        auto at = [](ld *M, int cols, int i, int j) -> ld & { return M[(size_t)i * cols + j]; };
        ld *Hhi = new ld[(size_t)NP * NP](), *Hh = new ld[(size_t)NP * NP](), *GHi = new ld[(size_t)NR * NP]();
        ld *Gh = new ld[(size_t)NR * NP](), *W = new ld[(size_t)NR * NR]();
        bool ok = true;
        for (int i = 0; i < NP; ++i)
            for (int j = C.Hhi_row[i]; j < C.Hhi_row[i + 1]; ++j) at(Hhi, NP, i, C.Hhi_col[j]) = -(ld)C.Hhi_val[j];
        for (int i = 0; i < NR; ++i)
            for (int j = C.GhHhi_row[i]; j < C.GhHhi_row[i + 1]; ++j) at(GHi, NP, i, C.GhHhi_col[j]) = -(ld)C.GhHhi_val[j];
        {   // Hh = Hhi^-1 by Gauss-Jordan (Hhi is block diagonal, symmetric positive definite)
            ld *Aug = new ld[(size_t)NP * 2 * NP]();
            for (int i = 0; i < NP; ++i) {
                for (int j = 0; j < NP; ++j) at(Aug, 2 * NP, i, j) = at(Hhi, NP, i, j);
                at(Aug, 2 * NP, i, NP + i) = 1;
            }
            for (int p = 0; p < NP && ok; ++p) {
                const ld piv = at(Aug, 2 * NP, p, p);
                if (!(piv > 0)) {
                    ok = false;
                    break;
                }
                for (int j = 0; j < 2 * NP; ++j) at(Aug, 2 * NP, p, j) /= piv;
                for (int i = 0; i < NP; ++i) {
                    if (i == p) continue;
                    const ld f = at(Aug, 2 * NP, i, p);
                    if (f == 0) continue;
                    for (int j = 0; j < 2 * NP; ++j) at(Aug, 2 * NP, i, j) -= f * at(Aug, 2 * NP, p, j);
                }
            }
            for (int i = 0; i < NP; ++i)
                for (int j = 0; j < NP; ++j) at(Hh, NP, i, j) = at(Aug, 2 * NP, i, NP + j);
            delete[] Aug;
        }
        for (int i = 0; i < NR; ++i)
            for (int k = 0; k < NP; ++k) {
                const ld g = at(GHi, NP, i, k);
                if (g == 0) continue;
                for (int j = 0; j < NP; ++j) at(Gh, NP, i, j) += g * at(Hh, NP, k, j);
            }
        for (int i = 0; i < NR; ++i)
            for (int j = 0; j < NR; ++j) {
                ld v = 0;
                for (int k = 0; k < NP; ++k) v += at(GHi, NP, i, k) * at(Gh, NP, j, k);
                at(W, NR, i, j) = v;
            }
        // element indices: x_b at zx(b), u_b at zu(b) (b = 0: u only), x_N at zx(N), t at DIM - 1, s_j at DIM + j
        auto zx = [](int b) { return m + (b - 1) * nm; };
        auto zu = [](int b) { return b == 0 ? 0 : m + (b - 1) * nm + n; };
        const int ROW_T = N * n, ROW_S0 = NEQ;                 // the two rows that pin t_hat = s_hat_0 = r
        int rows[NBK * n];                                     // rows of the coupled part of W, by block
        for (int b = 0; b < N; ++b)
            for (int j = 0; j < n; ++j) rows[b * n + j] = b * n + j;
        for (int j = 0; j < n; ++j) rows[N * n + j] = NEQ + 1 + j;
        // (1) the diagonal of Hh is stage-uniform, its only dense block is the one of x_N
        for (int i = 0; i < NP && ok; ++i)
            for (int j = 0; j < NP; ++j) {
                const bool inN = i >= zx(N) && i < zx(N) + n && j >= zx(N) && j < zx(N) + n;
                if (i != j && !inN && fabsl(at(Hhi, NP, i, j)) > 0) ok = false;
            }
        for (int b = 2; b < N && ok; ++b)
            for (int e = 0; e < nm; ++e)
                if (at(Hhi, NP, zx(b) + e, zx(b) + e) != at(Hhi, NP, zx(1) + e, zx(1) + e)) ok = false;
        for (int j = 0; j < m && ok; ++j)
            if (at(Hhi, NP, j, j) != at(Hhi, NP, zu(1) + j, zu(1) + j)) ok = false;
        for (int j = 0; j < NS && ok; ++j)
            if (fabsl(at(Hhi, NP, DIM + j, DIM + j) - (ld)C.rho_i) > EPS) ok = false;
        // (2) Gh: dynamics rows [A B] x_b u_b - x_{b+1}, t row, cone rows
        for (int b = 0; b < N && ok; ++b)
            for (int j = 0; j < n && ok; ++j)
                for (int c = 0; c < NP && ok; ++c) {
                    ld want = 0;
                    if (b >= 1 && c >= zx(b) && c < zx(b) + nm) want = at(Gh, NP, n + j, zx(1) + (c - zx(b)));       // [A B] of block 1
                    if (b == 0 && c < m) want = at(Gh, NP, n + j, zu(1) + c);
                    if (c >= zx(b + 1) && c < zx(b + 1) + n) want = (c - zx(b + 1) == j) ? -1 : 0;
                    if (fabsl(at(Gh, NP, b * n + j, c) - want) > EPS) ok = false;
                }
        for (int c = 0; c < NP && ok; ++c) {
            if (fabsl(at(Gh, NP, ROW_T, c) - (c == DIM - 1 ? 1 : 0)) > EPS) ok = false;
            if (fabsl(at(Gh, NP, ROW_S0, c) - (c == DIM - 1 ? -1 : (c == DIM ? 1 : 0))) > EPS) ok = false;
        }
        for (int j = 0; j < n && ok; ++j)
            for (int c = 0; c < NP && ok; ++c) {
                const bool inN = c >= zx(N) && c < zx(N) + n;
                const ld want = (c == DIM + 1 + j) ? 1 : 0;
                if (!inN && fabsl(at(Gh, NP, NEQ + 1 + j, c) - want) > EPS) ok = false;
            }
        // (3) block Cholesky W_c = R'R of the coupled rows (upper, block bidiagonal)
        constexpr int NC = NBK * n;
        ld *R = new ld[(size_t)NC * NC]();
        for (int i = 0; i < NC && ok; ++i)
            for (int j = i; j < NC; ++j) {
                ld v = at(W, NR, rows[i], rows[j]);
                for (int k = 0; k < i; ++k) v -= at(R, NC, k, i) * at(R, NC, k, j);
                if (i == j) {
                    if (!(v > 0)) {
                        ok = false;
                        break;
                    }
                    at(R, NC, i, i) = sqrtl(v);
                } else {
                    at(R, NC, i, j) = v / at(R, NC, i, i);
                }
            }
        for (int i = 0; i < NC && ok; ++i) {
            for (int j = (i / n + 2) * n; j < NC; ++j)
                if (fabsl(at(R, NC, i, j)) > EPS) ok = false;               // nothing beyond the first super-diagonal block
            if (fabsl(at(W, NR, rows[i], ROW_T)) > EPS || fabsl(at(W, NR, rows[i], ROW_S0)) > EPS) ok = false;   // the t rows decouple
        }
        if (ok) {
            struct Fac {
                double Beta[NBK][n][n], Alpha[NBK][n][n];
            };
            Fac *f = new Fac();
            for (int l = 0; l < NBK; ++l)
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < n; ++j) {
                        const ld u = at(R, NC, l * n + i, l * n + j);
                        f->Beta[l][i][j] = (double)(i == j ? 1 / u : u);                                   // diagonal stored inverted
                        f->Alpha[l][i][j] = l + 1 < NBK ? (double)at(R, NC, l * n + i, (l + 1) * n + j) : 0.0;
                    }
            typedef double Blk[n][n];
            Blk *Linv = new Blk[4 * NBK], *F = Linv + NBK, *Uinv = F + NBK, *G = Uinv + NBK;
            mma::block_inverses<NBK, n>(*f, Linv, F, Uinv, G);
            mma::recurrence_fragments<NBK, n, m, double>(Linv, F, Uinv, G, T.FWa, T.FWb, T.BWa, T.BWb);
            delete[] Linv;
            delete f;
            for (int c = 0; c < 8; ++c) {
                T.xat[c] = L::x_at(c);
                T.uat[c] = L::u_at(c);
                const int z = L::z_at(c);
                for (int b = 0; b < N; ++b) {
                    T.LB[b][c] = -1e300;
                    T.UB[b][c] = 1e300;
                    if (z < 0 || (b == 0 && z < n)) continue;
                    const int e = b == 0 ? z - n : zx(b) + z;
                    T.LB[b][c] = (double)C.LB[e];
                    T.UB[b][c] = (double)C.UB[e];
                }
                if (z < 0) continue;
                T.Hi[c] = (double)at(Hhi, NP, zx(1) + z, zx(1) + z);
                T.Qs[c] = z < n ? (double)C.Q[z][z] : (double)C.R[z - n][z - n];
            }
            for (int oc = 0; oc < 8; ++oc)
                for (int ic = 0; ic < 8; ++ic) {
                    const int xo = L::x_at(oc), xi = L::x_at(ic), zo = L::z_at(oc), zi = L::z_at(ic);
                    if (xo >= 0 && zi >= 0) T.NAB[oc * 8 + ic] = -(double)at(Gh, NP, n + xo, zx(1) + zi);
                    if (zo >= 0 && xi >= 0) T.ABt[oc * 8 + ic] = (double)at(Gh, NP, n + xi, zx(1) + zo);
                    if (xo >= 0 && xi >= 0) {
                        T.HiN[oc * 8 + ic] = (double)at(Hhi, NP, zx(N) + xo, zx(N) + xi);
                        T.NHiN[oc * 8 + ic] = -T.HiN[oc * 8 + ic];
                        T.Ph[oc * 8 + ic] = -(double)at(Gh, NP, NEQ + 1 + xo, zx(N) + xi);            // cone rows hold -P^(1/2)
                        T.NPht[oc * 8 + ic] = (double)at(Gh, NP, NEQ + 1 + xi, zx(N) + xo);
                    }
                }
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    T.Tm[i][j] = (double)C.T[i][j];
                    T.PhiP[i][j] = (double)C.PhiP[i][j];
                }
        }
        delete[] R;
        delete[] Hhi;
        delete[] Hh;
        delete[] GHi;
        delete[] Gh;
        delete[] W;
        return ok;
    }
}

__global__ void __launch_bounds__(BAND_BLOCK, 1) soc_band_kernel(const BatchIO io, const unsigned char *__restrict__ g_blob) {
    using mma::dmma;
    typedef mma::MmaLayout<n, m> L;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    const spcies_consts *C = reinterpret_cast<const spcies_consts *>(g_blob);
    const BandTables *T = reinterpret_cast<const BandTables *>(smem_raw);
    stage_constants(smem_raw, g_blob + BAND_OFFSET, (uint32_t)BAND_BYTES, &mbar);

    const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3, warp = threadIdx.x >> 5;
    const int cc[2] = {2 * t4, 2 * t4 + 1};
    const int xe[2] = {T->xat[cc[0]], T->xat[cc[1]]}, ue[2] = {T->uat[cc[0]], T->uat[cc[1]]};
    const bool xs[2] = {xe[0] >= 0, xe[1] >= 0}, us[2] = {ue[0] >= 0, ue[1] >= 0};
    const unsigned gmask = 0xFu << (4 * g);
    const bool leader = t4 == 0, lo2 = t4 < 2;
    constexpr int S0_COL = L::col_u(0);                      // s_0 / mu_0 live in the first input column of the cone block (an odd column)
    const bool own0 = cc[1] == S0_COL;
    // z_b at st[2 b], lambda_b at st[2 b + 1] (b < N), then z_N, s, mu
    double2 *st = reinterpret_cast<double2 *>(smem_raw + BAND_BYTES + warp * BAND_STATE_PER_WARP) + lane;
    constexpr int BLK_ZN = 2 * N, BLK_S = 2 * N + 1, BLK_MU = 2 * N + 2;
    auto LDZ = [&](int b) { return st[(2 * b) * 32]; };
    auto LDL = [&](int b) { return st[(2 * b + 1) * 32]; };
    auto FRAG = [&](const double *M) { return reinterpret_cast<const double2 *>(M)[lane]; };

    const double2 nab = FRAG(T->NAB), abt = FRAG(T->ABt);
    const double sigma_ = C->sigma, sigma_i_ = C->sigma_i, rho_ = C->rho, rho_i_ = C->rho_i;
    const double told = (double)tol_d, tolp = (double)tol_p;
    double hi[2], qs[2], maskx[2], tdz[2], tpz[2], td0[2], tp0[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        hi[i] = T->Hi[cc[i]];
        qs[i] = T->Qs[cc[i]];
        maskx[i] = xs[i] ? 1.0 : 0.0;
        tdz[i] = (xs[i] || us[i]) ? told : 1e300;            // blocks 1..N-1: every column of z
        tpz[i] = (xs[i] || us[i]) ? tolp : 1e300;
        td0[i] = us[i] ? told : 1e300;                       // block 0: the input columns only
        tp0[i] = us[i] ? tolp : 1e300;
    }

    const WorkQueue wq{io.queue, io.B, io.ready};
    wq.mark_start();
    unsigned long long stat_k = 0;
    unsigned int stat_nc = 0;
    long long inst = -1;
    int k = 0;
    bool live = false, drained = false;
    double q[2] = {0, 0}, nx0[2] = {0, 0}, qT[2] = {0, 0}, pxr[2] = {0, 0}, r_ = 0.0;

    for (;;) {
        // ---- refill                                                             code_ellipMPC_ADMM_soc_C.c:84-131
        const bool need = !live && !drained;
        if (__any_sync(FULL, need)) {
            long long slot = -1;
            if (need && leader) slot = wq.next();
            slot = __shfl_sync(FULL, slot, lane & ~3);
            if (need) {
                if (slot < 0) {
                    drained = true;
                    if (leader) wq.mark_drained();
                } else {
                    inst = slot;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double xr_ = xs[i] ? eng_x(C, io.xr, inst, n, xe[i]) : 0.0;
                        const double ur_ = us[i] ? eng_u(C, io.ur, inst, m, ue[i]) : 0.0;
                        q[i] = qs[i] * (xs[i] ? xr_ : ur_);                      // Q, R, T stored negated      :102-131
                        nx0[i] = xs[i] ? -eng_x(C, io.x0, inst, n, xe[i]) : 0.0;
                        double qt = 0.0, px = 0.0;
                        if (xs[i])
                            for (int j = 0; j < n; ++j) {
                                const double xj = eng_x(C, io.xr, inst, n, j);
                                qt = fma(T->Tm[xe[i]][j], xj, qt);
                                px = fma(T->PhiP[xe[i]][j], xj, px);                 // bh of the cone rows = -PhiP xr   :96-101
                            }
                        qT[i] = qt;
                        pxr[i] = px;
                    }
                    r_ = io.r[inst];
#pragma unroll 4
                    for (int e = 0; e < BAND_NST; ++e) st[e * 32] = make_double2(0.0, 0.0);
                    k = 0;
                    live = true;
                }
            }
            __syncwarp();
        }
        if (!__any_sync(FULL, live)) break;

        // ================= pass A: q_hat -> s -> r.h.s. -> forward recurrence =================
        double mup[NBK][2];   // nu'_b (with the second copies of x_4.. in register 1 of lanes 2,3)
        double zNh[2] = {0.0, 0.0};   // q_hat_N = qT - sigma z_N
        {
            double sp[2];
            {
                const double2 v = LDZ(0), la = LDL(0);
                const double qh0 = fma(-sigma_, v.x, q[0] + la.x), qh1 = fma(-sigma_, v.y, q[1] + la.y);
                sp[0] = xs[0] ? nx0[0] : hi[0] * qh0;           // block 0: the state columns carry -x0 (bh_0 = -A x0, :84-90)
                sp[1] = xs[1] ? nx0[1] : hi[1] * qh1;
            }
            constexpr int GRP = 4;   // blocks per group: their r.h.s. products are independent and issued back to back
#pragma unroll
            for (int b0 = 0; b0 < NBK; b0 += GRP) {
                double s[GRP + 1][2], r[GRP][2], e[GRP][2], cN[2] = {0.0, 0.0};
                s[0][0] = sp[0];
                s[0][1] = sp[1];
#pragma unroll
                for (int j = 0; j < GRP; ++j) {
                    const int b = b0 + j;
                    if (b >= NBK) break;
                    if (b + 1 < N) {
                        const double2 v = LDZ(b + 1), la = LDL(b + 1);
                        s[j + 1][0] = hi[0] * fma(-sigma_, v.x, q[0] + la.x);
                        s[j + 1][1] = hi[1] * fma(-sigma_, v.y, q[1] + la.y);
                    } else if (b + 1 == N) {
                        const double2 v = st[BLK_ZN * 32];
                        zNh[0] = xs[0] ? fma(-sigma_, v.x, qT[0]) : 0.0;
                        zNh[1] = xs[1] ? fma(-sigma_, v.y, qT[1]) : 0.0;
                        mma::mv(s[j + 1], FRAG(T->HiN), zNh, 0.0, 0.0);
                    } else {   // b == N: the cone rows, r_N = P_half s_N - (mu_s - rho s_s) / rho + PhiP xr
                        const double2 ss = st[BLK_S * 32], ms = st[BLK_MU * 32];
                        cN[0] = xs[0] ? fma(-rho_i_, ms.x, ss.x) + pxr[0] : 0.0;
                        cN[1] = xs[1] ? fma(-rho_i_, ms.y, ss.y) + pxr[1] : 0.0;
                        s[j + 1][0] = s[j + 1][1] = 0.0;
                    }
                }
                const double2 ph = FRAG(T->Ph);
#pragma unroll
                for (int j = 0; j < GRP; ++j) {
                    const int b = b0 + j;
                    if (b < N) dmma(e[j][0], e[j][1], s[j][0], nab.x, s[j + 1][0], s[j + 1][1]);
                    else if (b == N) dmma(e[j][0], e[j][1], s[j][0], ph.x, cN[0], cN[1]);
                }
#pragma unroll
                for (int j = 0; j < GRP; ++j) {
                    const int b = b0 + j;
                    if (b < N) dmma(r[j][0], r[j][1], s[j][1], nab.y, e[j][0], e[j][1]);
                    else if (b == N) dmma(r[j][0], r[j][1], s[j][1], ph.y, e[j][0], e[j][1]);
                }
#pragma unroll
                for (int j = 0; j < GRP; ++j)
                    if (b0 + j < NBK) dmma(e[j][0], e[j][1], r[j][0], T->FWa[b0 + j][lane].x, 0.0, 0.0);
#pragma unroll
                for (int j = 0; j < GRP; ++j) {
                    const int b = b0 + j;
                    if (b >= NBK) break;
                    if (b == 0) {
                        dmma(mup[0][0], mup[0][1], r[j][1], T->FWb[0][lane], e[j][0], e[j][1]);
                    } else {
                        double f0, f1;
                        dmma(f0, f1, mup[b - 1][0], T->FWa[b][lane].y, e[j][0], e[j][1]);
                        dmma(mup[b][0], mup[b][1], lo2 ? r[j][1] : mup[b - 1][1], T->FWb[b][lane], f0, f1);
                    }
                }
                constexpr int LASTJ = GRP;
                sp[0] = s[(b0 + GRP <= NBK) ? LASTJ : (NBK - b0)][0];
                sp[1] = s[(b0 + GRP <= NBK) ? LASTJ : (NBK - b0)][1];
            }
        }

        // ================= pass B: backward recurrence -> primal_hat -> primal -> dual -> residuals =================
        bool over = false;
        double mu[2], mun[2];   // nu_b, nu_{b-1}
        {
            double g0, g1;
            dmma(g0, g1, mup[N][0], T->BWa[N][lane].x, 0.0, 0.0);
            dmma(mu[0], mu[1], mup[N][1], T->BWb[N][lane], g0, g1);               // nu_N: the multipliers of the cone rows
        }
        {   // s = proj_SOC(s_hat + mu / rho), mu += rho (s_hat - s)                                       :220-254
            const double2 ss = st[BLK_S * 32], ms = st[BLK_MU * 32];
            const double sh[2] = {xs[0] ? fma(-rho_i_, ms.x + mu[0], ss.x) : 0.0, xs[1] ? fma(-rho_i_, ms.y + mu[1], ss.y) : 0.0};
            double w[2] = {xs[0] ? fma(rho_i_, ms.x, sh[0]) : 0.0, xs[1] ? fma(rho_i_, ms.y, sh[1]) : 0.0};
            const double w0 = fma(rho_i_, ms.y, r_);                              // s_hat_0 = r (own0 lanes)
            double part = fma(w[0], w[0], w[1] * w[1]);
            part += __shfl_xor_sync(FULL, part, 1);
            part += __shfl_xor_sync(FULL, part, 2);
            const double nrm = sqrt(part);
            const double x0c = __shfl_sync(FULL, w0, (lane & ~3) + S0_COL / 2);
            double s0n = x0c;
            if (nrm <= x0c) {
            } else if (nrm <= -x0c) {
                w[0] = w[1] = 0.0;
                s0n = 0.0;
            } else {
                const double step = (x0c + nrm) / (2.0 * nrm);
                s0n = step * nrm;
                w[0] *= step;
                w[1] *= step;
            }
            over = over || (xs[0] && ((fabs(ss.x - w[0]) > told) || (fabs(w[0] - sh[0]) > tolp))) ||
                   (xs[1] && ((fabs(ss.y - w[1]) > told) || (fabs(w[1] - sh[1]) > tolp))) ||
                   (own0 && ((fabs(ss.y - s0n) > told) || (fabs(s0n - r_) > tolp)));
            double2 sn, mn;
            sn.x = w[0];
            mn.x = xs[0] ? fma(rho_, sh[0] - w[0], ms.x) : 0.0;
            sn.y = own0 ? s0n : w[1];
            mn.y = own0 ? fma(rho_, r_ - s0n, ms.y) : (xs[1] ? fma(rho_, sh[1] - w[1], ms.y) : 0.0);
            st[BLK_S * 32] = sn;
            st[BLK_MU * 32] = mn;
        }
        {   // nu_{N-1}, then z_N = -Hi_N (q_hat_N - nu_{N-1} - P_half' nu_N)
            double g0, g1, h0, h1;
            dmma(g0, g1, mup[N - 1][0], T->BWa[N - 1][lane].x, 0.0, 0.0);
            dmma(h0, h1, mu[0], T->BWa[N - 1][lane].y, g0, g1);
            dmma(mun[0], mun[1], lo2 ? mup[N - 1][1] : mu[1], T->BWb[N - 1][lane], h0, h1);
            const double aux[2] = {fma(-maskx[0], mun[0], zNh[0]), fma(-maskx[1], mun[1], zNh[1])};
            double t_[2], zn[2];
            mma::mv(t_, FRAG(T->NPht), mu, aux[0], aux[1]);
            mma::mv(zn, FRAG(T->NHiN), t_, 0.0, 0.0);
            const double2 v = st[BLK_ZN * 32];
            over = over || (xs[0] && fabs(v.x - zn[0]) > told) || (xs[1] && fabs(v.y - zn[1]) > told);
            st[BLK_ZN * 32] = make_double2(xs[0] ? zn[0] : 0.0, xs[1] ? zn[1] : 0.0);
            mu[0] = mun[0];
            mu[1] = mun[1];
        }
        double u0v[2] = {0, 0};
#pragma unroll
        for (int b = N - 1; b >= 0; --b) {
            if (b > 0) {   // nu_{b-1} = Uinv nu'_{b-1} - G nu_b
                double g0, g1, h0, h1;
                dmma(g0, g1, mup[b - 1][0], T->BWa[b - 1][lane].x, 0.0, 0.0);
                dmma(h0, h1, mu[0], T->BWa[b - 1][lane].y, g0, g1);
                dmma(mun[0], mun[1], lo2 ? mup[b - 1][1] : mu[1], T->BWb[b - 1][lane], h0, h1);
            }
            const double2 v = LDZ(b), la = LDL(b);
            double c0 = fma(-sigma_, v.x, q[0] + la.x), c1 = fma(-sigma_, v.y, q[1] + la.y);      // q_hat_b
            if (b > 0) {
                c0 = fma(-maskx[0], mun[0], c0);                                                  // - [nu_{b-1}; 0]
                c1 = fma(-maskx[1], mun[1], c1);
            }
            double a[2];
            mma::mv(a, abt, mu, c0, c1);                                                          // + [A B]' nu_b
            const double2 lo = reinterpret_cast<const double2 *>(T->LB[b])[t4], up = reinterpret_cast<const double2 *>(T->UB[b])[t4];
            double2 vn, ln;
            {
                const double z0 = -hi[0] * a[0], z1 = -hi[1] * a[1];                              // z_hat_b
                vn.x = clip(fma(sigma_i_, la.x, z0), lo.x, up.x);
                vn.y = clip(fma(sigma_i_, la.y, z1), lo.y, up.y);
                const double d0 = z0 - vn.x, d1 = z1 - vn.y;
                const double e0 = b > 0 ? tdz[0] : td0[0], e1 = b > 0 ? tdz[1] : td0[1];
                const double p0 = b > 0 ? tpz[0] : tp0[0], p1 = b > 0 ? tpz[1] : tp0[1];
                over = over || (fabs(v.x - vn.x) > e0) || (fabs(d0) > p0) || (fabs(v.y - vn.y) > e1) || (fabs(d1) > p1);
                ln.x = fma(sigma_, d0, la.x);
                ln.y = fma(sigma_, d1, la.y);
            }
            if (b == 0) {   // block 0 only has input columns: keep the others at zero
                vn.x = us[0] ? vn.x : 0.0;
                vn.y = us[1] ? vn.y : 0.0;
                ln.x = us[0] ? ln.x : 0.0;
                ln.y = us[1] ? ln.y : 0.0;
                u0v[0] = vn.x;
                u0v[1] = vn.y;
            }
            st[(2 * b) * 32] = vn;
            st[(2 * b + 1) * 32] = ln;
            mu[0] = mun[0];
            mu[1] = mun[1];
        }

        // ================= exit condition                                            :260-283 =================
        if (live) k += 1;
        const bool gover = (__ballot_sync(FULL, over) & gmask) != 0u;
        if (live) {
            const int ef = !gover ? 1 : ((k >= k_max) ? -1 : 0);
            if (ef != 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (us[i]) io.u[inst * m + ue[i]] = eng_u_out(C, u0v[i], ue[i]);              // u_opt = z[0..m)   (:297-306)
                if (leader) {
                    io.k[inst] = k;
                    io.e[inst] = ef;
                    stat_k += (unsigned long long)k;
                    stat_nc += (ef < 0);
                }
                live = false;
            }
        }
    }
    flush_stats(io.queue, stat_k, stat_nc);
    wq.mark_end();
}
#endif
