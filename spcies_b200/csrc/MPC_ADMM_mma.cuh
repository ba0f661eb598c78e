// MPC_ADMM_mma.cuh -- tensor-core (DMMA) engine of the equMPC ADMM solver (included by MPC_ADMM.cuh, inside spcies::admm).
//
// Same mapping as the FISTA engine (MPC_FISTA_mma.cuh, spcies_mma.cuh): 8 instances per warp, 4 lanes per instance, 2 columns
// per lane, the shared block matrices ([A B], the explicit inverses of the banded Cholesky factor) as FP64 MMA B fragments.
// At N = 20 (BASELINE.json configs[2]) the iterates do not fit the register file (v, lambda: 2 N vectors), and with one thread
// per instance not even shared memory -- the scalar kernel runs that configuration from a global scratch.  Here they live in
// shared memory as [block][lane] double2 (one conflict-free 128-bit access per lane, block and pass; 20 KB per warp), the
// forward-substituted mu' stays in registers between the two passes.
//
// One ADMM iteration (code_equMPC_ADMM_C.c:291-553), blocks b = 0..N-1 (block 0 = the first m decision variables u_0, block
// l+1 = stage l of the reference's z / v / lambda arrays):
//   pass A  q_hat_b = q + lambda_b - rho v_b,  s_b = Hi o q_hat_b   (s_0: state columns = -x0, s_N := -xr)
//           r_b = s_{b+1} - [A B] s_b                                        2 MMA   (:332-353)
//           mu'_b = Linv_b r_b - F_b mu'_{b-1}                               3 MMA   (:356-381, merged half k-steps)
//   pass B  mu_b = Uinv_b mu'_b - G_b mu_{b+1}                               3 MMA   (:385-422)
//           z_b = -Hi o (q_hat_b - [mu_{b-1}; 0] + [A B]' mu_b)              2 MMA   (:424-440)
//           v_b = clip(z_b + lambda_b / rho),  lambda_b += rho (z_b - v_b),  exit test |v_old - v|, |z - v| > tol   (:447-524)
// 10 N - 2 = 198 DMMA and ~600 FP64-pipe instructions per warp-iteration at N = 20.
//
// Arithmetic: FAST (FMA, explicit block inverses, dot products in the MMA's order); EXACT mode, the debug payload, float,
// per-stage penalty / bounds (rho arrays, VAR_BOUNDS) and the laxMPC / ellipMPC terminal blocks use the scalar kernel.
#pragma once
// (spcies_mma.cuh is included by the parent header, outside its namespace)

#ifndef SPCIES_ADMM_MMA
#define SPCIES_ADMM_MMA 1
#endif
#ifndef SPCIES_ADMM_MMA_BLOCK
#define SPCIES_ADMM_MMA_BLOCK 0          // 0: as many warps as the iterates leave room for, 8 .. 12 (below)
#endif

#if defined(SCALAR_RHO) && !defined(VAR_BOUNDS)
#define SPCIES_ADMM_MMA_ELIGIBLE 1
#else
#define SPCIES_ADMM_MMA_ELIGIBLE 0
#endif

constexpr bool MMA_SHAPE_OK = n >= 5 && n <= 6 && nm <= 8 && N >= 3;
constexpr int MMA_NBLK = N + (HAS_TN ? 1 : 0);   // blocks of v / lambda: u_0, stages 0..N-2 [, the terminal state of laxMPC]
constexpr size_t MMA_STATE_PER_WARP = (size_t)2 * MMA_NBLK * 32 * sizeof(double2);

#if SPCIES_ADMM_MMA_ELIGIBLE
struct alignas(16) MmaTables {
    // by column (spcies_mma.cuh: MmaLayout); matrices row-major [output column][input column]
    double NAB[64];           // -[A B]   (out x, in z)
    double NAB0[64];          // the same for block 0, whose state columns carry -x0 (see fill_mma_tables: float constants)
    double ABt[64];           //  [A B]'  (out z, in x)
    double2 FWa[N][32], BWa[N][32];
    double FWb[N][32], BWb[N][32];
    double Hi[8];             // inverse of the (diagonal) Hessian + rho, by column of z; the same for every block (checked on the host)
    double HiN[64], NHiN[64]; // laxMPC: the dense terminal block Hi_N and its negative, [output column][input column] (x columns)
    double Tm[8][8];          // laxMPC / ellipMPC: T by component (qT = T xr, computed per instance at refill)
    // ellipMPC (terminal ellipsoid (v - c)' P (v - c) <= r^2 in the P^(1/2) metric): P_half, -rho P, rho_i Pinv_half, P by column
    double Ph[64], NRP[64], PihR[64], Pm[64];
    double cvec[8], rr, rv;   // centre by column, r^2, r
    double Qs[8];             // q = Qs o [xr; ur]   (Q, R stored negated)
    double LB[8], UB[8];
    int xat[8], uat[8];
};
constexpr size_t MMA_BYTES = (sizeof(MmaTables) + 15) / 16 * 16;
constexpr size_t CONSTS_BYTES_ = (sizeof(spcies_consts) + 15) / 16 * 16;
constexpr size_t MMA_OFFSET = CONSTS_BYTES_;
// Warps per SM: the kernel needs 168 registers (12 warps) and 2 N tiles of iterates per warp; the recurrences are latency bound,
// so every warp that fits helps (C3, N = 20: 9 warps 4.23 M solves/s against 4.16 M with 8; N = 10: 12 warps)
constexpr int MMA_FIT = MMA_BYTES + 64 >= 227 * 1024 ? 0 : (int)((227 * 1024 - 64 - MMA_BYTES) / MMA_STATE_PER_WARP);
constexpr int MMA_WARPS = MMA_FIT >= 12 ? 12 : (MMA_FIT >= 8 ? MMA_FIT : 8);
constexpr int MMA_BLOCK = SPCIES_ADMM_MMA_BLOCK > 0 ? SPCIES_ADMM_MMA_BLOCK : 32 * MMA_WARPS;
constexpr int MMA_IPB = MMA_BLOCK / 4;
constexpr size_t MMA_SMEM = MMA_BYTES + (MMA_BLOCK / 32) * MMA_STATE_PER_WARP;
constexpr bool HAS_MMA = SPCIES_ADMM_MMA != 0 && MMA_SHAPE_OK && sizeof(SPCIES_REAL) == 8 && MMA_SMEM <= 227 * 1024 - 64;

// false when the generated constants are not uniform over the horizon (then the engine is not used)
static inline bool fill_mma_tables(const spcies_consts &C, MmaTables &T) {
    typedef mma::MmaLayout<n, m> L;
    memset(&T, 0, sizeof T);
    for (int l = 0; l < N - 1; ++l)
        for (int j = 0; j < nm; ++j)
            if (C.Hi[l][j] != C.Hi[0][j]) return false;
    for (int j = 0; j < m; ++j)
        if (C.Hi_0[j] != C.Hi[0][n + j]) return false;
#if SPCIES_TERMINAL == 2
    for (int l = 0; l < N - 1; ++l)          // the engine keeps one set of bounds: they must not depend on the stage
        for (int j = 0; j < nm; ++j)
            if (C.LBz[l][j] != C.LBz[0][j] || C.UBz[l][j] != C.UBz[0][j]) return false;
    for (int j = 0; j < m; ++j)
        if (C.LBu0[j] != C.LBz[0][n + j] || C.UBu0[j] != C.UBz[0][n + j]) return false;
#endif
    for (int c = 0; c < 8; ++c) {
        T.xat[c] = L::x_at(c);
        T.uat[c] = L::u_at(c);
        const int z = L::z_at(c);
        T.LB[c] = -1e300;
        T.UB[c] = 1e300;
        if (z < 0) continue;
        T.Hi[c] = (double)C.Hi[0][z];
        T.Qs[c] = z < n ? (double)C.Q[z] : (double)C.R[z - n];
#if SPCIES_TERMINAL == 2
        T.LB[c] = (double)C.LBz[0][z];
        T.UB[c] = (double)C.UBz[0][z];
#else
        T.LB[c] = (double)C.LB[z];
        T.UB[c] = (double)C.UB[z];
#endif
    }
    for (int oc = 0; oc < 8; ++oc)
        for (int ic = 0; ic < 8; ++ic) {
            if (L::x_at(oc) >= 0 && L::z_at(ic) >= 0) {
                // r_b = s_{b+1} - [A B] s_b with s_b = Hi o q_hat_b.  The templates evaluate `AB[j][i]*Hi[l-1][i]*z[l-1][i]`
                // (code_equMPC_ADMM_C.c:340): with float constants the first product is rounded to float, so the fragment
                // applied to s_b is that product divided by Hi again; b = -A x0 (:272) takes the plain entries (block 0).
                const int x = L::x_at(oc), z = L::z_at(ic);
                T.NAB0[oc * 8 + ic] = T.NAB[oc * 8 + ic] = -(double)C.AB[x][z];
                if (sizeof(C.AB[0][0]) == 4) {
                    const double eff = -cprod_host(C.AB[x][z], C.Hi[0][z]) / (double)C.Hi[0][z];
                    T.NAB[oc * 8 + ic] = eff;
                    if (z >= n) T.NAB0[oc * 8 + ic] = eff;
                }
            }
            if (L::z_at(oc) >= 0 && L::x_at(ic) >= 0) T.ABt[oc * 8 + ic] = (double)C.AB[L::x_at(ic)][L::z_at(oc)];
        }
#if SPCIES_TERMINAL != 0
    for (int oc = 0; oc < 8; ++oc)
        for (int ic = 0; ic < 8; ++ic)
            if (L::x_at(oc) >= 0 && L::x_at(ic) >= 0) {
                const int i = L::x_at(oc), j = L::x_at(ic);
                T.HiN[oc * 8 + ic] = (double)C.Hi_N[i][j];
                T.NHiN[oc * 8 + ic] = -(double)C.Hi_N[i][j];
#if SPCIES_TERMINAL == 2
                T.Ph[oc * 8 + ic] = (double)C.P_half[i][j];
                T.NRP[oc * 8 + ic] = -(double)C.P[i][j] * (double)rho;
                T.PihR[oc * 8 + ic] = (double)C.Pinv_half[i][j] * (double)rho_i;
                T.Pm[oc * 8 + ic] = (double)C.P[i][j];
#endif
            }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) T.Tm[i][j] = (double)C.T[i][j];
#if SPCIES_TERMINAL == 2
    for (int c = 0; c < 8; ++c) T.cvec[c] = L::x_at(c) >= 0 ? (double)C.c[L::x_at(c)] : 0.0;
    T.rr = (double)C.r * (double)C.r;
    T.rv = (double)C.r;
#endif
#endif
    typedef double Blk[n][n];
    Blk *Linv = new Blk[4 * N], *F = Linv + N, *Uinv = F + N, *G = Uinv + N;
    mma::block_inverses<N, n>(C, Linv, F, Uinv, G);
    mma::recurrence_fragments<N, n, m, double>(Linv, F, Uinv, G, T.FWa, T.FWb, T.BWa, T.BWb);
    delete[] Linv;
    return true;
}

template <bool VARB>
__global__ void __launch_bounds__(MMA_BLOCK, 1) admm_mma_kernel(const BatchIO io, const unsigned char *__restrict__ g_blob) {
    using mma::dmma;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    const spcies_consts *C = reinterpret_cast<const spcies_consts *>(g_blob);   // engineering-unit scaling only
    (void)C;
    const MmaTables *T = reinterpret_cast<const MmaTables *>(smem_raw);
    stage_constants(smem_raw, g_blob + MMA_OFFSET, (uint32_t)MMA_BYTES, &mbar);

    const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3, warp = threadIdx.x >> 5;
    const int cc[2] = {2 * t4, 2 * t4 + 1};
    const int xe[2] = {T->xat[cc[0]], T->xat[cc[1]]}, ue[2] = {T->uat[cc[0]], T->uat[cc[1]]};
    const bool xs[2] = {xe[0] >= 0, xe[1] >= 0}, us[2] = {ue[0] >= 0, ue[1] >= 0};
    const unsigned gmask = 0xFu << (4 * g);
    const bool leader = t4 == 0, lo2 = t4 < 2;
    // v_b at st[2 b], lambda_b at st[2 b + 1]
    double2 *st = reinterpret_cast<double2 *>(smem_raw + MMA_BYTES + warp * MMA_STATE_PER_WARP) + lane;
    auto LDV = [&](int b) { return st[(2 * b) * 32]; };
    auto LDL = [&](int b) { return st[(2 * b + 1) * 32]; };

    const double2 nab = reinterpret_cast<const double2 *>(T->NAB)[lane], abt = reinterpret_cast<const double2 *>(T->ABt)[lane];
    const double rho_ = (double)rho, rhoi_ = (double)rho_i;
    double hi[2], qs[2], lb[2], ub[2], maskx[2], tolx[2], tol0[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        hi[i] = T->Hi[cc[i]];
        qs[i] = T->Qs[cc[i]];
        lb[i] = T->LB[cc[i]];
        ub[i] = T->UB[cc[i]];
        maskx[i] = xs[i] ? 1.0 : 0.0;
        tolx[i] = (xs[i] || us[i]) ? (double)tol : 1e300;     // blocks 1..N-1: every column of z
        tol0[i] = us[i] ? (double)tol : 1e300;                // block 0: the input columns only
    }

    const WorkQueue wq{io.queue, io.B, io.ready};
    wq.mark_start();
    unsigned long long stat_k = 0;
    unsigned int stat_nc = 0;
    long long inst = -1;
    int k = 0;
    bool live = false, drained = false;
    double q[2] = {0, 0}, nx0[2] = {0, 0}, nxr[2] = {0, 0};   // laxMPC: nxr holds qT = T xr instead of -xr
#if SPCIES_TERMINAL != 0
    const double2 hin = reinterpret_cast<const double2 *>(T->HiN)[lane], nhin = reinterpret_cast<const double2 *>(T->NHiN)[lane];
    const double tolxs[2] = {xs[0] ? (double)tol : 1e300, xs[1] ? (double)tol : 1e300};
#endif
#if SPCIES_TERMINAL == 2
    const double2 ph = reinterpret_cast<const double2 *>(T->Ph)[lane], nrp = reinterpret_cast<const double2 *>(T->NRP)[lane];
    const double2 pihr = reinterpret_cast<const double2 *>(T->PihR)[lane], pm = reinterpret_cast<const double2 *>(T->Pm)[lane];
    const double cv[2] = {T->cvec[cc[0]], T->cvec[cc[1]]}, r_ = T->rv, rr_ = T->rr;
#endif

    for (;;) {
        // ---- refill: a lane group without an instance pulls the next one          code_equMPC_ADMM_C.c:268-283
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
                        q[i] = qs[i] * (xs[i] ? xr_ : ur_);
                        nx0[i] = xs[i] ? -eng_x(C, io.x0, inst, n, xe[i]) : 0.0;
#if SPCIES_TERMINAL != 0
                        double qt = 0.0;                                   // qT = T xr (T dense, negated)   code_laxMPC_ADMM_C.c:292-295
                        if (xs[i])
                            for (int j = 0; j < n; ++j) qt = fma(T->Tm[xe[i]][j], eng_x(C, io.xr, inst, n, j), qt);
                        nxr[i] = qt;
#else
                        nxr[i] = -xr_;
#endif
                        if (VARB) {
                            const int ze = xs[i] ? xe[i] : n + ue[i];
                            lb[i] = (xs[i] || us[i]) ? io.LB[inst * nm + ze] : -1e300;
                            ub[i] = (xs[i] || us[i]) ? io.UB[inst * nm + ze] : 1e300;
                        }
                    }
#pragma unroll 4
                    for (int e = 0; e < 2 * MMA_NBLK; ++e) st[e * 32] = make_double2(0.0, 0.0);
                    k = 0;
                    live = true;
                }
            }
            __syncwarp();
        }
        if (!__any_sync(FULL, live)) break;

        // ================= pass A: q_hat -> s -> r.h.s. -> forward recurrence =================
        double mup[N][2];   // mu'_b (with the second copies of x_4.. in register 1 of lanes 2,3)
#if SPCIES_TERMINAL != 0
        double zNh[2] = {0.0, 0.0};   // z_N_hat = qT + lambda_N - rho v_N  (ellipMPC: in the P^(1/2) metric)
#endif
        {
            double sp[2];   // s_{b}
            {
                const double2 v = LDV(0), la = LDL(0);
                const double qh0 = fma(-rho_, v.x, q[0] + la.x), qh1 = fma(-rho_, v.y, q[1] + la.y);
                sp[0] = xs[0] ? nx0[0] : hi[0] * qh0;           // block 0: the state columns carry -x0 (r_0 = ... - b, b = -A x0)
                sp[1] = xs[1] ? nx0[1] : hi[1] * qh1;
            }
            const double2 nab0 = reinterpret_cast<const double2 *>(T->NAB0)[lane];
            constexpr int GRP = 5;   // blocks per group: their r.h.s. products are independent and issued back to back
#pragma unroll
            for (int b0 = 0; b0 < N; b0 += GRP) {
                double s[GRP + 1][2], r[GRP][2], e[GRP][2];
                s[0][0] = sp[0];
                s[0][1] = sp[1];
#pragma unroll
                for (int j = 0; j < GRP; ++j) {
                    const int b = b0 + j;               // r_b needs s_{b+1}
                    if (b >= N) break;
                    if (b + 1 < N) {
                        const double2 v = LDV(b + 1), la = LDL(b + 1);
                        s[j + 1][0] = hi[0] * fma(-rho_, v.x, q[0] + la.x);
                        s[j + 1][1] = hi[1] * fma(-rho_, v.y, q[1] + la.y);
                    } else {
#if SPCIES_TERMINAL == 1
                        // laxMPC: s_N = Hi_N (qT + lambda_N - rho v_N)                      code_laxMPC_ADMM_C.c:373-381
                        const double2 v = LDV(N), la = LDL(N);
                        zNh[0] = xs[0] ? fma(-rho_, v.x, nxr[0] + la.x) : 0.0;
                        zNh[1] = xs[1] ? fma(-rho_, v.y, nxr[1] + la.y) : 0.0;
                        mma::mv(s[j + 1], hin, zNh, 0.0, 0.0);
#elif SPCIES_TERMINAL == 2
                        // ellipMPC: z_N_hat = qT + P_half lambda_N - rho P v_N;  s_N = Hi_N z_N_hat   code_ellipMPC_ADMM_C.c:146-155
                        const double2 v = LDV(N), la = LDL(N);
                        const double lav[2] = {la.x, la.y}, vv[2] = {v.x, v.y};
                        double t_[2];
                        mma::mv(t_, ph, lav, nxr[0], nxr[1]);
                        mma::mv(zNh, nrp, vv, t_[0], t_[1]);
                        mma::mv(s[j + 1], hin, zNh, 0.0, 0.0);
#else
                        s[j + 1][0] = nxr[0];           // :351-353
                        s[j + 1][1] = nxr[1];
#endif
                    }
                }
#pragma unroll
                for (int j = 0; j < GRP; ++j)
                    if (b0 + j < N) dmma(e[j][0], e[j][1], s[j][0], b0 + j == 0 ? nab0.x : nab.x, s[j + 1][0], s[j + 1][1]);
#pragma unroll
                for (int j = 0; j < GRP; ++j)
                    if (b0 + j < N) dmma(r[j][0], r[j][1], s[j][1], b0 + j == 0 ? nab0.y : nab.y, e[j][0], e[j][1]);
#pragma unroll
                for (int j = 0; j < GRP; ++j)
                    if (b0 + j < N) dmma(e[j][0], e[j][1], r[j][0], T->FWa[b0 + j][lane].x, 0.0, 0.0);
#pragma unroll
                for (int j = 0; j < GRP; ++j) {
                    const int b = b0 + j;
                    if (b >= N) break;
                    if (b == 0) {
                        dmma(mup[0][0], mup[0][1], r[j][1], T->FWb[0][lane], e[j][0], e[j][1]);
                    } else {
                        double f0, f1;
                        dmma(f0, f1, mup[b - 1][0], T->FWa[b][lane].y, e[j][0], e[j][1]);
                        dmma(mup[b][0], mup[b][1], lo2 ? r[j][1] : mup[b - 1][1], T->FWb[b][lane], f0, f1);
                    }
                }
                constexpr int LASTJ = GRP;
                sp[0] = s[(b0 + GRP <= N) ? LASTJ : (N - b0)][0];
                sp[1] = s[(b0 + GRP <= N) ? LASTJ : (N - b0)][1];
            }
        }

        // ================= pass B: backward recurrence -> z -> v -> lambda -> residuals =================
        bool over = false;
        double mu[2], mun[2];   // mu_b, mu_{b-1}
        {
            double g0, g1;
            dmma(g0, g1, mup[N - 1][0], T->BWa[N - 1][lane].x, 0.0, 0.0);
            dmma(mu[0], mu[1], mup[N - 1][1], T->BWb[N - 1][lane], g0, g1);
        }
#if SPCIES_TERMINAL == 1
        {   // terminal block: z_N = -Hi_N (z_N_hat - mu_{N-1}), v_N, lambda_N           code_laxMPC_ADMM_C.c:476-485, :523-538, :561-569
            const double aux[2] = {fma(-maskx[0], mu[0], zNh[0]), fma(-maskx[1], mu[1], zNh[1])};
            double zn[2];
            mma::mv(zn, nhin, aux, 0.0, 0.0);
            const double2 v = LDV(N), la = LDL(N);
            double2 vn, ln;
            vn.x = xs[0] ? clip(fma(rhoi_, la.x, zn[0]), lb[0], ub[0]) : 0.0;
            vn.y = xs[1] ? clip(fma(rhoi_, la.y, zn[1]), lb[1], ub[1]) : 0.0;
            const double d0 = zn[0] - vn.x, d1 = zn[1] - vn.y;
            over = over || (fabs(v.x - vn.x) > tolxs[0]) || (fabs(d0) > tolxs[0]) || (fabs(v.y - vn.y) > tolxs[1]) || (fabs(d1) > tolxs[1]);
            ln.x = xs[0] ? fma(rho_, d0, la.x) : 0.0;
            ln.y = xs[1] ? fma(rho_, d1, la.y) : 0.0;
            st[(2 * N) * 32] = vn;
            st[(2 * N + 1) * 32] = ln;
        }
#elif SPCIES_TERMINAL == 2
        {   // terminal block with the radial projection onto the ellipsoid in the P metric   code_ellipMPC_ADMM_C.c:319-351, :375-387
            const double aux[2] = {fma(-maskx[0], mu[0], zNh[0]), fma(-maskx[1], mu[1], zNh[1])};
            double zn[2], vn[2], a2[2], ln[2];
            mma::mv(zn, nhin, aux, 0.0, 0.0);
            const double2 v = LDV(N), la = LDL(N);
            const double lav[2] = {la.x, la.y};
            mma::mv(vn, pihr, lav, zn[0], zn[1]);                                  // z_N + Pinv_half (lambda_N / rho)
            double dv[2] = {vn[0] - cv[0], vn[1] - cv[1]};
            mma::mv(a2, pm, dv, 0.0, 0.0);                                         // P (v - c)
            double vPv = (xs[0] ? dv[0] * a2[0] : 0.0) + (xs[1] ? dv[1] * a2[1] : 0.0);
            vPv += __shfl_xor_sync(FULL, vPv, 1);
            vPv += __shfl_xor_sync(FULL, vPv, 2);
            if (vPv > rr_) {
                const double sc = r_ / sqrt(vPv);
                vn[0] = fma(sc, dv[0], cv[0]);
                vn[1] = fma(sc, dv[1], cv[1]);
            }
            vn[0] = xs[0] ? vn[0] : 0.0;
            vn[1] = xs[1] ? vn[1] : 0.0;
            const double a3[2] = {xs[0] ? rho_ * (zn[0] - vn[0]) : 0.0, xs[1] ? rho_ * (zn[1] - vn[1]) : 0.0};
            mma::mv(ln, ph, a3, la.x, la.y);                                       // lambda_N + P_half rho (z_N - v_N)
            over = over || (fabs(v.x - vn[0]) > tolxs[0]) || (fabs(zn[0] - vn[0]) > tolxs[0]) || (fabs(v.y - vn[1]) > tolxs[1]) ||
                   (fabs(zn[1] - vn[1]) > tolxs[1]);
            st[(2 * N) * 32] = make_double2(vn[0], vn[1]);
            st[(2 * N + 1) * 32] = make_double2(xs[0] ? ln[0] : 0.0, xs[1] ? ln[1] : 0.0);
        }
#endif
        double u0v[2] = {0, 0};
#pragma unroll
        for (int b = N - 1; b >= 0; --b) {
            if (b > 0) {   // mu_{b-1} = Uinv mu'_{b-1} - G mu_b
                double g0, g1, h0, h1;
                dmma(g0, g1, mup[b - 1][0], T->BWa[b - 1][lane].x, 0.0, 0.0);
                dmma(h0, h1, mu[0], T->BWa[b - 1][lane].y, g0, g1);
                dmma(mun[0], mun[1], lo2 ? mup[b - 1][1] : mu[1], T->BWb[b - 1][lane], h0, h1);
            }
            const double2 v = LDV(b), la = LDL(b);
            double c0 = fma(-rho_, v.x, q[0] + la.x), c1 = fma(-rho_, v.y, q[1] + la.y);      // q_hat_b
            if (b > 0) {
                c0 = fma(-maskx[0], mun[0], c0);                                              // - [mu_{b-1}; 0]
                c1 = fma(-maskx[1], mun[1], c1);
            }
            double a[2];
            mma::mv(a, abt, mu, c0, c1);                                                      // + [A B]' mu_b
            double2 vn, ln;
            {
                const double z0 = -hi[0] * a[0], z1 = -hi[1] * a[1];
                vn.x = clip(fma(rhoi_, la.x, z0), lb[0], ub[0]);
                vn.y = clip(fma(rhoi_, la.y, z1), lb[1], ub[1]);
                const double d0 = z0 - vn.x, d1 = z1 - vn.y;
                const double t0 = b > 0 ? tolx[0] : tol0[0], t1 = b > 0 ? tolx[1] : tol0[1];
                over = over || (fabs(v.x - vn.x) > t0) || (fabs(d0) > t0) || (fabs(v.y - vn.y) > t1) || (fabs(d1) > t1);
                ln.x = fma(rho_, d0, la.x);
                ln.y = fma(rho_, d1, la.y);
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

        // ================= exit condition                                            :526-553 =================
        if (live) k += 1;
        const bool gover = (__ballot_sync(FULL, over) & gmask) != 0u;
        if (live) {
            const int ef = !gover ? 1 : ((k >= k_max) ? -1 : 0);
            if (ef != 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (us[i]) io.u[inst * m + ue[i]] = eng_u_out(C, u0v[i], ue[i]);                               // u_opt = v_0   (:557-566)
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
#else
constexpr bool HAS_MMA = false;
#endif
