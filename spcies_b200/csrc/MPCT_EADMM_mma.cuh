// MPCT_EADMM_mma.cuh -- tensor-core (DMMA) engine of the MPCT EADMM solver (included by MPCT_EADMM.cuh, inside spcies::eadmm).
//
// Same mapping as the other engines (spcies_mma.cuh): 8 instances per warp, 4 lanes per instance, 2 columns per lane.  The
// three-block EADMM iteration (code_MPCT_EADMM_C.c:85-459, IS_DIAG path) is component-wise work on nm-vectors except for
//   * rhs_l = H3i_{l+1} q3_{l+1} - [A B] (H3i_l q3_l)                2 MMA per stage   (:176-184)
//   * the banded-Cholesky solve (explicit block inverses, merged)    3 + 3 MMA         (:221-287)
//   * z3_{l+1} = -H3i (q3 - [mu_l; 0] + [A B]' mu_{l+1})             2 MMA             (:291-320)
//   * z2 = W2 q2                                                     2 MMA per iteration (:145-149)
// With one thread per instance the 12.5 KB of iterates of an N = 50 instance (BASELINE.json configs[4]) only fit a global
// scratch.  Here z3 and lambda live in shared memory as [block][lane] double2 (2 N + 4 blocks = 53 KB per warp at N = 50 -> FOUR
// warps per SM, one per scheduler; round 1 also kept z1 and mu' there: 105 KB per warp, two warps per SM, 2.9 % of the warp
// slots -- the occupancy problem the round-1 profile showed):
//   * z1 is not stored: z1_l = clip(H1i_l o (rho_l (z3_l + z2) + lambda_{l+1})) is recomputed from the *old* z3_l, lambda_{l+1}
//     and z2 where the forward and the backward sweep need it (they still hold the old values at that point: a stage's z3 and
//     lambda are overwritten at the end of its own backward step);
//   * the forward-substituted mu' is a pure stream -- written stage by stage in the forward sweep, read in reverse in the
//     backward sweep -- and goes to a per-warp global scratch (L2 resident: 25 KB per warp), read back PF stages ahead through a
//     register ring like the recurrence fragments;
//   * the per-stage component constants (rho, H1i, H3i) are one 64-byte row per stage, the recurrence fragments are streamed from
//     global memory (L2) PF stages ahead.
//
// Arithmetic: FAST (FMA, explicit block inverses, dot products in the MMA's order, q2 accumulated in two interleaved partial
// sums; [T xr; S ur] computed once per instance).  EXACT mode, float and the debug payload use the scalar kernel.
#pragma once
// (spcies_mma.cuh is included by the parent header, outside its namespace)

#ifndef SPCIES_EADMM_MMA
#define SPCIES_EADMM_MMA 1
#endif

constexpr bool MMA_SHAPE_OK = n >= 5 && n <= 6 && nm <= 8 && N >= 3;
#ifndef SPCIES_EADMM_PF
#define SPCIES_EADMM_PF 4
#endif
#ifndef SPCIES_EADMM_P1_UNROLL
#define SPCIES_EADMM_P1_UNROLL 4       // measured on C5b (256 Ki instances): 1.694 / 1.726 / 1.648 M solves/s at 2 / 4 / 5; PF 2 / 4 / 6: 1.678 / 1.694 / 1.704
#endif
constexpr int P1U = SPCIES_EADMM_P1_UNROLL;                        // unroll factor of the (stage-parallel) P1 loop
constexpr int PF = SPCIES_EADMM_PF;                               // prefetch distance (stages) of the recurrence fragments
constexpr int MMA_NBLK = 2 * N + 4;                               // z3[N+1], lambda[N+3]   (z1 recomputed, mu' streamed through L2)
constexpr int BLK_Z3 = 0, BLK_LAM = N + 1;
constexpr size_t MMA_STATE_PER_WARP = (size_t)MMA_NBLK * 32 * sizeof(double2);

struct alignas(16) MmaSmall {   // staged into shared memory
    double NAB[64], ABt[64], W2[64], TS[64];                      // -[A B] | [A B]' | W2 | blkdiag(T, S), by column
    double rho[N + 1][8], H1i[N + 1][8], H3i[N + 1][8];
    double rho_0[8], rho_s[8], LB[8], UB[8], LB_0[8], UB_0[8], LB_s[8], UB_s[8];
    int xat[8], uat[8];
};
struct alignas(16) MmaFrag {    // recurrence fragments: shared memory when they fit, else global memory
    double2 FWa[N][32], BWa[N][32];
    double FWb[N][32], BWb[N][32];
};
constexpr size_t SMALL_BYTES = (sizeof(MmaSmall) + 15) / 16 * 16;
constexpr size_t FRAG_BYTES = (sizeof(MmaFrag) + 15) / 16 * 16;
constexpr size_t CONSTS_BYTES_ = (sizeof(spcies_consts) + 15) / 16 * 16;
constexpr size_t MMA_OFFSET = CONSTS_BYTES_;                      // blob: spcies_consts | MmaSmall | MmaFrag
constexpr size_t SMEM_LIMIT = 227 * 1024 - 64;
constexpr int warps_fit(size_t fixed) { return fixed >= SMEM_LIMIT ? 0 : (int)((SMEM_LIMIT - fixed) / MMA_STATE_PER_WARP); }
// fragments in shared memory only if that does not cost a warp (up to one warp per SM scheduler); else streamed from global memory
constexpr int cap4(int w) { return w > 4 ? 4 : w; }
constexpr bool FRAG_SMEM = cap4(warps_fit(SMALL_BYTES + FRAG_BYTES)) >= cap4(warps_fit(SMALL_BYTES));
constexpr int MMA_WARPS_RAW = warps_fit(SMALL_BYTES + (FRAG_SMEM ? FRAG_BYTES : 0));
constexpr int MMA_WARPS = MMA_WARPS_RAW > 8 ? 8 : MMA_WARPS_RAW;
constexpr size_t MUP_BYTES_PER_WARP = (size_t)N * 32 * sizeof(double2);   // the mu' stream of a warp (global scratch)
constexpr int MMA_BLOCK = MMA_WARPS * 32;
constexpr int MMA_IPB = MMA_WARPS * 8;
constexpr size_t MMA_STAGED = SMALL_BYTES + (FRAG_SMEM ? FRAG_BYTES : 0);
constexpr size_t MMA_SMEM = MMA_STAGED + (size_t)MMA_WARPS * MMA_STATE_PER_WARP;
constexpr bool HAS_MMA = SPCIES_EADMM_MMA != 0 && MMA_SHAPE_OK && sizeof(SPCIES_REAL) == 8 && MMA_WARPS >= 1;

static inline void fill_mma_tables(const spcies_consts &C, MmaSmall &S, MmaFrag &Fr) {
    typedef mma::MmaLayout<n, m> L;
    memset(&S, 0, sizeof S);
    memset(&Fr, 0, sizeof Fr);
    for (int c = 0; c < 8; ++c) {
        S.xat[c] = L::x_at(c);
        S.uat[c] = L::u_at(c);
        const int z = L::z_at(c);
        S.LB[c] = S.LB_0[c] = S.LB_s[c] = -1e300;
        S.UB[c] = S.UB_0[c] = S.UB_s[c] = 1e300;
        if (z < 0) continue;
        for (int l = 0; l <= N; ++l) {
            S.rho[l][c] = (double)C.rho[l][z];
            S.H1i[l][c] = (double)C.H1i[l][z];
            S.H3i[l][c] = (double)C.H3i[l][z];
        }
        S.rho_0[c] = (double)C.rho_0[z];
        S.rho_s[c] = (double)C.rho_s[z];
        S.LB[c] = (double)C.LB[z];
        S.UB[c] = (double)C.UB[z];
        S.LB_0[c] = (double)C.LB_0[z];
        S.UB_0[c] = (double)C.UB_0[z];
        S.LB_s[c] = (double)C.LB_s[z];
        S.UB_s[c] = (double)C.UB_s[z];
    }
    for (int oc = 0; oc < 8; ++oc)
        for (int ic = 0; ic < 8; ++ic) {
            const int zo = L::z_at(oc), zi = L::z_at(ic);
            if (L::x_at(oc) >= 0 && zi >= 0) S.NAB[oc * 8 + ic] = -(double)C.AB[L::x_at(oc)][zi];
            if (zo >= 0 && L::x_at(ic) >= 0) S.ABt[oc * 8 + ic] = (double)C.AB[L::x_at(ic)][zo];
            if (zo >= 0 && zi >= 0) {
                S.W2[oc * 8 + ic] = (double)C.W2[zo][zi];
                if (zo < n && zi < n) S.TS[oc * 8 + ic] = (double)C.T[zo][zi];
                if (zo >= n && zi >= n) S.TS[oc * 8 + ic] = (double)C.S[zo - n][zi - n];
            }
        }
    typedef double Blk[n][n];
    Blk *Linv = new Blk[4 * N], *F = Linv + N, *Uinv = F + N, *G = Uinv + N;
    mma::block_inverses<N, n>(C, Linv, F, Uinv, G);
    mma::recurrence_fragments<N, n, m, double>(Linv, F, Uinv, G, Fr.FWa, Fr.FWb, Fr.BWa, Fr.BWb);
    delete[] Linv;
}

__global__ void __launch_bounds__(MMA_BLOCK, 1) eadmm_mma_kernel(const BatchIO io, const unsigned char *__restrict__ g_blob,
                                                                 double2 *__restrict__ g_mup) {
    using mma::dmma;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    const spcies_consts *C = reinterpret_cast<const spcies_consts *>(g_blob);   // engineering-unit scaling only
    (void)C;
    const MmaSmall *T = reinterpret_cast<const MmaSmall *>(smem_raw);
    stage_constants(smem_raw, g_blob + MMA_OFFSET, (uint32_t)MMA_STAGED, &mbar);
    const MmaFrag *Fr = FRAG_SMEM ? reinterpret_cast<const MmaFrag *>(smem_raw + SMALL_BYTES)
                                  : reinterpret_cast<const MmaFrag *>(g_blob + MMA_OFFSET + SMALL_BYTES);

    const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3, warp = threadIdx.x >> 5;
    const int cc[2] = {2 * t4, 2 * t4 + 1};
    const int xe[2] = {T->xat[cc[0]], T->xat[cc[1]]}, ue[2] = {T->uat[cc[0]], T->uat[cc[1]]};
    const bool xs[2] = {xe[0] >= 0, xe[1] >= 0}, us[2] = {ue[0] >= 0, ue[1] >= 0};
    const unsigned gmask = 0xFu << (4 * g);
    const bool leader = t4 == 0, lo2 = t4 < 2;
    double2 *st = reinterpret_cast<double2 *>(smem_raw + MMA_STAGED + warp * MMA_STATE_PER_WARP) + lane;
    double2 *gm = g_mup + ((size_t)blockIdx.x * MMA_WARPS + warp) * (size_t)(N * 32) + lane;    // mu'_l of this lane at gm[l * 32]
    auto LD = [&](int blk) { return st[blk * 32]; };
    auto ST = [&](int blk, double2 v) { st[blk * 32] = v; };
    auto ROW = [&](const double (*tab)[8], int l) { return reinterpret_cast<const double2 *>(tab[l])[t4]; };   // (tab[l][2t], tab[l][2t+1])
    auto ROW1 = [&](const double *row) { return reinterpret_cast<const double2 *>(row)[t4]; };
    auto FWA = [&](int l) { return FRAG_SMEM ? Fr->FWa[l][lane] : __ldg(&Fr->FWa[l][lane]); };
    auto FWB = [&](int l) { return FRAG_SMEM ? Fr->FWb[l][lane] : __ldg(&Fr->FWb[l][lane]); };
    auto BWA = [&](int l) { return FRAG_SMEM ? Fr->BWa[l][lane] : __ldg(&Fr->BWa[l][lane]); };
    auto BWB = [&](int l) { return FRAG_SMEM ? Fr->BWb[l][lane] : __ldg(&Fr->BWb[l][lane]); };

    const double2 nab = reinterpret_cast<const double2 *>(T->NAB)[lane], abt = reinterpret_cast<const double2 *>(T->ABt)[lane];
    const double2 w2 = reinterpret_cast<const double2 *>(T->W2)[lane];
    const double2 rho0 = ROW1(T->rho_0), rhos = ROW1(T->rho_s), lbi = ROW1(T->LB), ubi = ROW1(T->UB);
    const double2 lb0 = ROW1(T->LB_0), ub0 = ROW1(T->UB_0), lbs = ROW1(T->LB_s), ubs = ROW1(T->UB_s);
    const double tol_ = (double)tol;
    const double tolx[2] = {xs[0] ? tol_ : 1e300, xs[1] ? tol_ : 1e300};     // tests that only look at state columns
    const double tolz[2] = {(xs[0] || us[0]) ? tol_ : 1e300, (xs[1] || us[1]) ? tol_ : 1e300};
    const double maskx[2] = {xs[0] ? 1.0 : 0.0, xs[1] ? 1.0 : 0.0};

    const WorkQueue wq{io.queue, io.B, io.ready};
    wq.mark_start();
    unsigned long long stat_k = 0;
    unsigned int stat_nc = 0;
    long long inst = -1;
    int k = 0;
    bool live = false, drained = false;
    double x0v[2] = {0, 0}, tq[2] = {0, 0}, z2[2] = {0, 0};

    for (;;) {
        // ---- refill                                                             code_MPCT_EADMM_C.c:30-83
        const bool need = !live && !drained;
        if (__any_sync(FULL, need)) {
            long long slot = -1;
            if (need && leader) slot = wq.next();
            slot = __shfl_sync(FULL, slot, lane & ~3);
            double ref[2] = {0, 0};
            if (need) {
                if (slot < 0) {
                    drained = true;
                    if (leader) wq.mark_drained();
                } else {
                    inst = slot;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        x0v[i] = xs[i] ? eng_x(C, io.x0, inst, n, xe[i]) : 0.0;       // x0 zero-padded to nm (:35)
                        ref[i] = xs[i] ? eng_x(C, io.xr, inst, n, xe[i]) : (us[i] ? eng_u(C, io.ur, inst, m, ue[i]) : 0.0);
                    }
#pragma unroll 4
                    for (int e = 0; e < MMA_NBLK; ++e) ST(e, make_double2(0.0, 0.0));
                    z2[0] = z2[1] = 0.0;
                    k = 0;
                    live = true;
                }
            }
            // [T xr; S ur] of the instances that just arrived (the product is warp-collective: the other groups recompute theirs)
            double nt[2];
            mma::mv(nt, reinterpret_cast<const double2 *>(T->TS)[lane], ref, 0.0, 0.0);
            if (need && !drained) {
                tq[0] = nt[0];
                tq[1] = nt[1];
            }
            __syncwarp();
        }
        if (!__any_sync(FULL, live)) break;
        bool over = false;
        const double z2o[2] = {z2[0], z2[1]};       // z2 of the previous iteration: what every z1 of this iteration is computed from
        const double2 lam0o = LD(BLK_LAM + 0);      // lambda_0 of the previous iteration (enters z1_0)
        // z1_l, l < N, from the old z3_l / lambda_{l+1} / z2                                 :97-110
        auto z1_of = [&](int l, const double2 z3l, const double2 l1) {
            const double2 rl = ROW(T->rho, l), h1 = ROW(T->H1i, l);
            double2 v;
            if (l == 0) {
                v.x = clip((fma(rho0.x, x0v[0], rl.x * (z3l.x + z2o[0])) + l1.x - lam0o.x) * h1.x, lb0.x, ub0.x);
                v.y = clip((fma(rho0.y, x0v[1], rl.y * (z3l.y + z2o[1])) + l1.y - lam0o.y) * h1.y, lb0.y, ub0.y);
            } else {
                v.x = clip(fma(rl.x, z3l.x + z2o[0], l1.x) * h1.x, lbi.x, ubi.x);
                v.y = clip(fma(rl.y, z3l.y + z2o[1], l1.y) * h1.y, lbi.y, ubi.y);
            }
            return v;
        };

        // ---------- P1, last block first, and the head of q2                              :112-117, :123-136
        double q2a[2], q2b[2] = {0.0, 0.0}, z1N[2];
        {
            const double2 z3N = LD(BLK_Z3 + N), lA = LD(BLK_LAM + N + 1), lB = LD(BLK_LAM + N + 2);
            const double2 rN = ROW(T->rho, N), h1 = ROW(T->H1i, N);
            const double z3v[2] = {z3N.x, z3N.y}, la[2] = {lA.x, lA.y}, lb_[2] = {lB.x, lB.y}, rn[2] = {rN.x, rN.y}, hh[2] = {h1.x, h1.y};
            const double rsv[2] = {rhos.x, rhos.y}, lo[2] = {lbs.x, lbs.y}, hi[2] = {ubs.x, ubs.y};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double rs = rn[i] + rsv[i], rz = rn[i] * z3v[i];
                const double v = clip((fma(rs, z2[i], rz) + la[i] + lb_[i]) * hh[i], lo[i], hi[i]);
                z1N[i] = v;
                q2a[i] = fma(-rs, v, rz) + la[i] + lb_[i] + tq[i];
            }
        }
        // ---------- P1 for l = 0..N-1 fused with the q2 accumulation                      :97-110, :137-141
        {
#pragma unroll P1U
            for (int l = 0; l < N; ++l) {
                const double2 z3l = LD(BLK_Z3 + l), l1 = LD(BLK_LAM + l + 1), rl = ROW(T->rho, l);
                const double2 v = z1_of(l, z3l, l1);
                if (l & 1) {
                    q2b[0] += fma(rl.x, z3l.x - v.x, l1.x);
                    q2b[1] += fma(rl.y, z3l.y - v.y, l1.y);
                } else {
                    q2a[0] += fma(rl.x, z3l.x - v.x, l1.x);
                    q2a[1] += fma(rl.y, z3l.y - v.y, l1.y);
                }
            }
        }
        // ---------- P2: z2 = W2 q2                                                         :145-149, :412-418
        {
            const double q2[2] = {q2a[0] + q2b[0], q2a[1] + q2b[1]};
            double z2n[2];
            mma::mv(z2n, w2, q2, 0.0, 0.0);
            over = over || (fabs(z2[0] - z2n[0]) > tolz[0]) || (fabs(z2[1] - z2n[1]) > tolz[1]);
            z2[0] = z2n[0];
            z2[1] = z2n[1];
        }
        // ---------- P3 forward: t_l = H3i_l o q3_l,  rhs_l = t_{l+1} - [A B] t_l,  mu'_l   :157-184, :221-251
        auto t_of = [&](int l, double (&t)[2]) {     // q3_l = lambda_{l+1} + rho_l (z2 - z1_l)
            const double2 la = LD(BLK_LAM + l + 1), rl = ROW(T->rho, l), h3 = ROW(T->H3i, l);
            const double2 z1 = l == N ? make_double2(z1N[0], z1N[1]) : z1_of(l, LD(BLK_Z3 + l), la);
            t[0] = h3.x * fma(rl.x, z2[0] - z1.x, la.x);
            t[1] = h3.y * fma(rl.y, z2[1] - z1.y, la.y);
        };
        {
            double ta[2], tb[2], mup[2] = {0.0, 0.0};
            t_of(0, ta);
            // the recurrence fragments are fetched PF stages ahead into a register ring (they stream from L2 when they do
            // not fit shared memory; the ring index is static because the stage loop is unrolled by PF)
            double2 fa[PF];
            double fb[PF];
#pragma unroll
            for (int j = 0; j < PF; ++j) {
                fa[j] = FWA(j < N ? j : N - 1);
                fb[j] = FWB(j < N ? j : N - 1);
            }
#pragma unroll 1
            for (int l0 = 0; l0 < N; l0 += PF) {
#pragma unroll
                for (int j = 0; j < PF; ++j) {
                    const int l = l0 + j;
                    if (l < N) {
                        t_of(l + 1, tb);
                        double r[2], e0, e1, f0, f1;
                        mma::mv(r, nab, ta, tb[0], tb[1]);
                        dmma(e0, e1, r[0], fa[j].x, 0.0, 0.0);
                        dmma(f0, f1, mup[0], fa[j].y, e0, e1);               // F_0 = 0
                        dmma(mup[0], mup[1], lo2 ? r[1] : mup[1], fb[j], f0, f1);
                        gm[l * 32] = make_double2(mup[0], mup[1]);          // mu'_l: streamed out, read back by the backward sweep
                        ta[0] = tb[0];
                        ta[1] = tb[1];
                        const int ln = l + PF < N ? l + PF : N - 1;
                        fa[j] = FWA(ln);
                        fb[j] = FWB(ln);
                    }
                }
            }
        }
        double2 z10 = make_double2(0.0, 0.0);     // z1_0 of this iteration (u_opt, res_0): recomputed in the last backward stage
        // ---------- P3 backward + z3 + residual + lambda + exit tests                      :254-320, :371-449
        // close_stage: res = z2 + z3_l - z1_l, lambda_{l+1} += rho_l res, |res|, |z3_prev - z3| tests
        auto close_stage = [&](int l, const double (&z3n)[2], const double2 z1, const double2 z3o, const double2 la) {
            const double2 rl = ROW(T->rho, l);
            const double r0 = (z2[0] + z3n[0]) - z1.x, r1 = (z2[1] + z3n[1]) - z1.y;
            over = over || (fabs(r0) > tolz[0]) || (fabs(r1) > tolz[1]) || (fabs(z3o.x - z3n[0]) > tolz[0]) || (fabs(z3o.y - z3n[1]) > tolz[1]);
            ST(BLK_Z3 + l, make_double2(z3n[0], z3n[1]));
            ST(BLK_LAM + l + 1, make_double2(fma(rl.x, r0, la.x), fma(rl.y, r1, la.y)));
        };
        auto q3_of = [&](int l, double (&q)[2], const double2 z1, const double2 la) {
            const double2 rl = ROW(T->rho, l);
            q[0] = fma(rl.x, z2[0] - z1.x, la.x);
            q[1] = fma(rl.y, z2[1] - z1.y, la.y);
        };
        // old z3_l, lambda_{l+1} and the z1_l they give: read once per stage of the backward sweep, before they are overwritten
        auto stage_in = [&](int l, double2 &z1, double2 &z3o, double2 &la) {
            z3o = LD(BLK_Z3 + l);
            la = LD(BLK_LAM + l + 1);
            z1 = l == N ? make_double2(z1N[0], z1N[1]) : z1_of(l, z3o, la);
        };
        {
            double mu[2], mun[2];     // mu_{l+1}, mu_l
            {
                const double2 mp = gm[(N - 1) * 32], ba = BWA(N - 1);
                const double bb = BWB(N - 1);
                double g0, g1;
                dmma(g0, g1, mp.x, ba.x, 0.0, 0.0);
                dmma(mu[0], mu[1], mp.y, bb, g0, g1);
            }
            {   // z3_N = -H3i_N o (q3_N - [mu_{N-1}; 0])
                double q[2], z3n[2];
                double2 z1, z3o, la;
                stage_in(N, z1, z3o, la);
                q3_of(N, q, z1, la);
                const double2 h3 = ROW(T->H3i, N);
                z3n[0] = -h3.x * fma(-maskx[0], mu[0], q[0]);
                z3n[1] = -h3.y * fma(-maskx[1], mu[1], q[1]);
                close_stage(N, z3n, z1, z3o, la);
            }
            double2 ba[PF], mpr[PF];
            double bb[PF];
#pragma unroll
            for (int j = 0; j < PF; ++j) {
                ba[j] = BWA(N - 2 - j >= 0 ? N - 2 - j : 0);
                bb[j] = BWB(N - 2 - j >= 0 ? N - 2 - j : 0);
                mpr[j] = gm[(N - 2 - j >= 0 ? N - 2 - j : 0) * 32];
            }
#pragma unroll 1
            for (int l0 = N - 2; l0 >= 0; l0 -= PF) {
#pragma unroll
                for (int j = 0; j < PF; ++j) {
                    const int l = l0 - j;
                    if (l >= 0) {
                        const double2 mp = mpr[j];
                        double g0, g1, h0, h1;
                        dmma(g0, g1, mp.x, ba[j].x, 0.0, 0.0);
                        dmma(h0, h1, mu[0], ba[j].y, g0, g1);
                        dmma(mun[0], mun[1], lo2 ? mp.y : mu[1], bb[j], h0, h1);
                        // z3_{l+1} = -H3i_{l+1} o (q3_{l+1} - [mu_l; 0] + [A B]' mu_{l+1})
                        double q[2], a[2], z3n[2];
                        double2 z1, z3o, la;
                        stage_in(l + 1, z1, z3o, la);
                        q3_of(l + 1, q, z1, la);
                        mma::mv(a, abt, mu, fma(-maskx[0], mun[0], q[0]), fma(-maskx[1], mun[1], q[1]));
                        const double2 h3 = ROW(T->H3i, l + 1);
                        z3n[0] = -h3.x * a[0];
                        z3n[1] = -h3.y * a[1];
                        close_stage(l + 1, z3n, z1, z3o, la);
                        mu[0] = mun[0];
                        mu[1] = mun[1];
                        const int ln = l - PF >= 0 ? l - PF : 0;
                        ba[j] = BWA(ln);
                        bb[j] = BWB(ln);
                        mpr[j] = gm[ln * 32];
                    }
                }
            }
            {   // z3_0 = -H3i_0 o (q3_0 + [A B]' mu_0)
                double q[2], a[2], z3n[2];
                double2 z1, z3o, la;
                stage_in(0, z1, z3o, la);
                z10 = z1;
                q3_of(0, q, z1, la);
                mma::mv(a, abt, mu, q[0], q[1]);
                const double2 h3 = ROW(T->H3i, 0);
                z3n[0] = -h3.x * a[0];
                z3n[1] = -h3.y * a[1];
                close_stage(0, z3n, z1, z3o, la);
            }
        }
        // res_0 = z1_0[0:n] - x0, lambda_0;  res_{N+2} = z2 - z1_N, lambda_{N+2}             :371-402
        {
            const double2 l0 = lam0o, lS = LD(BLK_LAM + N + 2);
            const double r0 = z10.x - x0v[0], r1 = z10.y - x0v[1];
            const double s0 = z2[0] - z1N[0], s1 = z2[1] - z1N[1];
            over = over || (fabs(r0) > tolx[0]) || (fabs(r1) > tolx[1]) || (fabs(s0) > tolz[0]) || (fabs(s1) > tolz[1]);
            ST(BLK_LAM + 0, make_double2(xs[0] ? fma(rho0.x, r0, l0.x) : 0.0, xs[1] ? fma(rho0.y, r1, l0.y) : 0.0));
            ST(BLK_LAM + N + 2, make_double2(fma(rhos.x, s0, lS.x), fma(rhos.y, s1, lS.y)));
        }

        // ================= exit condition                                            :408-457 =================
        if (live) k += 1;
        const bool gover = (__ballot_sync(FULL, over) & gmask) != 0u;
        if (live) {
            const int ef = !gover ? 1 : ((k >= k_max) ? -1 : 0);
            if (ef != 0) {
                if (us[0]) io.u[inst * m + ue[0]] = eng_u_out(C, z10.x, ue[0]);                  // u_opt = z1[0][n..]  (:470-478)
                if (us[1]) io.u[inst * m + ue[1]] = eng_u_out(C, z10.y, ue[1]);
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
