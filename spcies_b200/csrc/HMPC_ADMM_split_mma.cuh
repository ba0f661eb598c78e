// HMPC_ADMM_split_mma.cuh -- tensor-core (DMMA) engine of the HMPC (S)ADMM_split solver (included by HMPC_ADMM_split.cuh,
// inside spcies::hmpc).
//
// The hot loop of code_HMPC_ADMM_split_C.c:176-188 is the dense product primal_hat = M2 bh - M1 q_hat with
// M1 = (dim + n_s)^2 = 442^2 at N = 50: 195 k sequential FMA per instance and iteration for a thread that owns an instance,
// reading a 1.5 MB matrix.  For a batch that shares M1 it is a GEMM: here a warp owns 8 instances (the rows of the m8n8k4 FP64
// MMA), every (z, s) vector is NT = 56 tiles of 8 columns (lane (g, t) of instance g holds columns 2t, 2t+1 of each tile: the
// C/D and -- one register per k-step -- the A fragment layout, spcies_mma.cuh), and [-M1 | M2] is a table of B fragments
// [output tile][input tile][lane] (double2) streamed from L2 through a register ring, two output tiles at a time so that two
// independent accumulators hide the 26-cycle MMA latency.  The update of an output tile (dual step, z = clip(z_hat +
// lambda / sigma), exit tests) is fused behind its product, so primal_hat is never stored.
//
// Column order: z_j -> tile j / 8, column j % 8; the s part is permuted so that the three tiles after z hold y_e, y_s, y_c
// (s[3 g + c] -> tile ZT + c, column g): a lane then owns complete (y_e, y_s, y_c) triples for the diamond-set projections
// (:255-258), no shuffles.  q has 2 n + m non-zeros (x_e, x_c, u_e): kept in registers for the tiles they fall in.
//
// State per warp in shared memory: primal, dual, q_hat = 3 NT tiles of 512 B (84 KB at N = 50 -> two warps per SM).
// Arithmetic: FAST (FMA, MMA accumulation order); EXACT mode, float and the debug payload use the scalar kernel.
#pragma once
// (spcies_mma.cuh is included by the parent header, outside its namespace)

#ifndef SPCIES_HMPC_MMA
#define SPCIES_HMPC_MMA 1
#endif

constexpr int ZT = (DIM + 7) / 8;                 // tiles of z
constexpr int NT = ZT + 3;                        // + y_e, y_s, y_c
constexpr int NIN = NT + 1;                       // input tiles of the product: q_hat, then bh
constexpr int NCLIP = DIM - 3 * n - 3 * m;        // clipped entries of z (:233)
constexpr int QT0 = Q0 / 8, QT1 = (Q0 + 3 * n + m - 1) / 8, QTN = QT1 - QT0 + 1;   // tiles that hold non-zeros of q
#ifndef SPCIES_HMPC_MMA_NB
#define SPCIES_HMPC_MMA_NB 4                      // output tiles per pass over q_hat (independent accumulators)
#endif
#ifndef SPCIES_HMPC_MMA_PF
#define SPCIES_HMPC_MMA_PF 8                      // prefetch distance of the fragment stream (input tiles): PF x NB x 32 cycles of MMA
#endif                                            // issue must cover the L2 latency; the ring is PF x NB x 512 B per warp
constexpr int NBZ = SPCIES_HMPC_MMA_NB, PFZ = SPCIES_HMPC_MMA_PF;
constexpr bool MMA_SHAPE_OK = NS == 3 * nm && nm <= 8 && n <= 8 && QTN <= 4;
constexpr size_t MMA_RING_PER_WARP = (size_t)PFZ * NBZ * 32 * sizeof(double2);
constexpr int NST = 3 * NT + 2;                   // primal, dual, q_hat tiles; then bh (the last input tile) and a dummy
constexpr size_t MMA_STATE_PER_WARP = (size_t)NST * 32 * sizeof(double2) + MMA_RING_PER_WARP;
constexpr int BLK_P = 0, BLK_D = NT, BLK_QH = 2 * NT;

struct alignas(16) MmaSmall {                     // staged into shared memory
    double LB[ZT][8], UB[ZT][8];                  // bounds of z by tile / column (+-1e300 where z is not clipped)
    double LBy[8], UBy[8];
};
constexpr size_t SMALL_BYTES = (sizeof(MmaSmall) + 15) / 16 * 16;
constexpr size_t FRAG_BYTES = ((size_t)NT * NIN + 16) * 32 * sizeof(double2);   // [-M1 | M2] fragments (+ padding), global memory
constexpr size_t CONSTS_BYTES_ = (sizeof(spcies_consts) + 15) / 16 * 16;
constexpr size_t MMA_OFFSET = CONSTS_BYTES_;      // blob: spcies_consts | MmaSmall | fragments
constexpr size_t SMEM_LIMIT = 227 * 1024 - 64;
constexpr int MMA_WARPS_RAW = SMALL_BYTES >= SMEM_LIMIT ? 0 : (int)((SMEM_LIMIT - SMALL_BYTES) / MMA_STATE_PER_WARP);
constexpr int MMA_WARPS = MMA_WARPS_RAW > 8 ? 8 : MMA_WARPS_RAW;
constexpr int MMA_BLOCK = MMA_WARPS * 32;
constexpr int MMA_IPB = MMA_WARPS * 8;
constexpr size_t MMA_SMEM = SMALL_BYTES + (size_t)MMA_WARPS * MMA_STATE_PER_WARP;
constexpr bool HAS_MMA = SPCIES_HMPC_MMA != 0 && MMA_SHAPE_OK && sizeof(SPCIES_REAL) == 8 && MMA_WARPS >= 1;

// element of (z, s) held by (tile, column), or -1
static inline int elem_at(int tile, int col) {
    if (tile < ZT) {
        const int j = tile * 8 + col;
        return j < DIM ? j : -1;
    }
    return col < nm ? DIM + 3 * col + (tile - ZT) : -1;
}

static inline void fill_mma_tables(const spcies_consts &C, MmaSmall &S, double2 *frag) {
    memset(&S, 0, sizeof S);
    for (int t = 0; t < ZT; ++t)
        for (int c = 0; c < 8; ++c) {
            const int j = t * 8 + c;
            S.LB[t][c] = j < NCLIP ? (double)C.LB[j] : -1e300;
            S.UB[t][c] = j < NCLIP ? (double)C.UB[j] : 1e300;
        }
    for (int g = 0; g < 8; ++g) {
        S.LBy[g] = g < nm ? (double)C.LBy[g] : 0.0;
        S.UBy[g] = g < nm ? (double)C.UBy[g] : 0.0;
    }
    for (int ot = 0; ot < NT; ++ot)
        for (int it = 0; it < NIN; ++it)
            for (int lane = 0; lane < 32; ++lane) {
                const int row = elem_at(ot, lane / 4), t = lane % 4;
                double v[2] = {0.0, 0.0};
                for (int i = 0; i < 2; ++i) {
                    const int c = 2 * t + i;
                    if (row < 0) continue;
                    if (it < NT) {
                        const int col = elem_at(it, c);
                        if (col >= 0) v[i] = -(double)C.M1[row][col];
                    } else if (c < n) {
                        v[i] = (double)C.M2[row][c];
                    }
                }
                frag[((size_t)ot * NIN + it) * 32 + lane] = make_double2(v[0], v[1]);
            }
}

// acc[b] = sum over the NIN input tiles of (fragment of output tile ot + b, input tile it) x (q_hat tile it | bh): NB independent
// accumulators.  The fragments stream from L2 into a per-warp shared-memory ring with cp.async (LDGSTS), PF input tiles ahead;
// every lane copies and later reads only its own 16 bytes, so the ring needs no cross-lane synchronisation.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int K> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(K) : "memory"); }

// The loop is kept free of integer overhead (it was issue bound: 3 IMAD per DMMA): pointers advance by constants, bh is tile NT
// of the q_hat array (tile NT + 1 is a dummy for the last prefetch), the fragment table is padded by PF tiles so that the refill
// needs no bound check, and the body is unrolled by two so that the "next -> current" operand hand-over is a renaming.
template <int NB, int PF>
__device__ __forceinline__ void hmpc_product(const double2 *__restrict__ fr, const double2 *qh /* q_hat tile 0 of this lane */,
                                             double2 *ring /* [PF][NB][32], + lane */, double (&acc)[NB][2]) {
    using mma::dmma;
#pragma unroll
    for (int j = 0; j < PF; ++j) {
#pragma unroll
        for (int b = 0; b < NB; ++b) cp_async16(ring + (j * NB + b) * 32, fr + ((size_t)b * NIN + j) * 32);
        cp_async_commit();
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b][0] = acc[b][1] = 0.0;
    // operands of an input tile (q_hat tile, fragments) are read from shared memory one tile ahead of their MMAs
    double2 v = qh[0], f[NB];
    cp_async_wait<PF - 1>();                                               // the group of input tile 0 has landed
#pragma unroll
    for (int b = 0; b < NB; ++b) f[b] = ring[b * 32];
    const double2 *gnext = fr + PF * 32;                                   // fragments of input tile it + PF
    const double2 *qn = qh + 32;                                           // q_hat tile it + 1
    double2 *rcur = ring;                                                  // ring slot of input tile it
    double2 *const rend = ring + PF * NB * 32;
#pragma unroll 2
    for (int it = 0; it < NIN; ++it) {
        double2 *rnxt = rcur + NB * 32;
        rnxt = rnxt == rend ? ring : rnxt;
        const double2 vn = *qn;
        double2 fn[NB];
        cp_async_wait<PF - 2>();                                            // ... and the group of input tile it + 1
#pragma unroll
        for (int b = 0; b < NB; ++b) fn[b] = rnxt[b * 32];
#pragma unroll
        for (int b = 0; b < NB; ++b) dmma(acc[b][0], acc[b][1], v.x, f[b].x, acc[b][0], acc[b][1]);
#pragma unroll
        for (int b = 0; b < NB; ++b) dmma(acc[b][0], acc[b][1], v.y, f[b].y, acc[b][0], acc[b][1]);
#pragma unroll
        for (int b = 0; b < NB; ++b) cp_async16(rcur + b * 32, gnext + (size_t)b * NIN * 32);
        cp_async_commit();
        gnext += 32;
        qn += 32;
        rcur = rnxt;
        v = vn;
#pragma unroll
        for (int b = 0; b < NB; ++b) f[b] = fn[b];
    }
    cp_async_wait<0>();
}

__global__ void __launch_bounds__(MMA_BLOCK, 1) hmpc_mma_kernel(const BatchIO io, const unsigned char *__restrict__ g_blob) {
    using mma::dmma;
    typedef Arith<double, false> A;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    const spcies_consts *C = reinterpret_cast<const spcies_consts *>(g_blob);
    const MmaSmall *T = reinterpret_cast<const MmaSmall *>(smem_raw);
    stage_constants(smem_raw, g_blob + MMA_OFFSET, (uint32_t)SMALL_BYTES, &mbar);
    const double2 *frag = reinterpret_cast<const double2 *>(g_blob + MMA_OFFSET + SMALL_BYTES);

    const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3, warp = threadIdx.x >> 5;
    const unsigned gmask = 0xFu << (4 * g);
    const bool leader = t4 == 0;
    double2 *st = reinterpret_cast<double2 *>(smem_raw + SMALL_BYTES + warp * MMA_STATE_PER_WARP) + lane;
    double2 *ring = st + NST * 32;
    auto LD = [&](int blk) { return st[blk * 32]; };
    auto ST = [&](int blk, double2 v) { st[blk * 32] = v; };

    const double sigma_ = C->sigma, sigma_i_ = C->sigma_i, rho_ = C->rho, rho_i_ = C->rho_i;
    const double as = SYMMETRIC ? (double)SPCIES_ALPHA * sigma_ : sigma_;      // alpha_SADMM * sigma
    const double ar = SYMMETRIC ? (double)SPCIES_ALPHA * rho_ : rho_;
    const double told = (double)tol_d, tolp = (double)tol_p;
    const double lby[2] = {T->LBy[2 * t4], T->LBy[2 * t4 + 1]}, uby[2] = {T->UBy[2 * t4], T->UBy[2 * t4 + 1]};
    const bool trip[2] = {2 * t4 < nm, 2 * t4 + 1 < nm};

    const WorkQueue wq{io.queue, io.B, io.ready};
    wq.mark_start();
    unsigned long long stat_k = 0;
    unsigned int stat_nc = 0;
    long long inst = -1;
    int k = 0;
    bool live = false, drained = false;
    double bh[2] = {0, 0}, qv[QTN][2];
#pragma unroll
    for (int q = 0; q < QTN; ++q) qv[q][0] = qv[q][1] = 0.0;

    for (;;) {
        // ---- refill: bh = -A x0;  q: -Te xr - QQ x0 at x_e, -QQ x0 at x_c, -Se ur at u_e     code_HMPC_ADMM_split_C.c:99-129
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
                    double x0[n], xr[n], ur[m];
#pragma unroll
                    for (int i = 0; i < n; ++i) {
                        x0[i] = eng_x(C, io.x0, inst, n, i);
                        xr[i] = eng_x(C, io.xr, inst, n, i);
                    }
#pragma unroll
                    for (int i = 0; i < m; ++i) ur[i] = eng_u(C, io.ur, inst, m, i);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int c = 2 * t4 + i;
                        double b = 0.0;
                        if (c < n)
                            for (int j = 0; j < n; ++j) b = fma(-C->A[c][j], x0[j], b);
                        bh[i] = b;
#pragma unroll
                        for (int q = 0; q < QTN; ++q) {
                            const int e = (QT0 + q) * 8 + c - Q0;          // offset into (x_e, x_s, x_c, u_e, ...)
                            double v = 0.0;
                            if (e >= 0 && e < n) {
                                for (int j = 0; j < n; ++j) v -= fma(C->QQ[e][j], x0[j], C->Te[e][j] * xr[j]);
                            } else if (e >= 2 * n && e < 3 * n) {
                                for (int j = 0; j < n; ++j) v = fma(-C->QQ[e - 2 * n][j], x0[j], v);
                            } else if (e >= 3 * n && e < 3 * n + m) {
                                for (int j = 0; j < m; ++j) v = fma(-C->Se[e - 3 * n][j], ur[j], v);
                            }
                            qv[q][i] = v;
                        }
                    }
#pragma unroll 4
                    for (int e = 0; e < 2 * NT; ++e) ST(e, make_double2(0.0, 0.0));
                    ST(BLK_QH + NT, make_double2(bh[0], bh[1]));
                    ST(BLK_QH + NT + 1, make_double2(0.0, 0.0));
                    k = 0;
                    live = true;
                }
            }
            __syncwarp();
        }
        if (!__any_sync(FULL, live)) break;

        // ---- q_hat = [sigma z - q - lambda ; rho s - mu]                                              :149-154
#pragma unroll 1
        for (int t = 0; t < NT; ++t) {
            const double2 p = LD(BLK_P + t), d = LD(BLK_D + t);
            double2 qh;
            if (t < ZT) {
                double q0 = 0.0, q1 = 0.0;
#pragma unroll
                for (int q = 0; q < QTN; ++q)
                    if (t == QT0 + q) {
                        q0 = qv[q][0];
                        q1 = qv[q][1];
                    }
                qh.x = sigma_ * p.x - q0 - d.x;
                qh.y = sigma_ * p.y - q1 - d.y;
            } else {
                qh.x = rho_ * p.x - d.x;
                qh.y = rho_ * p.y - d.y;
            }
            ST(BLK_QH + t, qh);
        }
        __syncwarp();

        // ---- primal_hat = M2 bh - M1 q_hat, NB output tiles at a time, fused with their update        :176-188, :215-334
        bool over = false;
        // z tile: [SADMM dual step], z = clip(z_hat + lambda / sigma), dual step, exit tests        :215-219, :230-238, :288-334
        auto update_z = [&](int t, const double (&zh)[2]) {
            const double2 zo = LD(BLK_P + t), lam0 = LD(BLK_D + t);
            const double2 lo = reinterpret_cast<const double2 *>(T->LB[t])[t4], hi = reinterpret_cast<const double2 *>(T->UB[t])[t4];
            const double zov[2] = {zo.x, zo.y}, lov[2] = {lo.x, lo.y}, hiv[2] = {hi.x, hi.y};
            double lam[2] = {lam0.x, lam0.y}, z[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (SYMMETRIC) lam[i] = fma(as, zh[i] - zov[i], lam[i]);
                z[i] = clip(fma(sigma_i_, lam[i], zh[i]), lov[i], hiv[i]);
                lam[i] = fma(as, zh[i] - z[i], lam[i]);
                over = over || (fabs(zov[i] - z[i]) > told) || (fabs(z[i] - zh[i]) > tolp);
            }
            ST(BLK_P + t, make_double2(z[0], z[1]));
            ST(BLK_D + t, make_double2(lam[0], lam[1]));
        };
        {
            int t = 0;
#pragma unroll 1
            for (; t + NBZ <= ZT; t += NBZ) {
                double acc[NBZ][2];
                hmpc_product<NBZ, PFZ>(frag + ((size_t)t * NIN) * 32 + lane, st + BLK_QH * 32, ring, acc);
#pragma unroll
                for (int b = 0; b < NBZ; ++b) update_z(t + b, acc[b]);
            }
            // the remaining ZT % NBZ tiles of z in one pass of their own (a single-tile pass would be a 2 NIN-deep dependent chain)
            constexpr int REM = ZT % NBZ, REMD = REM > 0 ? REM : 1;   // (this kernel is not a template: the branch is compiled for REM = 0 too)
            if constexpr (REM > 0) {
                double acc[REMD][2];
                hmpc_product<REMD, PFZ>(frag + ((size_t)t * NIN) * 32 + lane, st + BLK_QH * 32, ring, acc);
#pragma unroll
                for (int b = 0; b < REM; ++b) update_z(t + b, acc[b]);
            }
        }
        {   // s, mu: diamond-set projection of each (y_e, y_s, y_c) triple                              :220-225, :241-258, :294-311
            double acc[3][2];
            hmpc_product<3, PFZ>(frag + ((size_t)ZT * NIN) * 32 + lane, st + BLK_QH * 32, ring, acc);
            const double2 so0 = LD(BLK_P + ZT), so1 = LD(BLK_P + ZT + 1), so2 = LD(BLK_P + ZT + 2);
            const double2 mu0 = LD(BLK_D + ZT), mu1 = LD(BLK_D + ZT + 1), mu2 = LD(BLK_D + ZT + 2);
            const double so[3][2] = {{so0.x, so0.y}, {so1.x, so1.y}, {so2.x, so2.y}};
            double mu[3][2] = {{mu0.x, mu0.y}, {mu1.x, mu1.y}, {mu2.x, mu2.y}}, sn[3][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double sv[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (SYMMETRIC) mu[c][i] = fma(ar, acc[c][i] - so[c][i], mu[c][i]);
                    sv[c] = fma(rho_i_, mu[c][i], acc[c][i]);
                }
                proj_soc3<A>(sv, 1.0, lby[i]);
                proj_soc3<A>(sv, -1.0, uby[i]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    sn[c][i] = trip[i] ? sv[c] : 0.0;
                    mu[c][i] = trip[i] ? fma(ar, acc[c][i] - sv[c], mu[c][i]) : 0.0;
                    over = over || (trip[i] && ((fabs(so[c][i] - sv[c]) > told) || (fabs(sv[c] - acc[c][i]) > tolp)));
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                ST(BLK_P + ZT + c, make_double2(sn[c][0], sn[c][1]));
                ST(BLK_D + ZT + c, make_double2(mu[c][0], mu[c][1]));
            }
        }

        // ================= exit condition                                            :318-347 =================
        if (live) k += 1;
        const bool gover = (__ballot_sync(FULL, over) & gmask) != 0u;
        if (live) {
            const int ef = !gover ? 1 : ((k >= k_max) ? -1 : 0);
            if (ef != 0) {
                const double2 z0 = LD(BLK_P + 0);                        // u_opt = z[0..m)   (:359-368)
                if (2 * t4 < m) io.u[inst * m + 2 * t4] = eng_u_out(C, z0.x, 2 * t4);
                if (2 * t4 + 1 < m) io.u[inst * m + 2 * t4 + 1] = eng_u_out(C, z0.y, 2 * t4 + 1);
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
