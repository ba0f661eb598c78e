// ellipMPC_ADMM_soc_mma.cuh -- tensor-core (DMMA) engine of the ellipMPC ADMM_soc solver (included by ellipMPC_ADMM_soc.cuh,
// inside spcies::soc).
//
// The reference's iteration (code_ellipMPC_ADMM_soc_C.c:149-205) is a chain of sparse products with run-time index arrays,
//     rhs = GhHhi q_hat - bh;   W rhs = rhs (CSC L D L');   primal_hat = Hhi q_hat + HhiGh rhs,
// i.e. one linear map  primal_hat = M1 q_hat - Mb bh  with  M1 = Hhi + HhiGh W^-1 GhHhi  (88 x 88 at N = 10) and
// Mb = HhiGh W^-1 restricted to the 2 n + 1 non-zero rows of bh.  For a batch that shares the model this is the GEMM of the
// HMPC engine (HMPC_ADMM_split_mma.cuh), only small enough to keep the whole fragment table in shared memory:
// a warp owns 8 instances, (z, s) is ZT + 1 tiles of 8 columns (2 columns per lane), [M1 | -Mb] is a table of FP64 MMA B
// fragments [output tile][input tile][lane], the update of an output tile is fused behind its product.  M1 and Mb are formed on
// the host in extended precision from the generated CSR / CSC-LDL constants (the same L, Dinv the reference solves with).
// The SOC projection of s (n + 1 components in one tile) reduces the norm over the 4 lanes of the instance with two shuffles.
//
// Arithmetic: FAST (explicit W^-1, FMA, MMA accumulation order); EXACT mode, float and the debug payload use the scalar kernel.
#pragma once
// (spcies_mma.cuh is included by the parent header, outside its namespace)

#ifndef SPCIES_SOC_MMA
#define SPCIES_SOC_MMA 1
#endif

constexpr int ZT = (DIM + 7) / 8;                 // tiles of z
constexpr int NT = ZT + 1;                        // + s
constexpr int NIN = NT + 2;                       // input tiles: q_hat, then (b = -A x0, r) and (-PhiP xr)
constexpr int NCLIP = DIM - n - 1;                // clipped entries of z (:212)
constexpr bool MMA_SHAPE_OK = NS <= 8 && n <= 7 && nrow_GhHhi == NR && nrow_HhiGh == NP && nrow_Hhi == NP;
constexpr int BLK_P = 0, BLK_D = NT, BLK_QH = 2 * NT;            // q_hat is followed by the two bh tiles
constexpr int NST = 3 * NT + 2;
constexpr size_t MMA_STATE_PER_WARP = (size_t)NST * 32 * sizeof(double2);
constexpr int NB = 4;                             // output tiles per pass over the input (independent accumulators)

struct alignas(16) MmaSmall {
    double LB[ZT][8], UB[ZT][8];                  // bounds of z by tile / column (+-1e300 where z is not clipped)
};
constexpr size_t SMALL_BYTES = (sizeof(MmaSmall) + 15) / 16 * 16;
constexpr int NT_PAD = (NT + NB - 1) / NB * NB;   // output tiles padded to whole passes (zero fragments)
constexpr size_t FRAG_BYTES = (size_t)NT_PAD * NIN * 32 * sizeof(double2);
constexpr size_t CONSTS_BYTES_ = (sizeof(spcies_consts) + 15) / 16 * 16;
constexpr size_t MMA_OFFSET = CONSTS_BYTES_;      // blob: spcies_consts | MmaSmall | fragments
constexpr size_t MMA_STAGED = SMALL_BYTES + FRAG_BYTES;
constexpr size_t SMEM_LIMIT = 227 * 1024 - 64;
constexpr int MMA_WARPS_RAW = MMA_STAGED >= SMEM_LIMIT ? 0 : (int)((SMEM_LIMIT - MMA_STAGED) / MMA_STATE_PER_WARP);
constexpr int MMA_WARPS = MMA_WARPS_RAW > 8 ? 8 : MMA_WARPS_RAW;
constexpr int MMA_BLOCK = MMA_WARPS * 32;
constexpr int MMA_IPB = MMA_WARPS * 8;
constexpr size_t MMA_SMEM = MMA_STAGED + (size_t)MMA_WARPS * MMA_STATE_PER_WARP;
constexpr bool HAS_MMA = SPCIES_SOC_MMA != 0 && MMA_SHAPE_OK && sizeof(SPCIES_REAL) == 8 && MMA_WARPS >= 2;

// element of (z, s) held by (tile, column), or -1
static inline int elem_at(int tile, int col) {
    if (tile < ZT) {
        const int j = tile * 8 + col;
        return j < DIM ? j : -1;
    }
    return (tile == ZT && col < NS) ? DIM + col : -1;
}
// row of bh held by (bh tile, column), or -1
static inline int bh_row_at(int which, int col) {
    if (which == 0) return col < n ? col : (col == n ? NEQ - 1 : -1);
    return col < n ? NEQ + 1 + col : -1;
}

static inline void fill_mma_tables(const spcies_consts &C, MmaSmall &S, double2 *frag) {
    typedef long double ld;
    memset(&S, 0, sizeof S);
    for (int t = 0; t < ZT; ++t)
        for (int c = 0; c < 8; ++c) {
            const int j = t * 8 + c;
            S.LB[t][c] = j < NCLIP ? (double)C.LB[j] : -1e300;
            S.UB[t][c] = j < NCLIP ? (double)C.UB[j] : 1e300;
        }
    // W^-1 applied to the columns of GhHhi and to the unit vectors of the non-zero rows of bh (the reference's LDL solve, :172-188)
    auto ldl_solve = [&](ld *x) {
        for (int i = 0; i < NR; ++i)
            for (int j = C.L_col[i]; j < C.L_col[i + 1]; ++j) x[C.L_row[j]] -= (ld)C.L_val[j] * x[i];
        for (int i = 0; i < NR; ++i) x[i] *= (ld)C.Dinv[i];
        for (int i = NR - 1; i >= 0; --i)
            for (int j = C.L_col[i]; j < C.L_col[i + 1]; ++j) x[i] -= (ld)C.L_val[j] * x[C.L_row[j]];
    };
    ld *X = new ld[(size_t)NR * NP]();          // X[:, c] = W^-1 GhHhi[:, c]   (stored [c][r])
    for (int r = 0; r < NR; ++r)
        for (int j = C.GhHhi_row[r]; j < C.GhHhi_row[r + 1]; ++j) X[(size_t)C.GhHhi_col[j] * NR + r] = (ld)C.GhHhi_val[j];
    for (int c = 0; c < NP; ++c) ldl_solve(X + (size_t)c * NR);
    constexpr int NBH = 2 * 8;
    ld *Y = new ld[(size_t)NBH * NR]();         // Y[k] = W^-1 e_{row(k)}
    for (int k = 0; k < NBH; ++k) {
        const int row = bh_row_at(k / 8, k % 8);
        if (row < 0) continue;
        Y[(size_t)k * NR + row] = 1;
        ldl_solve(Y + (size_t)k * NR);
    }
    ld *M1 = new ld[(size_t)NP * NP]();         // Hhi + HhiGh X
    ld *Mb = new ld[(size_t)NP * NBH]();        // HhiGh Y
    for (int i = 0; i < NP; ++i) {
        for (int j = C.Hhi_row[i]; j < C.Hhi_row[i + 1]; ++j) M1[(size_t)i * NP + C.Hhi_col[j]] += (ld)C.Hhi_val[j];
        for (int j = C.HhiGh_row[i]; j < C.HhiGh_row[i + 1]; ++j) {
            const int r = C.HhiGh_col[j];
            const ld v = (ld)C.HhiGh_val[j];
            for (int c = 0; c < NP; ++c) M1[(size_t)i * NP + c] += v * X[(size_t)c * NR + r];
            for (int k = 0; k < NBH; ++k) Mb[(size_t)i * NBH + k] += v * Y[(size_t)k * NR + r];
        }
    }
    for (int ot = 0; ot < NT_PAD; ++ot)
        for (int it = 0; it < NIN; ++it)
            for (int lane = 0; lane < 32; ++lane) {
                const int row = ot < NT ? elem_at(ot, lane / 4) : -1, t = lane % 4;
                double v[2] = {0.0, 0.0};
                for (int i = 0; i < 2 && row >= 0; ++i) {
                    const int c = 2 * t + i;
                    if (it < NT) {
                        const int col = elem_at(it, c);
                        if (col >= 0) v[i] = (double)M1[(size_t)row * NP + col];
                    } else if (bh_row_at(it - NT, c) >= 0) {
                        v[i] = -(double)Mb[(size_t)row * NBH + (it - NT) * 8 + c];
                    }
                }
                frag[((size_t)ot * NIN + it) * 32 + lane] = make_double2(v[0], v[1]);
            }
    delete[] X;
    delete[] Y;
    delete[] M1;
    delete[] Mb;
}

__global__ void __launch_bounds__(MMA_BLOCK, 1) soc_mma_kernel(const BatchIO io, const unsigned char *__restrict__ g_blob) {
    using mma::dmma;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    const spcies_consts *C = reinterpret_cast<const spcies_consts *>(g_blob);
    const MmaSmall *T = reinterpret_cast<const MmaSmall *>(smem_raw);
    stage_constants(smem_raw, g_blob + MMA_OFFSET, (uint32_t)MMA_STAGED, &mbar);

    const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3, warp = threadIdx.x >> 5;
    const unsigned gmask = 0xFu << (4 * g);
    const bool leader = t4 == 0;
    const double2 *frag = reinterpret_cast<const double2 *>(smem_raw + SMALL_BYTES) + lane;
    double2 *st = reinterpret_cast<double2 *>(smem_raw + MMA_STAGED + warp * MMA_STATE_PER_WARP) + lane;
    auto LD = [&](int blk) { return st[blk * 32]; };
    auto ST = [&](int blk, double2 v) { st[blk * 32] = v; };

    const double sigma_ = C->sigma, sigma_i_ = C->sigma_i, rho_ = C->rho, rho_i_ = C->rho_i;
    const double told = (double)tol_d, tolp = (double)tol_p;
    const int c0 = 2 * t4, c1 = 2 * t4 + 1;
    const bool sin0 = c0 < NS, sin1 = c1 < NS;        // columns of the s tile that hold a component

    const WorkQueue wq{io.queue, io.B, io.ready};
    wq.mark_start();
    unsigned long long stat_k = 0;
    unsigned int stat_nc = 0;
    long long inst = -1;
    int k = 0;
    bool live = false, drained = false;
    double qv[ZT][2];                                 // q by tile (it repeats (Q xr, R ur) per stage, then T xr)
#pragma unroll
    for (int t = 0; t < ZT; ++t) qv[t][0] = qv[t][1] = 0.0;

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
                    double x0[n], xr[n], ur[m], QX[n], QT[n], QU[m];
#pragma unroll
                    for (int i = 0; i < n; ++i) {
                        x0[i] = eng_x(C, io.x0, inst, n, i);
                        xr[i] = eng_x(C, io.xr, inst, n, i);
                    }
#pragma unroll
                    for (int i = 0; i < m; ++i) ur[i] = eng_u(C, io.ur, inst, m, i);
#pragma unroll
                    for (int j = 0; j < n; ++j) {
                        double qx = 0.0, qt = 0.0;
#pragma unroll
                        for (int i = 0; i < n; ++i) {
                            qx = fma(C->Q[j][i], xr[i], qx);          // Q, R, T stored negated
                            qt = fma(C->T[j][i], xr[i], qt);
                        }
                        QX[j] = qx;
                        QT[j] = qt;
                    }
#pragma unroll
                    for (int j = 0; j < m; ++j) {
                        double qu = 0.0;
#pragma unroll
                        for (int i = 0; i < m; ++i) qu = fma(C->R[j][i], ur[i], qu);
                        QU[j] = qu;
                    }
                    auto q_at = [&](int idx) -> double {               // :102-131
                        if (idx >= DIM) return 0.0;
                        if (idx < m) return QU[idx];
                        if (idx >= m + (N - 1) * nm) return (idx < DIM - 1) ? QT[idx - m - (N - 1) * nm] : 0.0;
                        const int t = (idx - m) % nm;
                        return (t < n) ? QX[t] : QU[t - n];
                    };
#pragma unroll
                    for (int t = 0; t < ZT; ++t) {
                        qv[t][0] = q_at(8 * t + c0);
                        qv[t][1] = q_at(8 * t + c1);
                    }
                    // bh: (b = -A x0, r) and (-PhiP xr)
                    double bA[2] = {0.0, 0.0}, bB[2] = {0.0, 0.0};
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int c = 2 * t4 + i;
                        if (c < n) {
                            for (int j = 0; j < n; ++j) {
                                bA[i] = fma(-C->A[c][j], x0[j], bA[i]);
                                bB[i] = fma(-C->PhiP[c][j], xr[j], bB[i]);
                            }
                        } else if (c == n) {
                            bA[i] = io.r[inst];
                        }
                    }
#pragma unroll 4
                    for (int e = 0; e < 2 * NT; ++e) ST(e, make_double2(0.0, 0.0));
                    ST(BLK_QH + NT, make_double2(bA[0], bA[1]));
                    ST(BLK_QH + NT + 1, make_double2(bB[0], bB[1]));
                    k = 0;
                    live = true;
                }
            }
            __syncwarp();
        }
        if (!__any_sync(FULL, live)) break;

        // ---- q_hat = [q + lambda - sigma z ; mu - rho s]                                              :149-154
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const double2 p = LD(BLK_P + t), d = LD(BLK_D + t);
            double2 qh;
            if (t < ZT) {
                qh.x = fma(-sigma_, p.x, qv[t][0] + d.x);
                qh.y = fma(-sigma_, p.y, qv[t][1] + d.y);
            } else {
                qh.x = fma(-rho_, p.x, d.x);
                qh.y = fma(-rho_, p.y, d.y);
            }
            ST(BLK_QH + t, qh);
        }

        // ---- primal_hat = M1 q_hat - Mb bh, NB output tiles per pass, fused with their update         :157-281
        bool over = false;
#pragma unroll 1
        for (int t0 = 0; t0 < NT; t0 += NB) {
            double acc[NB][2];
#pragma unroll
            for (int b = 0; b < NB; ++b) acc[b][0] = acc[b][1] = 0.0;
            const double2 *fr = frag + (size_t)t0 * NIN * 32;
#pragma unroll 2
            for (int it = 0; it < NIN; ++it) {
                const double2 v = LD(BLK_QH + it);
                double2 f[NB];
#pragma unroll
                for (int b = 0; b < NB; ++b) f[b] = fr[((size_t)b * NIN + it) * 32];
#pragma unroll
                for (int b = 0; b < NB; ++b) dmma(acc[b][0], acc[b][1], v.x, f[b].x, acc[b][0], acc[b][1]);
#pragma unroll
                for (int b = 0; b < NB; ++b) dmma(acc[b][0], acc[b][1], v.y, f[b].y, acc[b][0], acc[b][1]);
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int t = t0 + b;
                if (t < ZT) {
                    // z = clip(z_hat + lambda / sigma), lambda += sigma (z_hat - z), exit tests          :210-217, :247-249, :260-270
                    const double2 zo = LD(BLK_P + t), lam = LD(BLK_D + t);
                    const double2 lo = reinterpret_cast<const double2 *>(T->LB[t])[t4], hi = reinterpret_cast<const double2 *>(T->UB[t])[t4];
                    const double z0 = clip(fma(sigma_i_, lam.x, acc[b][0]), lo.x, hi.x);
                    const double z1 = clip(fma(sigma_i_, lam.y, acc[b][1]), lo.y, hi.y);
                    over = over || (fabs(zo.x - z0) > told) || (fabs(z0 - acc[b][0]) > tolp) || (fabs(zo.y - z1) > told) ||
                           (fabs(z1 - acc[b][1]) > tolp);
                    ST(BLK_P + t, make_double2(z0, z1));
                    ST(BLK_D + t, make_double2(fma(sigma_, acc[b][0] - z0, lam.x), fma(sigma_, acc[b][1] - z1, lam.y)));
                } else if (t == ZT) {
                    // s = proj_SOC(s_hat + mu / rho), mu += rho (s_hat - s)                              :220-242, :252-254
                    const double2 so = LD(BLK_P + t), mu = LD(BLK_D + t);
                    double s0 = sin0 ? fma(rho_i_, mu.x, acc[b][0]) : 0.0, s1 = sin1 ? fma(rho_i_, mu.y, acc[b][1]) : 0.0;
                    double part = (c0 >= 1 ? s0 * s0 : 0.0) + s1 * s1;         // column 0 is the cone's axis
                    part += __shfl_xor_sync(FULL, part, 1);
                    part += __shfl_xor_sync(FULL, part, 2);
                    const double nrm = sqrt(part);
                    const double x0c = __shfl_sync(FULL, s0, lane & ~3);       // s[0] lives in lane 0 of the group
                    if (nrm <= x0c) {
                    } else if (nrm <= -x0c) {
                        s0 = s1 = 0.0;
                    } else {
                        const double step = (x0c + nrm) / (2.0 * nrm);
                        s0 = (c0 == 0) ? step * nrm : step * s0;
                        s1 = step * s1;
                    }
                    over = over || (sin0 && ((fabs(so.x - s0) > told) || (fabs(s0 - acc[b][0]) > tolp))) ||
                           (sin1 && ((fabs(so.y - s1) > told) || (fabs(s1 - acc[b][1]) > tolp)));
                    ST(BLK_P + t, make_double2(s0, s1));
                    ST(BLK_D + t, make_double2(sin0 ? fma(rho_, acc[b][0] - s0, mu.x) : 0.0, sin1 ? fma(rho_, acc[b][1] - s1, mu.y) : 0.0));
                }
            }
        }

        // ================= exit condition                                            :260-283 =================
        if (live) k += 1;
        const bool gover = (__ballot_sync(FULL, over) & gmask) != 0u;
        if (live) {
            const int ef = !gover ? 1 : ((k >= k_max) ? -1 : 0);
            if (ef != 0) {
                const double2 z0 = LD(BLK_P + 0);                        // u_opt = z[0..m)   (:297-306)
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
