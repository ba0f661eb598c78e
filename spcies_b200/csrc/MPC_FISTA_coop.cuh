// MPC_FISTA_coop.cuh -- cooperative tail kernel of the FISTA solver (included by MPC_FISTA.cuh, inside spcies::fista).
//
// The second launch of a large batch (io.phase = 2) only holds the ~1 % slowest instances, half of which run to k_max:
// its duration is (iterations left) x (time of ONE iteration of ONE instance), and a thread that owns a whole instance
// cannot run its ~5200-instruction iteration in less than ~6 us.  Here GRP = 8 lanes share one instance, lane j owning
// component j of every stage vector (z_l[j], r_l[j], mu_l[j], w_l[j], d_lambda_l[j], lambda_l[j], y_l[j]), and the
// iteration is reorganised around what is parallel and what is not:
//
//     step 1   z_l(y)            all stages at once: N independent chains per lane                (:494-537)
//     step 2   r_l = b - G z     all stages at once                                              (:546-574)
//              exit test         one ballot per group                                             (:337-361)
//     step 3   s_l = Linv_l r_l  all stages at once
//     step 4   mu_l = s_l - F_l mu_{l-1}      the only sequential part of the forward solve: one n-term dot product and
//              w_{l-1} = Uinv_{l-1} mu_{l-1}  one exchange per stage; w fills its latency slots
//     step 5   dl_l = w_l - G_l dl_{l+1}      sequential part of the backward solve, lambda / y update of stage l+1 in
//              lambda, y update               its latency slots                                   (:368-385)
//
// A lane needs whole vectors for the mat-vecs; they are exchanged through shared memory (own component stored,
// __syncwarp, 128-bit broadcast loads of the vector).  Warps never synchronise with each other, and the whole tail
// (every parked instance) is resident at once: 72 instances per SM.
//
// The arithmetic is the FAST arithmetic of fista_kernel, operation for operation and in the same order per component
// (zero-padded triangular rows add exact zeros), so an instance gets the same bits whichever kernel iterates it: results
// do not depend on when an instance was parked.  EXACT mode does not use this kernel (the reference's triangular
// substitutions are sequential in the component index); neither do calls that ask for the debug payload.
#pragma once

#ifndef SPCIES_FISTA_COOP_BLOCK
#define SPCIES_FISTA_COOP_BLOCK 576
#endif

constexpr int GRP = nm <= 8 ? 8 : (nm <= 16 ? 16 : 32);   // lanes per instance
constexpr int IPW = 32 / GRP;                             // instances per warp
constexpr int COOP_BLOCK = SPCIES_FISTA_COOP_BLOCK;
constexpr int COOP_INST = COOP_BLOCK / GRP;               // instances resident per CTA
constexpr int VEC = 16 / (int)sizeof(real);               // reals per 128-bit access
constexpr int al_(int x) { return (x + VEC - 1) / VEC * VEC; }
// per-instance shared memory (reals).  Z (step 1-2) and MU (step 4) share one region, R (step 2-3) and DL (step 5) another.
constexpr int C_Y = 0, C_LAM = al_(N * n);
constexpr int C_ZM = C_LAM + al_(N * n);                  // z[N-1][nm]  |  mu[N][n]
constexpr int C_ZM_LEN = al_((N - 1) * nm > N * n ? (N - 1) * nm : N * n);
constexpr int C_RD = C_ZM + C_ZM_LEN;                     // r[N][n]     |  dl[N][n]
constexpr int C_END = C_RD + al_(N * n);
// stride = 8 words (mod 32): the IPW groups of a warp start 8 banks apart, so a 128-bit broadcast load of the four
// groups is one wavefront and a 64-bit per-lane access the minimum two
constexpr int C_STRIDE = sizeof(real) == 8 ? ((C_END + 11) / 16 * 16 + 4) : ((C_END + 23) / 32 * 32 + 8);
constexpr size_t COOP_SMEM = BLOB_BYTES + (size_t)COOP_INST * C_STRIDE * sizeof(real);
constexpr bool VEC_OK = (n % VEC == 0 || (sizeof(real) == 8 && n % 2 == 0)) && (nm % 2 == 0);
constexpr bool HAS_COOP = nm <= 32 && N >= 3 && COOP_SMEM <= SMEM_MAX - 64;

// CNT consecutive reals -> registers; 128-bit loads when the layout guarantees the alignment (n, nm even; 16-byte bases)
template <int CNT> __device__ __forceinline__ void ldv(real (&v)[CNT], const real *p) {
    if constexpr (sizeof(real) == 8 && CNT % 2 == 0 && VEC_OK) {
#pragma unroll
        for (int c = 0; c < CNT / 2; ++c) {
            const double2 t2 = reinterpret_cast<const double2 *>(p)[c];
            v[2 * c] = (real)t2.x;
            v[2 * c + 1] = (real)t2.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < CNT; ++c) v[c] = p[c];
    }
}

template <bool VARB>
__global__ void __launch_bounds__(COOP_BLOCK, 1) fista_coop_kernel(const BatchIO io, const spcies_consts *__restrict__ g_consts) {
    typedef Arith<real, false> A;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    spcies_consts *C = reinterpret_cast<spcies_consts *>(smem_raw);
    const FistaDerived *D = reinterpret_cast<const FistaDerived *>(smem_raw + CONSTS_BYTES);
    stage_constants(C, g_consts, (uint32_t)BLOB_BYTES, &mbar);

    const int lane = threadIdx.x & 31;
    const int j = lane & (GRP - 1);                 // component owned by this lane
    const int gl0 = lane & ~(GRP - 1);              // first lane of the group
    const unsigned gmask = (GRP == 32 ? FULL : ((1u << GRP) - 1u)) << gl0;
    const bool is_x = j < n;                        // owns a state / dual component
    const bool is_z = j < nm;                       // owns a component of z
    const int jx = is_x ? j : n - 1;                // clamped indices for addressing
    const int jz = is_z ? j : nm - 1;
    real *st = reinterpret_cast<real *>(smem_raw + BLOB_BYTES) + (size_t)(threadIdx.x / GRP) * C_STRIDE;
    real *Y = st + C_Y, *LAM = st + C_LAM, *Z = st + C_ZM, *MU = st + C_ZM, *R = st + C_RD, *DL = st + C_RD;

    const real qri = C->QRi[jz];
    const real tol_ = (real)tol;

    const long long B = (long long)io.queue[6];
    const WorkQueue wq{io.queue + 7, B, nullptr};
    const WorkQueue marks{io.queue, B, nullptr};
    marks.mark_start();
    unsigned long long stat_k = 0;
    unsigned int stat_nc = 0;

    long long inst = -1;
    int k = 0;
    bool live = false, drained = false;
    real t = real(1), bj = 0, qj = 0, qTj = 0;
    real lbj = 0, ubj = 0;   // stage-invariant bounds of component jz (per-instance with VARB)
#ifndef VAR_BOUNDS
    if (!VARB) {
        lbj = C->LB[jz];
        ubj = C->UB[jz];
    }
#endif

    for (;;) {
        // ---- refill: the first lane of a group without an instance pulls the next parked record
        if (!live && !drained) {
            long long slot = -1;
            if (lane == gl0) slot = wq.next();
            slot = __shfl_sync(gmask, slot, gl0);
            if (slot < 0) {
                drained = true;
                if (lane == gl0) marks.mark_drained();
            } else {
                const double *pk = io.park + slot;
                inst = __double_as_longlong(pk[0]);
                k = (int)__double_as_longlong(pk[1 * io.park_cap]);
                t = (real)pk[2 * io.park_cap];
                for (int e = j; e < N * n; e += GRP) {
                    Y[e] = (real)pk[(3 + e) * io.park_cap];
                    LAM[e] = (real)pk[(3 + N * n + e) * io.park_cap];
                }
                // b = -A x0, q = [Q xr; R ur], qT = T xr | xr        code_laxMPC_FISTA_C.c:275-289 (same operations as fista_kernel)
                real b = real(0);
#pragma unroll
                for (int i = 0; i < n; ++i) b = A::nmsub(b, C->AB[jx][i], (real)io.x0[inst * n + i]);
                bj = b;
                const real xrj = (real)io.xr[inst * n + jx];
                qj = is_x ? A::mul(C->Q[jx], xrj) : A::mul(C->R[jz - n], (real)io.ur[inst * m + (jz - n)]);
#if SPCIES_TERMINAL
                qTj = A::mul(C->T[jx], xrj);
#else
                qTj = xrj;
#endif
                if (VARB) {
                    lbj = (real)io.LB[inst * nm + jz];
                    ubj = (real)io.UB[inst * nm + jz];
                }
                live = true;
            }
        }
        __syncwarp();
        if (!__any_sync(FULL, live)) break;

        auto LBs = [&](int l) -> real {
#ifdef VAR_BOUNDS
            return VARB ? lbj : C->LB[l][jz];
#else
            (void)l; return lbj;
#endif
        };
        auto UBs = [&](int l) -> real {
#ifdef VAR_BOUNDS
            return VARB ? ubj : C->UB[l][jz];
#else
            (void)l; return ubj;
#endif
        };

        // ================= step 1: z_l(y), all stages                        :474-537 =================
        real zown[N - 1], zN, u0;
        {
            real ABcol[n];                              // column jz of [A B]
#pragma unroll
            for (int i = 0; i < n; ++i) ABcol[i] = C->AB[i][jz];
            real yv[n];
            ldv<n>(yv, Y);
            u0 = qj;                                    // lanes n..nm-1: u_0[j-n]
#pragma unroll
            for (int i = 0; i < n; ++i) u0 = A::nmsub(u0, ABcol[i], yv[i]);
#ifdef VAR_BOUNDS
            u0 = clip(A::mul(u0, qri), VARB ? lbj : C->LB0[is_x ? 0 : jz - n], VARB ? ubj : C->UB0[is_x ? 0 : jz - n]);
#else
            u0 = clip(A::mul(u0, qri), lbj, ubj);
#endif
#pragma unroll
            for (int l = 0; l < N - 1; ++l) {
                ldv<n>(yv, Y + (l + 1) * n);
                real zj = qj;
#pragma unroll
                for (int i = 0; i < n; ++i) zj = A::nmsub(zj, ABcol[i], yv[i]);
                if (is_x) zj = A::add(zj, Y[l * n + jx]);
                zj = clip(A::mul(zj, qri), LBs(l), UBs(l));
                zown[l] = zj;
                if (is_z) Z[l * nm + jz] = zj;
            }
#if SPCIES_TERMINAL
            zN = A::add(qTj, Y[(N - 1) * n + jx]);
#ifdef VAR_BOUNDS
            zN = clip(A::mul(zN, C->Ti[jx]), VARB ? lbj : C->LBN[jx], VARB ? ubj : C->UBN[jx]);
#else
            zN = clip(A::mul(zN, C->Ti[jx]), lbj, ubj);
#endif
#else
            zN = qTj;                                   // xr                     code_equMPC_FISTA_C.c:549
#endif
        }
        __syncwarp();

        // ================= step 2: r_l, all stages; exit test                :546-574, :337-361 =================
        bool over = false;
        {
            real ABrow[nm];                             // row jx of [A B]
#pragma unroll
            for (int i = 0; i < nm; ++i) ABrow[i] = C->AB[jx][i];
            real rj = A::add(bj, zown[0]);
#pragma unroll
            for (int i = 0; i < m; ++i) rj = A::nmsub(rj, ABrow[n + i], __shfl_sync(FULL, u0, gl0 + n + i));
            over |= exceeds(rj, tol_);
            if (is_x) R[jx] = rj;
#pragma unroll
            for (int l = 1; l < N; ++l) {
                real zp[nm];
                ldv<nm>(zp, Z + (l - 1) * nm);
                rj = (l < N - 1) ? zown[l < N - 1 ? l : 0] : zN;
#pragma unroll
                for (int i = 0; i < nm; ++i) rj = A::nmsub(rj, ABrow[i], zp[i]);
                over |= exceeds(rj, tol_);
                if (is_x) R[l * n + jx] = rj;
            }
        }
        const bool over_g = (__ballot_sync(FULL, over && is_x) & gmask) != 0u;
        if (live) {
            k += 1;
            int ef = 0;
            if (!over_g) ef = 1;
            else if (k >= k_max) ef = -1;
            if (ef != 0) {
                if (!is_x && is_z) io.u[inst * m + (jz - n)] = (double)u0;
                if (lane == gl0) {
                    io.k[inst] = k;
                    io.e[inst] = ef;
                    stat_k += (unsigned long long)k;
                    stat_nc += (ef < 0);
                }
                live = false;
            }
        }
        __syncwarp();

        // ================= step 3: s_l = Linv_l r_l, all stages =================
        real sw[N];                                     // s_l[j], then w_l[j], then d_lambda_l[j]
#pragma unroll
        for (int l = 0; l < N; ++l) {
            real rv[n], Lr[n];
            ldv<n>(rv, R + l * n);
            ldv<n>(Lr, &D->Linv[l][jx][0]);
            real s = A::mul(Lr[0], rv[0]);
#pragma unroll
            for (int i = 1; i < n; ++i) s = fma(Lr[i], rv[i], s);
            sw[l] = s;
        }

        // ================= step 4: forward chain mu_l = s_l - F_l mu_{l-1};  w_l = Uinv_l mu_l =================
        if (is_x) MU[jx] = sw[0];
        __syncwarp();
#pragma unroll
        for (int l = 1; l <= N; ++l) {
            real mp[n];
            ldv<n>(mp, MU + (l - 1) * n);               // mu_{l-1}, whole vector
            if (l < N) {
                real Fr[n];
                ldv<n>(Fr, &D->F[l][jx][0]);
                real mu = sw[l];
#pragma unroll
                for (int i = 0; i < n; ++i) mu = fma(-Fr[i], mp[i], mu);
                if (is_x) MU[l * n + jx] = mu;
                __syncwarp();
            }
            real Ur[n];
            ldv<n>(Ur, &D->Uinv[l - 1][jx][0]);
            real wv = A::mul(Ur[0], mp[0]);
#pragma unroll
            for (int i = 1; i < n; ++i) wv = fma(Ur[i], mp[i], wv);
            sw[l - 1] = wv;
        }

        // ================= step 5: backward chain dl_l = w_l - G_l dl_{l+1};  lambda, y updates    :368-385 =================
        const real t1 = t;
        t = A::mul(real(0.5), A::add(real(1), A::sqrt(A::add(real(1), A::mul(A::mul(real(4), t1), t1)))));
        const real coef = A::sub(t1, real(1));
        const real beta = A::div(coef, t);
        auto update = [&](int l) {                      // lambda_l = y_l + dl_l;  y_l <- lambda_l + beta (lambda_l - lambda1_l)
            const real lam = A::add(Y[l * n + jx], sw[l]);
            const real d = A::sub(lam, LAM[l * n + jx]);
            const real ynew = A::madd(lam, beta, d);
            if (is_x) {
                LAM[l * n + jx] = lam;
                Y[l * n + jx] = ynew;
            }
        };
        if (is_x) DL[(N - 1) * n + jx] = sw[N - 1];
        __syncwarp();
#pragma unroll
        for (int l = N - 2; l >= 0; --l) {
            real dn[n], Gr[n];
            ldv<n>(dn, DL + (l + 1) * n);
            ldv<n>(Gr, &D->G[l][jx][0]);
            real dl = sw[l];
#pragma unroll
            for (int i = 0; i < n; ++i) dl = fma(-Gr[i], dn[i], dl);
            sw[l] = dl;
            if (l > 0) {
                if (is_x) DL[l * n + jx] = dl;
                __syncwarp();
            }
            update(l + 1);
        }
        update(0);
        __syncwarp();
    }
    flush_stats(io.queue, stat_k, stat_nc);
    marks.mark_end();
}
