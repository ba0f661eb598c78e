// MPC_FISTA_single.cuh -- latency engine of the laxMPC / equMPC FISTA solvers: ONE CTA PER INSTANCE (included by MPC_FISTA.cuh,
// inside spcies::fista).
//
// The reference's single-instance symbol is a batch of one; the throughput engines give such a call one warp (8 rows of the MMA,
// 7 of them idle) and ~1.6 us per iteration of sequential block recurrences.  Here a whole CTA works on one instance, and the
// iteration (code_laxMPC_FISTA_C.c:323-389) is restated in the space of the primal variable so that it is ONE data-parallel phase
// and one barrier.  With E = the equality-constraint matrix in the reference's sign convention (row block l: +x_{l+1} -
// [A B] (x_l, u_l)), z = clip(Hd o (q + E' y)) (:471-539), r = b + E z (:546-574), lambda+ = y + W^-1 r (:577-651, :364-369),
// y+ = lambda+ + beta_k (lambda+ - lambda) (:372-385), the images v = E' y and mu = E' lambda obey
//     mu+ = v + g + P z,   v+ = mu+ + beta_k (mu+ - mu),   z+ = clip(Hd o (q + v+)),      P = E' W^-1 E,  g = E' W^-1 b,
// i.e. per iteration one dense |z| x |z| product with the shared vector z followed by component-wise work.  A pair of threads owns
// two rows of P: each has both rows' coefficients for half of the columns in registers, reads its half of z from shared memory
// (128-bit loads, two addresses per warp) and exchanges one partial sum with its partner, then updates its own row.  The shared-
// memory / shuffle pipe is what bounds a one-CTA iteration: a load fills 4 bytes per lane and cycle whatever its width and
// uniformity (measured per iteration: four threads per row with two butterfly steps 1080 cycles, one thread per row reading all
// of z 800, this scheme 745; tools/probes/lat_probe.cu, tools/latency_probe.py).  The residual rows (<= nm + 1 terms) run in two
// more warps, and one __syncthreads_or carries the exit decision (:337-361) and the new z (double buffered).  P and g are formed on the host in extended precision
// from Alpha / Beta; beta_k comes from a table (the t-sequence does not depend on the instance).
// FAST arithmetic only (the sums run in a different order than the reference's); used for host-buffer batches
// of at most 64 instances (spcies_host.cuh: run_small), one CTA each; the inputs of up to 4 instances travel as kernel arguments
// (no PCIe read before the first iteration), the results and a completion flag go to mapped host memory.
#pragma once

#ifndef SPCIES_FISTA_SINGLE
#define SPCIES_FISTA_SINGLE 1
#endif

constexpr int SG_ROWS = N * n;                                   // dual variables
constexpr int SG_ZLEN = TERMINAL ? N * nm : N * nm - n;          // primal variables
constexpr int SG_ZPAD = (SG_ZLEN + 7) / 8 * 8;
constexpr int SG_H = SG_ZPAD / 2;                                 // columns per thread: two rows x half the columns
constexpr int SG_MR = (SG_ZLEN + 31) / 32 * 32;                  // threads [0, SG_MR): one row of P each
constexpr int SG_RR = (SG_ROWS + 31) / 32 * 32;                  // threads [SG_MR, SG_MR + SG_RR): one residual row each
constexpr int SG_BLOCK = SG_MR + SG_RR;
constexpr int SG_NZ2 = nm + 1;
constexpr int SG_KTAB = k_max + 2;
constexpr int SG_NARG = 4;                                       // instances whose inputs fit the kernel arguments
constexpr bool HAS_SINGLE = SPCIES_FISTA_SINGLE != 0 && sizeof(real) == 8 && SG_BLOCK <= 1024 && SG_ZPAD <= 96 && SG_MR % 2 == 0 && SG_KTAB <= 16384;

struct SingleArgs {
    int count;                                                   // 0: read the inputs through BatchIO
    double x0[SG_NARG][n], xr[SG_NARG][n], ur[SG_NARG][m];
};

struct alignas(16) SingleTables {
    // [term][thread]: coalesced loads into registers at the start of the kernel
    double P[SG_ZPAD][SG_MR];              // P[col][row]; thread e works on rows (e & ~1, e | 1), columns [(e & 1) SG_H, .. + SG_H)
    double Hd[SG_MR], cq[SG_MR];           // q_e = cq_e ref[qsrc_e], ref = (xr, ur)   (Q, R, T stored negated)
    double LB[SG_MR], UB[SG_MR];
    double cg[2 * n][SG_MR];               // g_e = sum_j cg[j] x0_j + sum_j cg[n + j] xr_j
    double c2[SG_NZ2][SG_RR];              // thread SG_MR + i: r_i = b_i + sum_t c2[t][i] z[i2[t][i]]
    double cb[n][SG_RR];                   // b_i = sum_j cb[j] x0_j (rows of the first block: -A) [+ xr_{bx} for equMPC's last block]
    double beta[SG_KTAB];                  // momentum coefficient (t_{k-1} - 1) / t_k of iteration k   (:374-381)
    int i2[SG_NZ2][SG_RR];
    int qsrc[SG_MR], comp[SG_MR];          // component of (x, u) a primal variable is (per-instance bounds)
    int bx[SG_RR];                         // equMPC: component of xr added to row i (last block), else -1
};
constexpr size_t SINGLE_BYTES = (sizeof(SingleTables) + 15) / 16 * 16;
constexpr size_t SINGLE_OFFSET = HAS_MMA ? TOTAL_BLOB_BYTES : BLOB_BYTES;
constexpr size_t SINGLE_SMEM = (size_t)(2 * SG_ZPAD + 2 * nm) * sizeof(double) + 16;

// ---- the iteration in the space of the primal variable, as dense host-side matrices (extended precision) -------------------------
// Shared by the latency engine below and by the dense tensor-core policy (MPC_FISTA_dense.cuh).
struct PrimalForm {
    typedef long double ld;
    static constexpr int ROWS = SG_ROWS, ZLEN = SG_ZLEN;
    ld *E, *P, *G0, *G1;            // E [ROWS][ZLEN]; P = E' W^-1 E [ZLEN][ZLEN]; g = G0 x0 + G1 xr, [ZLEN][n] each (G1 = 0 for laxMPC)
    ld (*bA)[n];                    // b = bA x0 [+ xr in the last block for equMPC], [ROWS][n]
    double Hd[ZLEN], cq[ZLEN], LB[ZLEN], UB[ZLEN];     // z_e = clip(Hd_e (q_e + v_e)), q_e = cq_e ref[qsrc_e], ref = (xr, ur)
    int qsrc[ZLEN], comp[ZLEN];     // comp: the component of (x, u) a primal variable is (per-instance bounds)
    double beta[SG_KTAB];           // momentum coefficient (t_{k-1} - 1) / t_k of iteration k   (:374-381)
    // primal variable e: u_0[j] | z[l][j] (x_{l+1}, u_{l+1}) | z_N[j]
    static int el_u0(int j) { return j; }
    static int el_z(int l, int j) { return m + l * nm + j; }
    static int el_zN(int j) { return m + (N - 1) * nm + j; }
    void primal(int e, double hd, double cq_, int src, int cmp, double lb, double ub) {
        Hd[e] = hd;
        cq[e] = cq_;
        qsrc[e] = src;
        comp[e] = cmp;
        LB[e] = lb;
        UB[e] = ub;
    }
    explicit PrimalForm(const spcies_consts &C) {
        for (int j = 0; j < m; ++j)                                               // :478-495
#ifdef VAR_BOUNDS
            primal(el_u0(j), (double)C.QRi[n + j], (double)C.R[j], n + j, n + j, (double)C.LB0[j], (double)C.UB0[j]);
#else
            primal(el_u0(j), (double)C.QRi[n + j], (double)C.R[j], n + j, n + j, (double)C.LB[n + j], (double)C.UB[n + j]);
#endif
        for (int l = 0; l < N - 1; ++l)                                           // :498-521
            for (int j = 0; j < nm; ++j)
#ifdef VAR_BOUNDS
                primal(el_z(l, j), (double)C.QRi[j], j < n ? (double)C.Q[j] : (double)C.R[j - n], j, j, (double)C.LB[l][j], (double)C.UB[l][j]);
#else
                primal(el_z(l, j), (double)C.QRi[j], j < n ? (double)C.Q[j] : (double)C.R[j - n], j, j, (double)C.LB[j], (double)C.UB[j]);
#endif
#if SPCIES_TERMINAL
        for (int j = 0; j < n; ++j)                                               // :524-537
#ifdef VAR_BOUNDS
            primal(el_zN(j), (double)C.Ti[j], (double)C.T[j], j, j, (double)C.LBN[j], (double)C.UBN[j]);
#else
            primal(el_zN(j), (double)C.Ti[j], (double)C.T[j], j, j, (double)C.LB[j], (double)C.UB[j]);
#endif
#endif
        // E (rows of the residual, :546-574) as a dense matrix; b = -A x0 (:275-280)
        E = new ld[(size_t)ROWS * ZLEN]();
        bA = new ld[ROWS][n]();
        for (int l = 0; l < N; ++l)
            for (int j = 0; j < n; ++j) {
                ld *row = E + (size_t)(l * n + j) * ZLEN;
                if (l == 0) {
                    for (int c = 0; c < m; ++c) row[el_u0(c)] = -(ld)C.AB[j][n + c];
                    for (int c = 0; c < n; ++c) bA[j][c] = -(ld)C.AB[j][c];
                } else {
                    for (int c = 0; c < nm; ++c) row[el_z(l - 1, c)] = -(ld)C.AB[j][c];
                }
                if (l < N - 1) row[el_z(l, j)] = 1;
                else if (TERMINAL) row[el_zN(j)] = 1;
            }
        // W^-1: the reference's solve (forward / backward substitution with Alpha, Beta) applied to the unit vectors   :577-651
        ld *Wi = new ld[(size_t)ROWS * ROWS]();
        for (int c = 0; c < ROWS; ++c) {
            ld mu[N][n] = {};
            mu[c / n][c % n] = 1;
            for (int l = 0; l < N; ++l)                                            // forward substitution
                for (int j = 0; j < n; ++j) {
                    if (l > 0)
                        for (int i = 0; i < n; ++i) mu[l][j] -= (ld)C.Alpha[l - 1][i][j] * mu[l - 1][i];
                    for (int i = 0; i < j; ++i) mu[l][j] -= (ld)C.Beta[l][i][j] * mu[l][i];
                    mu[l][j] *= (ld)C.Beta[l][j][j];
                }
            for (int l = N - 1; l >= 0; --l)                                       // backward substitution
                for (int j = n - 1; j >= 0; --j) {
                    if (l < N - 1)
                        for (int i = 0; i < n; ++i) mu[l][j] -= (ld)C.Alpha[l][j][i] * mu[l + 1][i];
                    for (int i = n - 1; i > j; --i) mu[l][j] -= (ld)C.Beta[l][j][i] * mu[l][i];
                    mu[l][j] *= (ld)C.Beta[l][j][j];
                }
            for (int r = 0; r < ROWS; ++r) Wi[(size_t)r * ROWS + c] = mu[r / n][r % n];
        }
        ld *M = new ld[(size_t)ROWS * ZLEN]();                                     // W^-1 E
        for (int i = 0; i < ROWS; ++i)
            for (int c = 0; c < ZLEN; ++c) {
                ld v = 0;
                for (int r = 0; r < ROWS; ++r) v += Wi[(size_t)i * ROWS + r] * E[(size_t)r * ZLEN + c];
                M[(size_t)i * ZLEN + c] = v;
            }
        P = new ld[(size_t)ZLEN * ZLEN]();
        G0 = new ld[(size_t)ZLEN * n]();
        G1 = new ld[(size_t)ZLEN * n]();
        for (int e = 0; e < ZLEN; ++e) {
            for (int c = 0; c < n; ++c) {                                          // g = E' W^-1 b, b = bA x0 [+ xr in the last block]
                ld v0 = 0, v1 = 0;
                for (int i = 0; i < ROWS; ++i) {
                    ld wb = 0;
                    for (int r = 0; r < n; ++r) wb += Wi[(size_t)i * ROWS + r] * bA[r][c];
                    v0 += E[(size_t)i * ZLEN + e] * wb;
                    if (!TERMINAL) v1 += E[(size_t)i * ZLEN + e] * Wi[(size_t)i * ROWS + (N - 1) * n + c];
                }
                G0[(size_t)e * n + c] = v0;
                G1[(size_t)e * n + c] = v1;
            }
            for (int col = 0; col < ZLEN; ++col) {
                ld v = 0;
                for (int i = 0; i < ROWS; ++i) v += E[(size_t)i * ZLEN + e] * M[(size_t)i * ZLEN + col];
                P[(size_t)e * ZLEN + col] = v;
            }
        }
        delete[] M;
        delete[] Wi;
        // momentum coefficients: t_0 = 1, t_k = (1 + sqrt(1 + 4 t_{k-1}^2)) / 2, beta_k = (t_{k-1} - 1) / t_k, in double like the reference
        double t = 1.0;
        beta[0] = 0.0;
        for (int k = 1; k < SG_KTAB; ++k) {
            const double t1 = t;
            t = 0.5 * (1.0 + sqrt(1.0 + 4.0 * t1 * t1));
            beta[k] = (t1 - 1.0) / t;
        }
    }
    ~PrimalForm() {
        delete[] E;
        delete[] P;
        delete[] G0;
        delete[] G1;
        delete[] bA;
    }
    PrimalForm(const PrimalForm &) = delete;
    PrimalForm &operator=(const PrimalForm &) = delete;
};

#if SPCIES_FISTA_SINGLE
static inline void fill_single_tables(const PrimalForm &F, SingleTables &T) {
    memset(&T, 0, sizeof T);
    for (int t = 0; t < SG_MR; ++t) {
        T.LB[t] = -1e300;
        T.UB[t] = 1e300;
    }
    for (int t = 0; t < SG_RR; ++t) T.bx[t] = -1;
    for (int e = 0; e < SG_ZLEN; ++e) {
        T.Hd[e] = F.Hd[e];
        T.cq[e] = F.cq[e];
        T.qsrc[e] = F.qsrc[e];
        T.comp[e] = F.comp[e];
        T.LB[e] = F.LB[e];
        T.UB[e] = F.UB[e];
        for (int c = 0; c < n; ++c) {
            T.cg[c][e] = (double)F.G0[(size_t)e * n + c];
            T.cg[n + c][e] = (double)F.G1[(size_t)e * n + c];
        }
        for (int col = 0; col < SG_ZLEN; ++col) T.P[col][e] = (double)F.P[(size_t)e * SG_ZLEN + col];
    }
    for (int i = 0; i < SG_ROWS; ++i) {
        int t = 0;
        for (int c = 0; c < SG_ZLEN; ++c) {
            const long double v = F.E[(size_t)i * SG_ZLEN + c];
            if (v == 0) continue;
            T.c2[t][i] = (double)v;
            T.i2[t][i] = c;
            ++t;
        }
        for (int c = 0; c < n; ++c) T.cb[c][i] = (double)F.bA[i][c];
        if (!TERMINAL && i >= (N - 1) * n) T.bx[i] = i - (N - 1) * n;          // equMPC: x_N = xr   (code_equMPC_FISTA_C.c:549)
    }
    memcpy(T.beta, F.beta, sizeof T.beta);
}

// ---- per-thread coefficients (registers), loaded once per kernel
struct SingleLane {
    double w[SG_ZPAD];          // mat: half rows of the pair (w[j], w[SG_H + j]);  else: the first SG_NZ2 entries hold the residual row's terms
    int i2[SG_NZ2];
    double hd, lb, ub;
    int e;
    bool mat, owner;            // mat is warp-uniform: rows of P | rows of the residual
    __device__ __forceinline__ void load(const SingleTables *T, int tid) {
        mat = tid < SG_MR;
        e = mat ? tid : tid - SG_MR;
        owner = mat && e < SG_ZLEN;
        hd = lb = ub = 0.0;
        if (mat) {
#pragma unroll
            for (int j = 0; j < SG_H; ++j) {
                w[j] = T->P[(e & 1) * SG_H + j][e & ~1];
                w[SG_H + j] = T->P[(e & 1) * SG_H + j][e | 1];
            }
            hd = T->Hd[e];
            lb = T->LB[e];
            ub = T->UB[e];
#pragma unroll
            for (int t = 0; t < SG_NZ2; ++t) i2[t] = 0;
        } else {
#pragma unroll
            for (int t = 0; t < SG_NZ2; ++t) {
                w[t] = T->c2[t][e];
                i2[t] = T->i2[t][e];
            }
        }
    }
};

// One instance, the whole CTA.  s_ref = (xr, ur, x0) in scaled units (visible to every thread); returns k, e_flag and z of the last pass.
__device__ __forceinline__ void single_solve(const SingleTables *T, const SingleLane &L, double lb, double ub, double *s_z,
                                             const double *s_ref, int &k_out, int &ef_out, const double *&z_out) {
    const int e = L.e;
    double qe = 0.0, ce = 0.0;                                       // mat: q_e, g_e;  else: b_i in ce
    if (L.mat) {
        qe = T->cq[e] * s_ref[T->qsrc[e]];
#pragma unroll
        for (int j = 0; j < n; ++j) {
            ce = fma(T->cg[j][e], s_ref[nm + j], ce);
            ce = fma(T->cg[n + j][e], s_ref[j], ce);
        }
    } else {
#pragma unroll
        for (int j = 0; j < n; ++j) ce = fma(T->cb[j][e], s_ref[nm + j], ce);
        if (T->bx[e] >= 0) ce += s_ref[T->bx[e]];
    }
    const double tol_ = (double)tol;
    double mu = 0.0, v = 0.0;
    if (L.owner) s_z[e] = clip(qe * L.hd, lb, ub);                   // z(lambda = 0) of the initial step   (:300-303)
    __syncthreads();
    int k = 0, ef = 0;
    const double *zc = s_z;                                          // z of this pass; the other buffer receives the next one
    for (;;) {                                                       // pass 0 = the initial step (:300-320): no exit test, y = lambda
        double *zn = s_z + ((k & 1) ? 0 : SG_ZPAD);
        bool over = false;
        if (L.mat) {
            const double bk = T->beta[k];
            // rows (e & ~1, e | 1) x this thread's half of the columns; the partner thread has the other half
            const double2 *z2 = reinterpret_cast<const double2 *>(zc + (e & 1) * SG_H);
            double a[4] = {0.0, 0.0, 0.0, 0.0}, b[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 0; j < SG_H; j += 4) {
                const double2 p0 = z2[j / 2], p1 = z2[j / 2 + 1];
                a[0] = fma(L.w[j], p0.x, a[0]);
                a[1] = fma(L.w[j + 1], p0.y, a[1]);
                a[2] = fma(L.w[j + 2], p1.x, a[2]);
                a[3] = fma(L.w[j + 3], p1.y, a[3]);
                b[0] = fma(L.w[SG_H + j], p0.x, b[0]);
                b[1] = fma(L.w[SG_H + j + 1], p0.y, b[1]);
                b[2] = fma(L.w[SG_H + j + 2], p1.x, b[2]);
                b[3] = fma(L.w[SG_H + j + 3], p1.y, b[3]);
            }
            double d0 = (a[0] + a[1]) + (a[2] + a[3]), d1 = (b[0] + b[1]) + (b[2] + b[3]);
            // each thread keeps the sum of its own row (row e): send the partner's partial sum, receive mine
            const double mine = (e & 1) ? d1 : d0, theirs = (e & 1) ? d0 : d1;
            const double d = ce + mine + __shfl_xor_sync(0xffffffffu, theirs, 1);
            const double mnew = v + d;                               // E' (y + W^-1 r)
            v = k == 0 ? mnew : fma(bk, mnew - mu, mnew);            // :311-320 | :372-385 (beta_1 = 0)
            mu = mnew;
            if (L.owner) zn[e] = clip((qe + v) * L.hd, lb, ub);      // (unused when this pass turns out to be the last)
        } else {
            double r0 = ce, r1 = 0.0;
#pragma unroll
            for (int t = 0; t + 1 < SG_NZ2; t += 2) {
                r0 = fma(L.w[t], zc[L.i2[t]], r0);
                r1 = fma(L.w[t + 1], zc[L.i2[t + 1]], r1);
            }
            if (SG_NZ2 & 1) r0 = fma(L.w[SG_NZ2 - 1], zc[L.i2[SG_NZ2 - 1]], r0);
            over = e < SG_ROWS && fabs(r0 + r1) > tol_;
        }
        const bool any_over = __syncthreads_or(over);
        if (k > 0) {
            ef = !any_over ? 1 : (k >= k_max ? -1 : 0);              // :350-361
            if (ef != 0) break;
        }
        k += 1;
        zc = zn;
    }
    k_out = k;
    ef_out = ef;
    z_out = zc;
}

template <bool VARB>
__global__ void __launch_bounds__(SG_BLOCK, 1) fista_single_kernel(const BatchIO io, const unsigned char *__restrict__ g_blob,
                                                                   const __grid_constant__ SingleArgs args) {
    const spcies_consts *C = reinterpret_cast<const spcies_consts *>(g_blob);      // engineering-unit scaling only
    (void)C;
    const SingleTables *T = reinterpret_cast<const SingleTables *>(g_blob + SINGLE_OFFSET);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_z = reinterpret_cast<double *>(smem_raw);             // [2][SG_ZPAD], zero padded
    double *s_ref = s_z + 2 * SG_ZPAD;                              // xr, ur, x0
    const int tid = threadIdx.x;
    const long long inst = blockIdx.x;
    if (inst >= io.B) return;
    SingleLane L;
    L.load(T, tid);
    double lb = L.lb, ub = L.ub;
    if (VARB && L.owner) {
        lb = io.LB[inst * nm + T->comp[L.e]];
        ub = io.UB[inst * nm + T->comp[L.e]];
    }
    // ---- the instance                                                       code_laxMPC_FISTA_C.c:255-289
    {
        const bool in_args = args.count > 0;
        const double *px0 = in_args ? &args.x0[0][0] : io.x0, *pxr = in_args ? &args.xr[0][0] : io.xr, *pur = in_args ? &args.ur[0][0] : io.ur;
        if (tid < n) {
            s_ref[tid] = eng_x(C, pxr, inst, n, tid);
            s_ref[nm + tid] = eng_x(C, px0, inst, n, tid);
        } else if (tid < nm) {
            s_ref[tid] = eng_u(C, pur, inst, m, tid - n);
        }
    }
    for (int j = SG_ZLEN + tid; j < SG_ZPAD; j += SG_BLOCK) s_z[j] = s_z[SG_ZPAD + j] = 0.0;
    __syncthreads();
    int k, ef;
    const double *zc;
    single_solve(T, L, lb, ub, s_z, s_ref, k, ef, zc);
    // ---- results                                                            :392-407
    if (tid < m) io.u[inst * m + tid] = eng_u_out(C, zc[tid], tid);
    if (tid == 0) {
        io.k[inst] = k;
        io.e[inst] = ef;
    }
    if (io.done != nullptr) {                 // completion flag in mapped host memory: the host spins on it instead of synchronising
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int *>(io.done + inst) = io.done_seq;
        }
    }
}

// ---- lingering server for the reference's single-instance symbol --------------------------------------------------------------
// A launch costs ~10 us end to end, more than the solve.  The first single-instance call starts this one-CTA kernel; it keeps its
// coefficients in registers and polls a mailbox in mapped host memory (spcies_common.cuh: SingleMailbox): the host posts (x0, xr,
// ur) and a sequence number, the kernel answers with (u_opt, k, e_flag) and the same number.  It exits by itself `linger_ns` after
// the last request (or when told to), so that nothing waits long on a device-wide synchronisation; the host restarts it on demand
// (spcies_host.cuh: run_server).
typedef SingleMailbox<n, m> Mailbox;
__global__ void __launch_bounds__(SG_BLOCK, 1) fista_server_kernel(const unsigned char *__restrict__ g_blob, Mailbox *mb, unsigned int seq0,
                                                                   unsigned long long linger_ns) {
    const spcies_consts *C = reinterpret_cast<const spcies_consts *>(g_blob);
    (void)C;
    const SingleTables *T = reinterpret_cast<const SingleTables *>(g_blob + SINGLE_OFFSET);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_z = reinterpret_cast<double *>(smem_raw);
    double *s_ref = s_z + 2 * SG_ZPAD;
    __shared__ __align__(16) unsigned int s_in[Mailbox::NL * 16];
    __shared__ int s_cmd;                                            // 0: nothing yet, 1: request, 2: exit
    const int tid = threadIdx.x;
    SingleLane L;
    L.load(T, tid);
    for (int j = SG_ZLEN + tid; j < SG_ZPAD; j += SG_BLOCK) s_z[j] = s_z[SG_ZPAD + j] = 0.0;
    unsigned int last = seq0;
    unsigned long long t_idle = globaltimer_ns();
    for (;;) {
        if (tid < 32) {
            const volatile unsigned int *src = reinterpret_cast<const volatile unsigned int *>(mb->in);
            for (int wd = tid; wd < Mailbox::NL * 16; wd += 32) s_in[wd] = src[wd];       // one PCIe read per 64-byte line
            __syncwarp();
            if (tid == 0) {
                bool req = true;
                for (int l = 0; l < Mailbox::NL; ++l) req = req && s_in[l * 16 + 14] == last + 1;
                int cmd = req ? 1 : 0;
                if (!req && (s_in[15] != 0 || globaltimer_ns() - t_idle > linger_ns)) cmd = 2;
                s_cmd = cmd;
            }
        }
        __syncthreads();
        const int cmd = s_cmd;
        if (cmd == 2) break;
        if (cmd == 1) {
            // payload: x0[n], xr[n], ur[m], seven doubles per line                  code_laxMPC_FISTA_C.c:255-289
            const double *raw = reinterpret_cast<const double *>(s_in);
            auto pay = [&](int j) { return raw[(j / 7) * 8 + j % 7]; };
            if (tid < n) {
                const double x0v = pay(tid), xrv = pay(n + tid);
                s_ref[tid] = eng_x_value(C, xrv, tid);
                s_ref[nm + tid] = eng_x_value(C, x0v, tid);
            } else if (tid < nm) {
                s_ref[tid] = eng_u_value(C, pay(2 * n + tid - n), tid - n);
            }
            __syncthreads();
            int k, ef;
            const double *zc;
            single_solve(T, L, L.lb, L.ub, s_z, s_ref, k, ef, zc);
            if (tid == 0) {
                for (int j = 0; j < m; ++j) mb->out.u[j] = eng_u_out(C, zc[j], j);
                mb->out.k = k;
                mb->out.e = ef;
                __threadfence_system();
                *reinterpret_cast<volatile unsigned int *>(&mb->out.seq) = last + 1;
                t_idle = globaltimer_ns();
            }
            last += 1;
        }
        __syncthreads();                                             // s_in / s_cmd are rewritten by the next poll
    }
    if (tid == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int *>(&mb->alive) = 0u;
    }
}
#endif
