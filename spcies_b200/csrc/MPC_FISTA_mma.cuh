// MPC_FISTA_mma.cuh -- tensor-core (DMMA) engine of the FISTA solver (included by MPC_FISTA.cuh, inside spcies::fista).
//
// Why.  ncu on the one-thread-per-instance kernel (profiles/r1_fista_v3_ncu_summary.txt) shows the batched
// shared-matrix block products as the bottleneck: every DFMA of the W-solve takes a warp-uniform constant that costs
// one shared-memory wavefront per 8 bytes (shared pipe 69 % busy, FP64 pipe 55 %), the per-instance iterates (200
// doubles) do not fit the register file, and the ~5200-instruction iteration of one thread makes the tail of the
// slowest instances latency bound (5.6 ms of a 13.9 ms step).  Every product of the FAST arithmetic is a dense
// (<= 8 x 8) shared matrix times a per-instance vector, i.e. one small GEMM per warp with the batch as the M dimension.
//
// Mapping.  A warp owns 8 instances (one per row of the m8n8k4 FP64 MMA, SASS DMMA.8x8x4); the 4 lanes of a lane group
// hold the instance's vectors, 2 components per lane: lane (g = lane / 4, t = lane % 4) holds components 2t and 2t+1 of
// every vector of instance g.  That is exactly the C/D fragment layout, and -- with the shared matrix's columns split
// even / odd over two MMAs -- also the A fragment layout, so products chain with no shuffle:
//
//     v' = c + M v :   D = A0 * B0 + C,  D = A1 * B1 + D      A0[g][t] = v_g[2t], A1[g][t] = v_g[2t+1]   (registers of v)
//                                                              B0[t][o] = M[o][2t], B1[t][o] = M[o][2t+1] (lane reads the
//                                                              double2 at index `lane` of the row-major 8 x 8 M)
//
// All iterates (y, lambda, w: 3 N vectors = 60 doubles per lane at N = 10) and the [A B] fragments live in registers; the
// 4 N stage matrices (Linv, -F, Uinv, -G) are read from shared memory as one conflict-free LDS.128 per lane and matrix
// (4 wavefronts per 8 instances instead of 36 per 32).  116 DMMA per warp iteration (N = 10); the component-wise work
// (scaling, clipping, exit test, momentum) is 2 components per lane.  The t-sequence of FISTA does not depend on the
// instance, so the momentum coefficient (t_{k-1} - 1) / t_k is a table indexed by k (no sqrt / divide in the loop).
//
// Arithmetic: the FAST arithmetic of fista_kernel (explicit block inverses, FMA) with the dot products accumulated in
// the MMA's order; results differ from the reference by rounding only (same gate: u_opt <= 1e-9 relative, e_flag
// identical, |dk| <= 1).  EXACT mode, float precision, the debug payload and problems with nn_ + mm_ > 8 use the
// one-thread-per-instance kernel.  Refill is per lane group; park & resume (io.phase) use the same record format.
#pragma once
// (spcies_mma.cuh is included by the parent header, outside its namespace)

#ifndef SPCIES_FISTA_MMA
#define SPCIES_FISTA_MMA 1           // 0: never use the tensor-core engine
#endif
#ifndef SPCIES_FISTA_MMA_PRESCALE
#define SPCIES_FISTA_MMA_PRESCALE 1  // 1: fold the QRi / Ti scalings of z into [A B]' and q (one FP64 operation less per component of z;
                                     // the constants are rounded once more: same measured accuracy, tools/diag_equ.py)
#endif
#ifndef SPCIES_FISTA_MMA_INT_CLIP
#define SPCIES_FISTA_MMA_INT_CLIP 0      // clip with integer compares on order-preserving keys of the doubles (measured)
#endif
#ifndef SPCIES_FISTA_MMA_INT_TEST
#define SPCIES_FISTA_MMA_INT_TEST 0      // exit test with integer compares (tools/engine_variants.py: measured)
#endif
#ifndef SPCIES_FISTA_MMA_MERGE
#define SPCIES_FISTA_MMA_MERGE 1     // 1: permuted component layout that lets two n-vector products share a k-step (5 <= nn_ <= 6)
#endif
#ifndef SPCIES_FISTA_MMA_BLOCK
#define SPCIES_FISTA_MMA_BLOCK 256   // threads per CTA: 2 warps per SM scheduler, up to 255 registers (the low-latency configuration)
#endif
#ifndef SPCIES_FISTA_MMA_BLOCK_BULK
#define SPCIES_FISTA_MMA_BLOCK_BULK 384   // first launch of a large batch: 3 warps per scheduler at 168 registers (a few spills) keep the
                                          // FP64 pipe busier; resumed rounds and small batches are latency bound and use MMA_BLOCK
#endif

constexpr int MMA_BLOCK = SPCIES_FISTA_MMA_BLOCK;
constexpr int MMA_BLOCK_BULK = SPCIES_FISTA_MMA_BLOCK_BULK;
constexpr long long MMA_BULK_MIN = 16LL * 148 * (MMA_BLOCK_BULK / 4);   // batches of at least 16 waves use the bulk configuration
constexpr bool MMA_SHAPE_OK = SPCIES_FISTA_MMA != 0 && sizeof(real) == 8 && nm <= 8 && N >= 2 && N <= 12;
constexpr int MMA_KTAB = k_max + 2;
constexpr bool PRESCALE = SPCIES_FISTA_MMA_PRESCALE != 0;

// Component layout.  A vector of an instance is 8 "columns" (column c lives in lane t = c / 2 of the group, register c % 2).
// An MMA k-step consumes one register of the 4 lanes, i.e. columns {0,2,4,6} or {1,3,5,7}.
//   plain   column = component: x_e -> e, u_j -> n + j.  Every product with an n-vector (n = 6) wastes a quarter of both k-steps.
//   MERGE   x_0..x_3 -> 0,2,4,6 (a full k-step), x_4.. -> 1,3 (half a k-step), u_j -> the columns after them (5,7).  The
//           vectors of the two recurrences (mu, d_lambda) carry a second copy of x_4.. in columns 5,7 -- free, the matrix
//           rows are duplicated -- so the half k-steps of two products that add up (Linv_l r_l + (-F_l) mu_{l-1}, and
//           Uinv_l mu_l + (-G_l) d_lambda_{l+1}) merge into one:  a = lane < 2 ? r.reg1 : mu.reg1.  3 MMAs instead of 4, twice
//           per stage: 98 instead of 116 DMMA per iteration at N = 10.
constexpr bool MERGE = SPCIES_FISTA_MMA_MERGE != 0 && PRESCALE && n >= 5 && n <= 6 && nm <= 8;
__host__ __device__ constexpr int col_x(int e) { return MERGE ? (e < 4 ? 2 * e : 2 * (e - 4) + 1) : e; }
__host__ __device__ constexpr int col_u(int j) { return MERGE ? 2 * (n - 4 + j) + 1 : n + j; }
__host__ __device__ constexpr int col_dup(int e) { return 2 * (e - 4 + 2) + 1; }   // MERGE: second copy of x_e, e >= 4
__host__ __device__ constexpr int x_at(int c) {
    for (int e = 0; e < n; ++e)
        if (col_x(e) == c) return e;
    return -1;
}
__host__ __device__ constexpr int u_at(int c) {
    for (int j = 0; j < m; ++j)
        if (col_u(j) == c) return j;
    return -1;
}
__host__ __device__ constexpr int z_at(int c) { return x_at(c) >= 0 ? x_at(c) : (u_at(c) >= 0 ? n + u_at(c) : -1); }   // component of z = (x, u)
// component of mu / d_lambda produced in output column c (MERGE: including the second copies)
__host__ __device__ constexpr int xo_at(int c) {
    if (x_at(c) >= 0) return x_at(c);
    if (MERGE)
        for (int e = 4; e < n; ++e)
            if (col_dup(e) == c) return e;
    return -1;
}

// order-preserving integer key of a double (sign-magnitude -> two's complement): a > b  <=>  key(a) > key(b) for non-NaN values
__device__ __forceinline__ long long dkey(double x) {
    const long long b = __double_as_longlong(x);
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
// clip() of spcies_common.cuh with the two compares on the integer pipe
__device__ __forceinline__ double clip_mma(double v, double lo, double hi) {
#if SPCIES_FISTA_MMA_INT_CLIP
    const long long kv = dkey(v), kl = dkey(lo), kh = dkey(hi);
    const bool above = kv > kl;
    const double t = above ? v : lo;
    const long long kt = above ? kv : kl;
    return kt > kh ? hi : t;
#else
    return clip(v, lo, hi);
#endif
}

struct alignas(16) MmaTables {
    // every table is indexed by column (see the layout above); matrices are row-major [output column][input column]
    double SNABt[64];         // -[A B]'  (out z, in x)   [PRESCALE: -diag(QRi) [A B]', the scaling of z folded into the product]
    double NAB[64];           // -[A B]   (out x, in z)
    // plain layout: one 8 x 8 matrix per product and stage
    double Linv[MERGE ? 1 : N][64], NF[MERGE ? 1 : N][64], Uinv[MERGE ? 1 : N][64], NG[MERGE ? 1 : N][64];
    // MERGE layout, per stage and lane (o = lane / 4 output column, t = lane % 4):
    //   FWa = (Linv[o][x_t], -F[o][x_t]),  FWb = t < 2 ? Linv[o][x_{4+t}] : -F[o][x_{4+t-2}]      (F = 0 at stage 0)
    //   BWa = (Uinv[o][x_t], -G[o][x_t]),  BWb = t < 2 ? Uinv[o][x_{4+t}] : -G[o][x_{4+t-2}]      (G = 0 at stage N-1)
    double2 FWa[MERGE ? N : 1][32], BWa[MERGE ? N : 1][32];
    double FWb[MERGE ? N : 1][32], BWb[MERGE ? N : 1][32];
    double QRiz[8], QRiy[8];  // scaling of z by column: all of z  |  state columns only (multiplies y_l)
    double Ti[8];             // component scalings are stored negated by the generator, like Q, R, T
    double Qs[8], Ts[8];      // q = Qs o [xr; ur],  qT = Ts o xr (lax) | xr (equ)   [PRESCALE: times QRi, Ti]
    double LBs[N + 1][8];     // row 0: u_0, rows 1..N-1: stage l = row - 1, row N: terminal state; -1e300 / 1e300 where unused
    double UBs[N + 1][8];
    int xat[8], uat[8];       // component of x / u held by a column, or -1
    double beta[MMA_KTAB];    // momentum coefficient of the pass that follows the k-th exit test (0 for k = 0, 1)
};
constexpr size_t MMA_BYTES = (sizeof(MmaTables) + 15) / 16 * 16;
constexpr size_t MMA_OFFSET = BLOB_BYTES;                 // position in the device constant blob
constexpr size_t TOTAL_BLOB_BYTES = BLOB_BYTES + MMA_BYTES;
// the tables (the momentum table has k_max + 2 entries) must fit shared memory; otherwise the scalar kernel runs
constexpr bool HAS_MMA = MMA_SHAPE_OK && MMA_BYTES <= 200 * 1024;

static inline void fill_mma_tables(const spcies_consts &C, const FistaDerived &D, MmaTables &T) {
    memset(&T, 0, sizeof T);
    for (int c = 0; c < 8; ++c) {
        T.xat[c] = x_at(c);
        T.uat[c] = u_at(c);
    }
    for (int oc = 0; oc < 8; ++oc)
        for (int ic = 0; ic < 8; ++ic) {
            if (z_at(oc) >= 0 && x_at(ic) >= 0)
                T.SNABt[oc * 8 + ic] = -(PRESCALE ? (double)C.QRi[z_at(oc)] : 1.0) * (double)C.AB[x_at(ic)][z_at(oc)];
            if (x_at(oc) >= 0 && z_at(ic) >= 0) T.NAB[oc * 8 + ic] = -(double)C.AB[x_at(oc)][z_at(ic)];
        }
    for (int l = 0; l < N; ++l) {
        if constexpr (!MERGE) {
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    T.Linv[l][i * 8 + j] = (double)D.Linv[l][i][j];
                    T.NF[l][i * 8 + j] = -(double)D.F[l][i][j];
                    T.Uinv[l][i * 8 + j] = (double)D.Uinv[l][i][j];
                    T.NG[l][i * 8 + j] = -(double)D.G[l][i][j];
                }
        } else {
            for (int lane = 0; lane < 32; ++lane) {
                const int o = xo_at(lane / 4), t = lane % 4;
                if (o < 0) continue;
                const int e0 = x_at(2 * t);                      // component in the full k-step (columns 0,2,4,6)
                const int e1 = 4 + (t < 2 ? t : t - 2);          // component in the shared k-step
                if (e0 >= 0) {
                    T.FWa[l][lane] = make_double2((double)D.Linv[l][o][e0], -(double)D.F[l][o][e0]);
                    T.BWa[l][lane] = make_double2((double)D.Uinv[l][o][e0], -(double)D.G[l][o][e0]);
                }
                if (e1 < n) {
                    T.FWb[l][lane] = t < 2 ? (double)D.Linv[l][o][e1] : -(double)D.F[l][o][e1];
                    T.BWb[l][lane] = t < 2 ? (double)D.Uinv[l][o][e1] : -(double)D.G[l][o][e1];
                }
            }
        }
    }
    for (int s = 0; s <= N; ++s)
        for (int c = 0; c < 8; ++c) {
            T.LBs[s][c] = -1e300;
            T.UBs[s][c] = 1e300;
        }
    for (int c = 0; c < 8; ++c) {
        const int e = x_at(c), j = u_at(c), z = z_at(c);
        if (z >= 0) T.QRiz[c] = (double)C.QRi[z];
        if (e >= 0) {
            T.QRiy[c] = (double)C.QRi[e];
            T.Qs[c] = (PRESCALE ? (double)C.QRi[e] : 1.0) * (double)C.Q[e];
#if SPCIES_TERMINAL
            T.Ti[c] = (double)C.Ti[e];
            T.Ts[c] = (PRESCALE ? (double)C.Ti[e] : 1.0) * (double)C.T[e];
#else
            T.Ts[c] = 1.0;
#endif
        }
        if (j >= 0) T.Qs[c] = (PRESCALE ? (double)C.QRi[n + j] : 1.0) * (double)C.R[j];
#ifdef VAR_BOUNDS
        if (j >= 0) {
            T.LBs[0][c] = (double)C.LB0[j];
            T.UBs[0][c] = (double)C.UB0[j];
        }
        if (z >= 0)
            for (int l = 0; l < N - 1; ++l) {
                T.LBs[l + 1][c] = (double)C.LB[l][z];
                T.UBs[l + 1][c] = (double)C.UB[l][z];
            }
#if SPCIES_TERMINAL
        if (e >= 0) {
            T.LBs[N][c] = (double)C.LBN[e];
            T.UBs[N][c] = (double)C.UBN[e];
        }
#endif
#else
        if (z >= 0)
            for (int s = 0; s <= N; ++s) {
                T.LBs[s][c] = (double)C.LB[z];
                T.UBs[s][c] = (double)C.UB[z];
            }
#endif
    }
    // t_0 = 1;  t_k = (1 + sqrt(1 + 4 t_{k-1}^2)) / 2;  beta_k = (t_{k-1} - 1) / t_k      code_laxMPC_FISTA_C.c:370-385
    volatile double t = 1.0;
    T.beta[0] = 0.0;
    for (int k = 1; k < MMA_KTAB; ++k) {
        const double t1 = t;
        volatile double s = 4.0 * t1;
        s = s * t1;
        s = 1.0 + s;
        t = 0.5 * (1.0 + sqrt(s));
        T.beta[k] = (t1 - 1.0) / t;
    }
}

using mma::dmma;
// out = c + M v   (M = the lane's double2 of the row-major 8 x 8 matrix; out may alias c)
__device__ __forceinline__ void mma_mv(double (&out)[2], const double2 M, const double (&v)[2], const double c0, const double c1) {
    mma::mv(out, M, v, c0, c1);
}
__device__ __forceinline__ double2 mma_mat(const double *M, int lane) { return reinterpret_cast<const double2 *>(M)[lane]; }

// SS (bulk configuration, MERGE layout only): lambda and mu' live in shared memory as [vector][stage][thread] double2 (one
// conflict-free 128-bit access per lane) instead of registers; only y stays in registers, so that 4 warps per SM scheduler fit the
// register file without spills (128 registers) and keep the FP64 pipe busy.
// CL: closed-loop run (io.cl_steps sampling times per instance, see the block at the exit test)
template <bool VARB, int BLOCK, bool CL>
__global__ void __launch_bounds__(BLOCK, 1) fista_mma_kernel(const BatchIO io, const unsigned char *__restrict__ g_blob) {
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    const spcies_consts *C = reinterpret_cast<const spcies_consts *>(g_blob);   // engineering-unit scaling only
    (void)C;
    const MmaTables *T = reinterpret_cast<const MmaTables *>(smem_raw);
    stage_constants(smem_raw, g_blob + MMA_OFFSET, (uint32_t)MMA_BYTES, &mbar);

    const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int cc[2] = {2 * t4, 2 * t4 + 1};                       // the two columns of this lane
    const int xe[2] = {T->xat[cc[0]], T->xat[cc[1]]};             // component of x they hold (or -1) ...
    const int ue[2] = {T->uat[cc[0]], T->uat[cc[1]]};             // ... component of u (or -1)
    const bool xs[2] = {xe[0] >= 0, xe[1] >= 0};
    const bool us[2] = {ue[0] >= 0, ue[1] >= 0};
    const unsigned gmask = 0xFu << (4 * g);
    const bool leader = t4 == 0;
    const bool lo2 = t4 < 2;                                      // MERGE: lanes whose register 1 holds x_4.. (the others: second copies)

    const double2 nabt = mma_mat(T->SNABt, lane), nab = mma_mat(T->NAB, lane);
    double qri[2], qriy[2], qs[2], ts[2], lb[2], ub[2];
#if SPCIES_TERMINAL
    double ti[2];
#endif
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        qri[i] = T->QRiz[cc[i]];
        qriy[i] = T->QRiy[cc[i]];
        qs[i] = T->Qs[cc[i]];
        ts[i] = T->Ts[cc[i]];
#if SPCIES_TERMINAL
        ti[i] = T->Ti[cc[i]];
#endif
        lb[i] = T->LBs[1][cc[i]];
        ub[i] = T->UBs[1][cc[i]];
    }
    // bounds of stage row s (0: u_0, 1..N-1: stage s-1, N: terminal)
    auto bnd = [&](int s, int i, double &lo, double &hi) {
#ifdef VAR_BOUNDS
        if (!VARB) {
            lo = T->LBs[s][cc[i]];
            hi = T->UBs[s][cc[i]];
            return;
        }
#endif
        lo = lb[i];
        hi = ub[i];
    };

    const double tolv[2] = {xs[0] ? (double)tol : 1e300, xs[1] ? (double)tol : 1e300};   // the exit test looks at state components
    // io.phase 2: the "batch" is the list of records parked by the previous launch (count in queue[9], records in park_in);
    // with io.cap > 0 a resumed instance may be parked again (into io.park).
    const bool resume = io.phase == 2;
    const long long B = resume ? min((long long)io.queue[9], io.park_in_cap) : io.B;
    const WorkQueue wq{resume ? io.queue + 7 : io.queue, B, resume ? nullptr : io.ready};
    const WorkQueue marks{io.queue, B, nullptr};
    marks.mark_start();
    unsigned long long stat_k = 0;
    unsigned int stat_nc = 0;

    long long inst = -1;
    int k = 0, grace_left = io.grace;
    bool live = false, drained = false, park_ok = io.park != nullptr;
    double y[N][2], lam[N][2], w[N][2], q[2] = {0, 0}, qT[2] = {0, 0};
    auto getL = [&](int l, double (&v)[2]) {
        v[0] = lam[l][0];
        v[1] = lam[l][1];
    };
    auto setL = [&](int l, const double (&v)[2]) {
        lam[l][0] = v[0];
        lam[l][1] = v[1];
    };
    auto getW = [&](int l, double (&v)[2]) {
        v[0] = w[l][0];
        v[1] = w[l][1];
    };
    auto setW = [&](int l, const double (&v)[2]) {
        w[l][0] = v[0];
        w[l][1] = v[1];
    };
    int cl_step = 0;                            // CL: sampling time of the instance this lane group holds
    bool cl_restart = false;                    // CL: the instance finished a sampling time and starts the next one in place
    double xnext[2] = {0.0, 0.0};               // CL: its successor state x+ = A x + B u, by column
    double lo0[2] = {0, 0}, hi0[2] = {0, 0};   // "stage -1": lo = hi = x0 on the state components, the bounds of u_0 on the others
#pragma unroll
    for (int l = 0; l < N; ++l)
#pragma unroll
        for (int i = 0; i < 2; ++i) y[l][i] = lam[l][i] = w[l][i] = 0.0;

    for (;;) {
        // ---- refill: a lane group without an instance pulls the next one          code_laxMPC_FISTA_C.c:94-100, 275-289
        const bool need = !live && !drained;
        if (__any_sync(FULL, need)) {
            long long slot = -1;
            if (need && leader) slot = wq.next();
            slot = __shfl_sync(FULL, slot, lane & ~3);
            if (need) {
                if (slot < 0) {
                    drained = true;
                    if (leader) marks.mark_drained();
                } else {
                    const double *pk = io.park_in + slot;
                    inst = resume ? __double_as_longlong(pk[0]) : slot;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double xr_ = xs[i] ? eng_x(C, io.xr, inst, n, xe[i]) : 0.0;
                        const double ur_ = us[i] ? eng_u(C, io.ur, inst, m, ue[i]) : 0.0;
                        q[i] = qs[i] * (xs[i] ? xr_ : ur_);          // QRi o [Q xr; R ur]
                        qT[i] = ts[i] * xr_;                          // Ti o T xr (lax)  |  xr (equ)
                        if (VARB) {
                            const int ze = xs[i] ? xe[i] : n + ue[i];
                            lb[i] = (xs[i] || us[i]) ? io.LB[inst * nm + ze] : -1e300;
                            ub[i] = (xs[i] || us[i]) ? io.UB[inst * nm + ze] : 1e300;
                        }
                        double l0, h0;
                        bnd(0, i, l0, h0);
                        const double x0_ = xs[i] ? eng_x(C, io.x0, inst, n, xe[i]) : 0.0;
                        lo0[i] = xs[i] ? x0_ : l0;
                        hi0[i] = xs[i] ? x0_ : h0;
                    }
                    if (resume) {
                        k = (int)__double_as_longlong(pk[1 * io.park_in_cap]);
#pragma unroll
                        for (int l = 0; l < N; ++l)
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                y[l][i] = xs[i] ? pk[(3 + l * n + xe[i]) * io.park_in_cap] : 0.0;
                                lam[l][i] = xs[i] ? pk[(3 + N * n + l * n + xe[i]) * io.park_in_cap] : 0.0;
                            }
                    } else {
                        k = -1;                                       // the warm-up pass brings it to 0
#pragma unroll
                        for (int l = 0; l < N; ++l)
#pragma unroll
                            for (int i = 0; i < 2; ++i) y[l][i] = lam[l][i] = 0.0;
                        cl_step = 0;
                    }
                    live = true;
                }
            }
        }
        if (!__any_sync(FULL, live)) break;
        const bool dry = (io.phase == 1) && (*(volatile unsigned long long *)io.queue >= (unsigned long long)B);

        // ================= pass A: z(y) -> residual -> exit test -> forward step =================
        // A warp issues in order, and one DMMA takes 16 issue cycles of the tensor pipe but 26 cycles until its result can be
        // used: the source order below is the schedule.  Products that do not depend on each other (the same product of
        // every stage) are issued back to back, k-step 0 of all stages and then k-step 1 of all stages, so the only
        // latency-exposed MMAs are the two recurrences (mu_l forward, d_lambda_l backward), and the forward one is
        // interleaved with the independent w_{l-1} = Uinv_{l-1} mu_{l-1} products.
        // Components that a vector does not use hold finite don't-care values where the consuming matrix has zero columns
        // (the input part of z_l inside r_l), exact zeros elsewhere; the exit test only looks at state components.
        bool over = false;
        double u0v[2];
        double zz[N + 1][2];   // zz[0] = (x0, u_0), zz[l+1] = z_l (l < N-1), zz[N] = z_N (lax) | xr (equ)
        {
            // QRi o (q - [A B]' y_s) for every stage s                                      :474-519
            double e[N][2];
#pragma unroll
            for (int s_ = 0; s_ < N; ++s_) dmma(e[s_][0], e[s_][1], y[s_][0], nabt.x, q[0], q[1]);
#pragma unroll
            for (int s_ = 0; s_ < N; ++s_) dmma(zz[s_][0], zz[s_][1], y[s_][1], nabt.y, e[s_][0], e[s_][1]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)                                                        // u_0; lo0 = hi0 = x0 on the state part
            zz[0][i] = clip_mma(PRESCALE ? zz[0][i] : zz[0][i] * qri[i], lo0[i], hi0[i]);
        u0v[0] = zz[0][0];
        u0v[1] = zz[0][1];
#pragma unroll
        for (int l = 0; l < N - 1; ++l)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                double lo, hi;
                bnd(l + 1, i, lo, hi);
                zz[l + 1][i] = clip_mma(PRESCALE ? fma(qriy[i], y[l][i], zz[l + 1][i]) : (zz[l + 1][i] + y[l][i]) * qri[i], lo, hi);   // z_l  :494-519
            }
#if SPCIES_TERMINAL
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            double lo, hi;
            bnd(N, i, lo, hi);
            zz[N][i] = clip_mma(PRESCALE ? fma(ti[i], y[N - 1][i], qT[i]) : (qT[i] + y[N - 1][i]) * ti[i], lo, hi);   // z_N  :522-537
        }
#else
        zz[N][0] = qT[0];                                                                   // xr   code_equMPC_FISTA_C.c:549
        zz[N][1] = qT[1];
#endif
        // r_l = x_{l+1} - [A B] z_{l-1}   (r_0 = x_1 - A x0 - B u_0)                         :549-572
        double r[N][2];
        {
            double e[N][2];
#pragma unroll
            for (int l = 0; l < N; ++l) dmma(e[l][0], e[l][1], zz[l][0], nab.x, zz[l + 1][0], zz[l + 1][1]);
#pragma unroll
            for (int l = 0; l < N; ++l) dmma(r[l][0], r[l][1], zz[l][1], nab.y, e[l][0], e[l][1]);
        }
#if SPCIES_FISTA_MMA_INT_TEST
        {   // |r| > tol on the integer pipe (the FP64 datapath is the bound of this kernel): for non-negative doubles the bit patterns
            // order like the values, so |r| > tol  <=>  (bits(r) & ~sign) > bits(tol); a NaN residual counts as "not converged"
            const unsigned long long tb0 = (unsigned long long)__double_as_longlong(tolv[0]), tb1 = (unsigned long long)__double_as_longlong(tolv[1]);
            constexpr unsigned long long ABS = 0x7fffffffffffffffULL;
#pragma unroll
            for (int l = 0; l < N; ++l)
                over = over || (((unsigned long long)__double_as_longlong(r[l][0]) & ABS) > tb0) ||
                       (((unsigned long long)__double_as_longlong(r[l][1]) & ABS) > tb1);
        }
#else
#pragma unroll
        for (int l = 0; l < N; ++l) over = over || (fabs(r[l][0]) > tolv[0]) || (fabs(r[l][1]) > tolv[1]);
#endif
        // forward step: mu_l = Linv_l r_l - F_l mu_{l-1}   [plain layout: and w_l = Uinv_l mu_l, stored in w]
        if constexpr (MERGE) {
            // w[l] keeps mu_l.  Per stage: e_l = Linv r (full k-step, issued ahead), f = e_l - F mu_{l-1} (full k-step),
            // mu_l = f + [Linv | -F] (r_l, mu_{l-1}) (the shared k-step)
            double e[N][2];
#pragma unroll
            for (int l = 0; l < N; ++l) dmma(e[l][0], e[l][1], r[l][0], T->FWa[l][lane].x, 0.0, 0.0);
            double mu[2];
            dmma(mu[0], mu[1], r[0][1], T->FWb[0][lane], e[0][0], e[0][1]);
            setW(0, mu);
#pragma unroll
            for (int l = 1; l < N; ++l) {
                double f0, f1;
                dmma(f0, f1, mu[0], T->FWa[l][lane].y, e[l][0], e[l][1]);
                dmma(mu[0], mu[1], lo2 ? r[l][1] : mu[1], T->FWb[l][lane], f0, f1);
                setW(l, mu);
            }
        } else {
            double sl[N][2];
            {
                double e[N][2];
                double2 li[N];
#pragma unroll
                for (int l = 0; l < N; ++l) li[l] = mma_mat(T->Linv[l], lane);
#pragma unroll
                for (int l = 0; l < N; ++l) dmma(e[l][0], e[l][1], r[l][0], li[l].x, 0.0, 0.0);
#pragma unroll
                for (int l = 0; l < N; ++l) dmma(sl[l][0], sl[l][1], r[l][1], li[l].y, e[l][0], e[l][1]);
            }
            double mu[2] = {sl[0][0], sl[0][1]};
#pragma unroll
            for (int l = 1; l < N; ++l) {
                const double2 nf = mma_mat(T->NF[l], lane), ui = mma_mat(T->Uinv[l - 1], lane);
                double e0, e1, f0, f1, m0, m1;
                dmma(e0, e1, mu[0], nf.x, sl[l][0], sl[l][1]);       // recurrence, k-step 0
                dmma(f0, f1, mu[0], ui.x, 0.0, 0.0);                 // w_{l-1}, k-step 0
                dmma(m0, m1, mu[1], nf.y, e0, e1);                   // recurrence, k-step 1
                dmma(w[l - 1][0], w[l - 1][1], mu[1], ui.y, f0, f1); // w_{l-1}, k-step 1
                mu[0] = m0;
                mu[1] = m1;
            }
            mma_mv(w[N - 1], mma_mat(T->Uinv[N - 1], lane), mu, 0.0, 0.0);
        }
        // ================= exit condition                                            :337-361 =================
        if (live) k += 1;
        const bool gover = (__ballot_sync(FULL, over) & gmask) != 0u;
        if (live && k >= 1) {
            int ef = 0;
            if (!gover) ef = 1;
            else if (k >= k_max) ef = -1;
            if (ef != 0) {
                const long long o = CL ? (long long)cl_step * io.cl_ld + inst : inst;   // CL: trajectories are [step][instance]
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (us[i]) io.u[o * m + ue[i]] = eng_u_out(C, u0v[i], ue[i]);
                if (leader) {
                    io.k[o] = k;
                    io.e[o] = ef;
                    stat_k += (unsigned long long)k;
                    stat_nc += (ef < 0);
                }
                live = false;
                if (CL) cl_restart = true;      // the sampling time is over: successor state, then the next one (if any) in place
            }
        }
        if constexpr (CL) {
            // Closed loop (examples/cl_in_C/main_cl_in_C.c:100-117 for every instance of the batch, without leaving the chip): the
            // plant is the prediction model, and -[A B] (x, u_0) is the product the residual r_0 already uses, so the successor
            // state comes out of one more MMA in the column layout of x -- no shuffles.  The instance then restarts in place:
            // x0 <- x+, k <- -1 (warm-up pass), and either y = lambda = 0 (cold start, what a loop of reference calls does) or the
            // y of the exit test kept as the starting dual point (io.cl_warm: the `lambda` argument of
            // platforms/Matlab/spcies_laxMPC_FISTA_solver.m:161-164, :266-283 -- sol.lambda of the previous sampling time).
            if (__any_sync(FULL, cl_restart)) {
                double nx[2];
                mma_mv(nx, nab, zz[0], 0.0, 0.0);                  // -(A x + B u_0)
                if (cl_restart) {
                    cl_step += 1;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        xnext[i] = -nx[i];
                        if (xs[i]) {
                            lo0[i] = hi0[i] = xnext[i];
                            if (io.cl_x) io.cl_x[((long long)cl_step * io.cl_ld + inst) * n + xe[i]] = xnext[i];
                        }
                    }
                    if (cl_step < io.cl_steps) {
                        k = -1;
                        live = true;
                    } else {
                        cl_restart = false;                        // that was the last sampling time of this instance
                    }
                }
            }
        }

        // ================= pass B: backward step, lambda and y updates      :368-385, :619-648 =================
        const double beta = T->beta[(live && k > 0) ? k : 0];
        const bool hold = CL && cl_restart;                         // this pass belongs to no iteration of a restarting instance
        const double hkeep = (CL && io.cl_warm != 0) ? 1.0 : 0.0;
        auto update = [&](int l, const double (&d)[2]) {
            double l1[2], ln[2];
            getL(l, l1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                ln[i] = y[l][i] + d[i];                             // lambda_l = y_l + d_lambda_l
                const double yn = fma(beta, ln[i] - l1[i], ln[i]);  // y_l = lambda_l + beta (lambda_l - lambda1_l)
                y[l][i] = hold ? hkeep * y[l][i] : yn;
            }
            setL(l, ln);
        };
        if constexpr (MERGE) {
            // d_lambda_l = Uinv_l mu_l - G_l d_lambda_{l+1}:  g_l = Uinv mu (full k-step, issued one stage ahead),
            // h = g_l - G d_{l+1} (full k-step), d_l = h + [Uinv | -G] (mu_l, d_{l+1}) (the shared k-step)
            double d[2], g0, g1, gn0 = 0.0, gn1 = 0.0, wl[2], wp[2];   // mu'_l, mu'_{l-1}
            getW(N - 1, wl);
            getW(N - 2, wp);
            dmma(g0, g1, wl[0], T->BWa[N - 1][lane].x, 0.0, 0.0);
            dmma(gn0, gn1, wp[0], T->BWa[N - 2][lane].x, 0.0, 0.0);
            dmma(d[0], d[1], wl[1], T->BWb[N - 1][lane], g0, g1);
            update(N - 1, d);
#pragma unroll
            for (int l = N - 2; l >= 0; --l) {
                double h0, h1;
                wl[0] = wp[0];
                wl[1] = wp[1];
                if (l > 0) getW(l - 1, wp);
                dmma(h0, h1, d[0], T->BWa[l][lane].y, gn0, gn1);
                if (l > 0) dmma(gn0, gn1, wp[0], T->BWa[l - 1][lane].x, 0.0, 0.0);
                dmma(d[0], d[1], lo2 ? wl[1] : d[1], T->BWb[l][lane], h0, h1);
                update(l, d);
            }
        } else {
            double d[2] = {w[N - 1][0], w[N - 1][1]};
#pragma unroll
            for (int l = N - 1; l >= 0; --l) {
                if (l < N - 1) mma_mv(d, mma_mat(T->NG[l], lane), d, w[l][0], w[l][1]);   // d_lambda_l = w_l - G_l d_lambda_{l+1}
                update(l, d);
            }
        }

        if constexpr (CL) {
            // io.cl_warm == 2: the receding horizon moved by one stage, so does the dual starting point (lambda_l <- lambda_{l+1},
            // the last block is repeated)
            if (io.cl_warm == 2 && __any_sync(FULL, cl_restart)) {
#pragma unroll
                for (int l = 0; l < N - 1; ++l)
#pragma unroll
                    for (int i = 0; i < 2; ++i) y[l][i] = cl_restart ? y[l + 1][i] : y[l][i];
            }
            cl_restart = false;
        }

        // ---- parking: an instance that is still running `grace` iterations after the queue ran dry (io.phase 1), or that
        //      has reached the iteration cap of this launch (io.cap), hands its iterates to the next launch
        bool park_now = false;
        if (dry && live && k >= 1) {
            park_now = grace_left <= 0;
            grace_left -= 1;
        }
        if (io.cap > 0 && live && k >= io.cap) park_now = true;
        park_now = park_now && park_ok;
        if (__any_sync(FULL, park_now)) {
            long long pslot = 0;
            if (park_now && leader) pslot = (long long)atomicAdd(io.queue + 6, 1ULL);
            pslot = __shfl_sync(FULL, pslot, lane & ~3);
            if (park_now && pslot >= io.park_cap) {                 // no room: the instance stays where it is
                park_now = false;
                park_ok = false;
            }
            if (park_now) {
                double *pk = io.park + pslot;
                if (leader) {
                    atomicAdd(io.queue + 10, 1ULL);
                    pk[0] = __longlong_as_double(inst);
                    pk[1 * io.park_cap] = __longlong_as_double((long long)k);
                    pk[2 * io.park_cap] = 0.0;                      // t: implied by k in this engine
                }
#pragma unroll
                for (int l = 0; l < N; ++l) {
                    double lv[2];
                    getL(l, lv);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        if (xs[i]) {
                            pk[(3 + l * n + xe[i]) * io.park_cap] = y[l][i];
                            pk[(3 + N * n + l * n + xe[i]) * io.park_cap] = lv[i];
                        }
                }
                live = false;
            }
        }
    }
    flush_stats(io.queue, stat_k, stat_nc);
    marks.mark_end();
}
