// MPC_FISTA_dense.cuh -- policy of the laxMPC / equMPC FISTA solvers for the generic dense tensor-core engine
// (spcies_dense_mma.cuh); included by MPC_FISTA.cuh, inside spcies::fista, after MPC_FISTA_single.cuh.
//
// The banded engine (MPC_FISTA_mma.cuh) takes systems with nn_ + mm_ <= 8 and N <= 12 (everything in registers).  Any other
// system used to fall back to the one-thread-per-instance kernel; this policy puts it on the tensor cores instead, through the
// formulation of the latency engine (MPC_FISTA_single.cuh: PrimalForm): in the space of the primal variable the FISTA iteration
// of code_laxMPC_FISTA_C.c:323-389 is one product with a matrix that all instances share,
//     [ mu+ - v ]   [ P  G ] [ z ]        mu+ = v + g + P z,   v+ = mu+ + beta_k (mu+ - mu),   z+ = clip(Hd o (q + v+)),
//     [   r     ] = [ E  B ] [ c ],       exit test on r = b + E z of the z the pass started from,
// with w = z (|z| / 8 tiles), c = (x0, xr, ur), P = E' W^-1 E, G / B = the columns that turn c into g / b.  The first pass is the
// reference's initial step (no exit test, not counted: WARMUP = 1).  z is double buffered so that u_opt = z_0 of the pass whose
// residual met the tolerance.
#pragma once
// (spcies_dense_mma.cuh is included by the parent header, outside its namespace)

#ifndef SPCIES_FISTA_DENSE
#define SPCIES_FISTA_DENSE 1
#endif

struct DenseEngine {
    static constexpr int ZT = (SG_ZLEN + 7) / 8, RT = (SG_ROWS + 7) / 8;
    static constexpr int NO = ZT + RT, NW = ZT, NC = (2 * n + m + 7) / 8;
    static constexpr int NSTATE = 5 * ZT;            // mu, v, q, z (two buffers)
    static constexpr int BLK_MU = 0, BLK_V = ZT, BLK_Q = 2 * ZT, BLK_Z = 3 * ZT;
    static constexpr int NB = 4, TEAM = 1;
    static constexpr int WARMUP = 1;
    static constexpr bool OK = SPCIES_FISTA_DENSE != 0 && SG_KTAB <= 2048;
    static constexpr size_t BLOB_OFFSET = SINGLE_OFFSET + (HAS_SINGLE ? SINGLE_BYTES : 0);
    struct alignas(16) Small {
        double Hd[ZT][8], LB[ZT][8], UB[ZT][8], cq[ZT][8];
        int qsrc[ZT][8];
        double beta[SG_KTAB];
    };
    struct Lane {
        int pass;                                    // kept by the engine: index of the current pass (0 = the initial step)
    };
    __device__ static __forceinline__ void lane_reset(Lane &L) { L.pass = 0; }

    static inline void fill(const PrimalForm &F, Small &S, long double *Fm) {
        constexpr int NINC = (NW + NC) * 8;
        for (int t = 0; t < ZT; ++t)
            for (int c = 0; c < 8; ++c) {
                const int e = t * 8 + c;
                const bool in = e < SG_ZLEN;
                S.Hd[t][c] = in ? F.Hd[e] : 0.0;
                S.LB[t][c] = in ? F.LB[e] : 0.0;
                S.UB[t][c] = in ? F.UB[e] : 0.0;
                S.cq[t][c] = in ? F.cq[e] : 0.0;
                S.qsrc[t][c] = in ? F.qsrc[e] : 0;
            }
        memcpy(S.beta, F.beta, sizeof S.beta);
        // c = (x0 [n], xr [n], ur [m]) behind the NW tiles of z
        for (int e = 0; e < SG_ZLEN; ++e) {
            long double *row = Fm + (size_t)e * NINC;
            for (int col = 0; col < SG_ZLEN; ++col) row[col] = F.P[(size_t)e * SG_ZLEN + col];
            for (int c = 0; c < n; ++c) {
                row[NW * 8 + c] = F.G0[(size_t)e * n + c];
                row[NW * 8 + n + c] = F.G1[(size_t)e * n + c];
            }
        }
        for (int i = 0; i < SG_ROWS; ++i) {
            long double *row = Fm + (size_t)(ZT * 8 + i) * NINC;
            for (int col = 0; col < SG_ZLEN; ++col) row[col] = F.E[(size_t)i * SG_ZLEN + col];
            for (int c = 0; c < n; ++c) row[NW * 8 + c] = F.bA[i][c];
            if (!TERMINAL && i >= (N - 1) * n) row[NW * 8 + n + (i - (N - 1) * n)] = 1;      // equMPC: x_N = xr
        }
    }

    __device__ static __forceinline__ void init(Lane &L, const spcies_consts *C, const Small *S, const BatchIO &io, long long inst,
                                                double2 *st, double2 *cin, int t4, int /*rank*/) {
        L.pass = 0;
#pragma unroll 1
        for (int t = 0; t < ZT; ++t) {
            double q[2], z[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c = 2 * t4 + i, src = S->qsrc[t][c];
                const double ref = src < n ? eng_x(C, io.xr, inst, n, src) : eng_u(C, io.ur, inst, m, src - n);
                q[i] = S->cq[t][c] * ref;                                                 // :282-289 (Q, R, T stored negated)
                z[i] = clip(q[i] * S->Hd[t][c], S->LB[t][c], S->UB[t][c]);                 // z(lambda = 0) of the initial step
            }
            st[(BLK_MU + t) * 32] = make_double2(0.0, 0.0);
            st[(BLK_V + t) * 32] = make_double2(0.0, 0.0);
            st[(BLK_Q + t) * 32] = make_double2(q[0], q[1]);
            st[(BLK_Z + t) * 32] = make_double2(z[0], z[1]);
            st[(BLK_Z + ZT + t) * 32] = make_double2(0.0, 0.0);
        }
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
    __device__ static __forceinline__ double2 make_w(Lane &L, const spcies_consts *, const Small *, int t, const double2 *st, int /*t4*/) {
        return st[(BLK_Z + (L.pass & 1) * ZT + t) * 32];
    }
    __device__ static __forceinline__ void update(Lane &L, const spcies_consts *, const Small *S, int t0, const double (&acc)[NB][2],
                                                  double2 *st, int t4, bool &over) {
        const double tl = (double)tol;
        const int kp = L.pass;
        const double bk = S->beta[kp < SG_KTAB ? kp : SG_KTAB - 1];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int t = t0 + b;
            if (t >= NO) break;
            if (t < ZT) {
                const double2 mu = st[(BLK_MU + t) * 32], v = st[(BLK_V + t) * 32], q = st[(BLK_Q + t) * 32];
                const double2 hd = reinterpret_cast<const double2 *>(S->Hd[t])[t4], lo = reinterpret_cast<const double2 *>(S->LB[t])[t4],
                              hi = reinterpret_cast<const double2 *>(S->UB[t])[t4];
                const double m0 = v.x + acc[b][0], m1 = v.y + acc[b][1];                  // E' (y + W^-1 r)
                const double v0 = kp == 0 ? m0 : fma(bk, m0 - mu.x, m0), v1 = kp == 0 ? m1 : fma(bk, m1 - mu.y, m1);   // :311-320 | :372-385
                st[(BLK_MU + t) * 32] = make_double2(m0, m1);
                st[(BLK_V + t) * 32] = make_double2(v0, v1);
                st[(BLK_Z + ((kp + 1) & 1) * ZT + t) * 32] = make_double2(clip((q.x + v0) * hd.x, lo.x, hi.x), clip((q.y + v1) * hd.y, lo.y, hi.y));
            } else {
                over = over || (fabs(acc[b][0]) > tl) || (fabs(acc[b][1]) > tl);          // :337-348 (rows beyond N n are zero)
            }
        }
    }
    __device__ static __forceinline__ void finish(Lane &L, const spcies_consts *C, const BatchIO &io, long long inst, const double2 *st,
                                                  int t4) {
        // u_opt = z_0 of the pass whose residual was tested: the buffer that pass read   (:392-407)
        const int buf = L.pass & 1;
#pragma unroll
        for (int j = 0; j < m; ++j)
            if ((j % 8) / 2 == t4) {
                const double2 v = st[(BLK_Z + buf * ZT + j / 8) * 32];
                io.u[inst * m + j] = eng_u_out(C, (j & 1) ? v.y : v.x, j);
            }
    }
};
constexpr bool HAS_DENSE = !HAS_MMA && dense::Plan<DenseEngine>::HAS;
// AUTO picks it where it beats the one-thread-per-instance kernel (measured: n = 8, m = 4, N = 8, |z| = 96: 1.30x on 64 Ki instances,
// tools/fallback_times.py); smaller problems stay on the scalar kernel unless engine = SPCIES_CUDA_ENGINE_MMA asks for it
constexpr bool DENSE_PREFERRED = HAS_DENSE && SG_ZLEN >= 80;
