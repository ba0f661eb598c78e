// spcies_mma.cuh -- building blocks of the FP64 tensor-core engines (MPC_FISTA_mma.cuh, MPC_ADMM_mma.cuh).
//
// The engines run 8 instances per warp: instance g = lane / 4 is one row of the m8n8k4 FP64 MMA (SASS DMMA.8x8x4), its 4
// lanes hold every vector of the instance, 2 "columns" per lane (column c in lane t = c / 2, register c % 2).  That is the
// C/D fragment layout and -- one register per k-step -- the A fragment layout, so shared-matrix x per-instance-vector
// products chain without shuffles; the shared matrix is the B fragment (lane (o = lane / 4, t = lane % 4) holds
// M[output column o][input column of k-slot t]).
//
// MmaLayout<n, m>: the column assignment of the components of x (n of them) and u (m), n + m <= 8, 5 <= n <= 6:
//   x_0..x_3 -> columns 0,2,4,6 (register 0 of the four lanes: a full k-step), x_4.. -> 1,3 (register 1 of lanes 0,1: half a
//   k-step), u_j -> the register-1 columns after them.  Vectors produced by the block recurrences of the W solve carry a
//   second copy of x_4.. in register 1 of lanes 2,3 (duplicated matrix rows), so that the half k-steps of two products that add
//   up merge into one:  a = lane % 4 < 2 ? first.reg1 : second.reg1.
#pragma once
#include <cuda_runtime.h>
#include <string.h>

namespace spcies {
namespace mma {

// D = A B + C on one k-step (4 input columns) of the 8 instances of the warp
__device__ __forceinline__ void dmma(double &d0, double &d1, const double a, const double b, const double c0, const double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
// out = c + M v over both k-steps (M = the lane's double2 of a row-major 8 x 8 matrix [output column][input column])
__device__ __forceinline__ void mv(double (&out)[2], const double2 M, const double (&v)[2], const double c0, const double c1) {
    double e0, e1;
    dmma(e0, e1, v[0], M.x, c0, c1);
    dmma(out[0], out[1], v[1], M.y, e0, e1);
}

template <int n, int m> struct MmaLayout {
    // the layout exists for 5 <= nn_ <= 6, nn_ + mm_ <= 8; the engines test OK (their MMA_SHAPE_OK) before they are used -- for any
    // other system the tables are compiled but never filled or read, and the solver runs on its one-thread-per-instance kernel
    static constexpr bool OK = n >= 5 && n <= 6 && n + m <= 8;
    __host__ __device__ static constexpr int col_x(int e) { return e < 4 ? 2 * e : 2 * (e - 4) + 1; }
    __host__ __device__ static constexpr int col_u(int j) { return 2 * (n - 4 + j) + 1; }
    __host__ __device__ static constexpr int col_dup(int e) { return 2 * (e - 4 + 2) + 1; }   // second copy of x_e, e >= 4
    __host__ __device__ static constexpr int x_at(int c) {
        for (int e = 0; e < n; ++e)
            if (col_x(e) == c) return e;
        return -1;
    }
    __host__ __device__ static constexpr int u_at(int c) {
        for (int j = 0; j < m; ++j)
            if (col_u(j) == c) return j;
        return -1;
    }
    // component of z = (x, u) at a column
    __host__ __device__ static constexpr int z_at(int c) { return x_at(c) >= 0 ? x_at(c) : (u_at(c) >= 0 ? n + u_at(c) : -1); }
    // component of a recurrence vector produced in output column c (including the second copies)
    __host__ __device__ static constexpr int xo_at(int c) {
        if (x_at(c) >= 0) return x_at(c);
        for (int e = 4; e < n; ++e)
            if (col_dup(e) == c) return e;
        return -1;
    }
};

// Explicit inverses of the block-bidiagonal Cholesky factor W = R'R of the generated solvers (Alpha = super-diagonal blocks,
// Beta = diagonal blocks with the diagonal stored inverted; compute_laxMPC_FISTA_ingredients.m:137-150), in extended precision:
//   forward   mu_l = Linv_l r_l - F_l mu_{l-1},     Linv_l = (U_l')^-1,  F_l = Linv_l Alpha_{l-1}'
//   backward  d_l  = Uinv_l mu_l - G_l d_{l+1},     Uinv_l = U_l^-1,     G_l = Uinv_l Alpha_l
template <int N, int n, class C, class R>
static inline void block_inverses(const C &c, R (*Linv)[n][n], R (*F)[n][n], R (*Uinv)[n][n], R (*G)[n][n]) {
    typedef long double ld;
    for (int l = 0; l < N; ++l) {
        ld U[n][n] = {}, Ui[n][n] = {};
        for (int i = 0; i < n; ++i)
            for (int j = i; j < n; ++j) U[i][j] = (i == j) ? (ld)1 / (ld)c.Beta[l][j][j] : (ld)c.Beta[l][i][j];
        for (int col = 0; col < n; ++col)            // U * Ui[:, col] = e_col by back substitution
            for (int i = n - 1; i >= 0; --i) {
                ld v = (i == col) ? (ld)1 : (ld)0;
                for (int j = i + 1; j < n; ++j) v -= U[i][j] * Ui[j][col];
                Ui[i][col] = v / U[i][i];
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                Uinv[l][i][j] = (j >= i) ? (R)Ui[i][j] : R(0);
                Linv[l][i][j] = (j <= i) ? (R)Ui[j][i] : R(0);
                F[l][i][j] = R(0);
                G[l][i][j] = R(0);
            }
        if (l >= 1)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    ld v = 0;
                    for (int k = 0; k <= j; ++k) v += Ui[k][j] * (ld)c.Alpha[l - 1][i][k];   // Linv[j][k] * Alpha'[k][i]
                    F[l][j][i] = (R)v;
                }
        if (l <= N - 2)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    ld v = 0;
                    for (int k = j; k < n; ++k) v += Ui[j][k] * (ld)c.Alpha[l][k][i];
                    G[l][j][i] = (R)v;
                }
    }
}

// The four per-stage fragment tables of the merged recurrences, per lane (o = lane / 4 output column, t = lane % 4):
//   FWa = (Linv[o][x_t], -F[o][x_t]),  FWb = t < 2 ? Linv[o][x_{4+t}] : -F[o][x_{4+t-2}]
//   BWa = (Uinv[o][x_t], -G[o][x_t]),  BWb = t < 2 ? Uinv[o][x_{4+t}] : -G[o][x_{4+t-2}]
template <int N, int n, int m, class R>
static inline void recurrence_fragments(const R (*Linv)[n][n], const R (*F)[n][n], const R (*Uinv)[n][n], const R (*G)[n][n],
                                        double2 (*FWa)[32], double (*FWb)[32], double2 (*BWa)[32], double (*BWb)[32]) {
    typedef MmaLayout<n, m> L;
    for (int l = 0; l < N; ++l)
        for (int lane = 0; lane < 32; ++lane) {
            FWa[l][lane] = BWa[l][lane] = make_double2(0.0, 0.0);
            FWb[l][lane] = BWb[l][lane] = 0.0;
            const int o = L::xo_at(lane / 4), t = lane % 4;
            if (o < 0) continue;
            const int e0 = L::x_at(2 * t);               // component in the full k-step (columns 0,2,4,6)
            const int e1 = 4 + (t < 2 ? t : t - 2);      // component in the shared k-step
            if (e0 >= 0) {
                FWa[l][lane] = make_double2((double)Linv[l][o][e0], -(double)F[l][o][e0]);
                BWa[l][lane] = make_double2((double)Uinv[l][o][e0], -(double)G[l][o][e0]);
            }
            if (e1 < n) {
                FWb[l][lane] = t < 2 ? (double)Linv[l][o][e1] : -(double)F[l][o][e1];
                BWb[l][lane] = t < 2 ? (double)Uinv[l][o][e1] : -(double)G[l][o][e1];
            }
        }
}

}  // namespace mma
}  // namespace spcies
