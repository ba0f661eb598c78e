// spcies_sparse.cuh -- device versions of the reference's sparse helpers and cone projections, operating on the
// per-instance [element][thread] state (StateRef) with the shared, read-only index / value arrays of the generator.
//
//   spmv_csr       +sp_utils/smv.m:23-36       (inlined in code_ellipMPC_ADMM_soc_C.c:157-162, :193-205)
//   ldl_solve_csc  +sp_utils/LDLsolve.m:22-49  (QDLDL-style; inlined in code_ellipMPC_ADMM_soc_C.c:172-188,
//                                               code_HMPC_ADMM_split_C.c:193-209)
//   proj_soc       +sp_utils/proj_SOC.m:12-27  (inlined in code_ellipMPC_ADMM_soc_C.c:223-242)
//   proj_soc3      snippets/proj_SOC3.c:4-35   (== +sp_utils/proj_SSOC.m:14-29 for dimension 3)
//   proj_diamond3  +sp_utils/proj_D.m:19-23    (SSOC(+1, lb) then SSOC(-1, ub); code_HMPC_ADMM_split_C.c:255-258)
//
// The host-side generators of the formats (full2CSR / full2CSC / full2LDL) are in spcies_b200/sp_utils.py.
// Accumulation order is the reference's (row-major CSR scan, column-oriented forward / row-oriented backward
// substitution), so Arith<EXACT> is bit-identical.
#pragma once
#include "spcies_common.cuh"

namespace spcies {

// y[i] (+)= sum_j val[j] * x[col[j]],  j in [row[i], row[i+1])
template <class A, bool ACCUMULATE, class ST, typename real>
__device__ __forceinline__ void spmv_csr(const ST &s, int off_y, int off_x, int nrow, const real *val, const int *col,
                                         const int *row) {
    int j = row[0];
#pragma unroll 1
    for (int i = 0; i < nrow; ++i) {
        const int jend = row[i + 1];
        real acc = ACCUMULATE ? s.ld(off_y + i) : real(0);
#pragma unroll 1
        for (; j < jend; ++j) acc = A::madd(acc, val[j], s.ld(off_x + col[j]));
        s.st(off_y + i, acc);
    }
}

// in-place solve of L D L' x = b with CSC (L - I) and Dinv
template <class A, class ST, typename real>
__device__ __forceinline__ void ldl_solve_csc(const ST &s, int off, int nrow, const real *L_val, const int *L_row,
                                              const int *L_col, const real *Dinv) {
    // forward substitution (column oriented): x[row[j]] -= val[j] * x[i]
#pragma unroll 1
    for (int i = 0; i < nrow; ++i) {
        const int jend = L_col[i + 1];
        const real xi = s.ld(off + i);
#pragma unroll 1
        for (int j = L_col[i]; j < jend; ++j) {
            const int r = L_row[j];
            s.st(off + r, A::nmsub(s.ld(off + r), L_val[j], xi));
        }
    }
    // x *= Dinv
#pragma unroll 1
    for (int i = 0; i < nrow; ++i) s.st(off + i, A::mul(s.ld(off + i), Dinv[i]));
    // backward substitution (row oriented): x[i] -= val[j] * x[row[j]]
#pragma unroll 1
    for (int i = nrow - 1; i >= 0; --i) {
        const int jend = L_col[i + 1];
        real xi = s.ld(off + i);
#pragma unroll 1
        for (int j = L_col[i]; j < jend; ++j) xi = A::nmsub(xi, L_val[j], s.ld(off + L_row[j]));
        s.st(off + i, xi);
    }
}

// projection onto ||x[1:]|| <= x[0]
template <class A, int DIMS, typename real> __device__ __forceinline__ void proj_soc(real (&x)[DIMS]) {
    real nrm = real(0);
#pragma unroll
    for (int j = 1; j < DIMS; ++j) nrm = A::madd(nrm, x[j], x[j]);
    nrm = A::sqrt(nrm);
    if (nrm <= x[0]) {
    } else if (nrm <= -x[0]) {
#pragma unroll
        for (int j = 0; j < DIMS; ++j) x[j] = real(0);
    } else {
        const real step = A::div(A::add(x[0], nrm), A::mul(real(2), nrm));
        x[0] = A::mul(step, nrm);
#pragma unroll
        for (int j = 1; j < DIMS; ++j) x[j] = A::mul(step, x[j]);
    }
}

// projection onto the shifted cone ||x[1:3]|| <= alpha (x[0] - d), alpha = +-1
template <class A, typename real> __device__ __forceinline__ void proj_soc3(real (&x)[3], real alpha, real d) {
    const real x_0 = x[0];
    real nrm = real(0);
#pragma unroll
    for (int j = 1; j < 3; ++j) nrm = A::madd(nrm, x[j], x[j]);
    nrm = A::sqrt(nrm);
    const real c0 = A::mul(alpha, A::sub(x_0, d));
    if (nrm <= c0) {
    } else if (nrm <= -c0) {
        x[0] = d;
        x[1] = real(0);
        x[2] = real(0);
    } else {
        const real step = A::div(A::add(c0, nrm), A::mul(real(2), nrm));
        x[0] = A::add(A::mul(A::mul(step, nrm), alpha), d);
        x[1] = A::mul(step, x[1]);
        x[2] = A::mul(step, x[2]);
    }
}

template <class A, typename real> __device__ __forceinline__ void proj_diamond3(real (&x)[3], real lb, real ub) {
    proj_soc3<A>(x, real(1), lb);
    proj_soc3<A>(x, real(-1), ub);
}

}  // namespace spcies
