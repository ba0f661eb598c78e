/*
 * spcies_cuda.h -- C ABI of the B200 (sm_100a) batched solver backend for Spcies.
 *
 * Every solver generated with platform 'CUDA' is one shared library `<save_name>.so`.  It exports
 *
 *   (1) the reference's single-instance symbol, UNCHANGED (same name, same argument list), so a caller
 *       written for the generated plain-C solver (examples/cl_in_C/main_cl_in_C.c:103, or the MEX gateways
 *       formulations/+<F>/struct_<F>_<method>_C_Matlab.c) links against it as is.  It runs a batch of one;
 *   (2) a batched entry point `<func>_batch` (new): B independent instances that share the generated system
 *       model and differ in x0 / xr / ur (/ r_ellip / bounds);
 *   (3) the common `spcies_cuda_*` query / housekeeping symbols below.
 *
 * Reference interfaces replaced (file:line in the reference tree):
 *   laxMPC_FISTA        formulations/+laxMPC/header_laxMPC_FISTA_C.h:26      (code_laxMPC_FISTA_C.c:21)
 *   laxMPC_ADMM         formulations/+laxMPC/header_laxMPC_ADMM_C.h:27       (code_laxMPC_ADMM_C.c:21)
 *   equMPC_FISTA        formulations/+equMPC/header_equMPC_FISTA_C.h:25      (code_equMPC_FISTA_C.c:21)
 *   equMPC_ADMM         formulations/+equMPC/header_equMPC_ADMM_C.h:26       (code_equMPC_ADMM_C.c:21)
 *   ellipMPC_ADMM       formulations/+ellipMPC/header_ellipMPC_ADMM_C.h      (code_ellipMPC_ADMM_C.c)
 *   ellipMPC_ADMM_soc   formulations/+ellipMPC/header_ellipMPC_ADMM_soc_C.h  (code_ellipMPC_ADMM_soc_C.c:20)
 *   MPCT_EADMM          formulations/+MPCT/header_MPCT_EADMM_C.h             (code_MPCT_EADMM_C.c:18)
 *   MPCT_ADMM_cs        formulations/+MPCT/header_MPCT_ADMM_cs_C.h:25        (code_MPCT_ADMM_cs_C.c:18)
 *   MPCT_ADMM_semiband  formulations/+MPCT/header_MPCT_ADMM_semiband_C.h     (code_MPCT_ADMM_semiband_C.c:21)
 *   HMPC_ADMM           formulations/+HMPC/header_HMPC_ADMM_split_C.h:27     (code_HMPC_ADMM_split_C.c:19)
 *                       formulations/+HMPC/header_HMPC_ADMM_C.h:24           (code_HMPC_ADMM_C.c:18)
 *                       formulations/+HMPC/header_ellipHMPC_ADMM_C.h:24      (code_ellipHMPC_ADMM_C.c:18, six references)
 *
 * Conventions
 *   - all host arrays are row-major, instance-major: x0[B][nn_], xr[B][nn_], ur[B][mm_], r_ellip[B],
 *     u_opt[B][mm_], k[B], e_flag[B].  One column of a MATLAB nn_ x B matrix is one instance, so an mxArray
 *     can be passed without transposition.
 *   - e_flag per instance keeps the reference meaning: 1 converged, -1 k_max reached
 *     (code_laxMPC_FISTA_C.c:354-361).
 *   - the return value (new -- the reference functions are void) is 0 on success or a cudaError_t /
 *     SPCIES_CUDA_E* code for infrastructure errors.  There is no CPU fallback: without a usable CUDA device
 *     every entry point fails with a non-zero code (the single-instance symbol sets *e_flag = SPCIES_CUDA_EFLAG_DEVICE).
 *   - the caller owns every host buffer; the library owns its device buffers (created lazily per device,
 *     released by spcies_cuda_free()).
 *   - calls are blocking; one host thread at a time per library.
 *
 * The per-solver header `<save_name>.h` emitted next to `<save_name>.cu` carries the reference's #defines
 * (nn_, mm_, NN_, k_max, tol, ...), the `sol_<save_name>` struct and the prototypes; this file holds what is
 * common to all of them.
 */
#ifndef SPCIES_CUDA_H
#define SPCIES_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define SPCIES_CUDA_ABI_VERSION 2   /* 2: closed-loop entry points; spcies_batch_opts.warm_start / .plant_AB */

/* infrastructure error codes (positive values below 1000 are cudaError_t) */
#define SPCIES_CUDA_EINVAL      1001   /* bad argument (NULL buffer, B < 0, ...) */
#define SPCIES_CUDA_ENODEVICE   1002   /* no CUDA device / requested device absent */
#define SPCIES_CUDA_EUNSUPPORTED 1003  /* option not available in this generated solver */
#define SPCIES_CUDA_EFLAG_DEVICE (-100) /* e_flag value written by the single-instance symbol on device failure */

/* arithmetic modes */
#define SPCIES_CUDA_ARITH_FAST  0  /* fused multiply-add, reciprocal hoisting: u_opt within 1e-9 (double) of the reference */
#define SPCIES_CUDA_ARITH_EXACT 1  /* same operations, same order, no contraction: bit-identical to gcc -O3 (no -mfma) */

/* Tail handling.  Iteration counts are heavy-tailed; with AUTO, solvers that support it run a large batch as two
 * launches: the first parks the few instances that are still iterating shortly after the instance queue ran dry, the
 * second resumes them with one warp per SM scheduler (much faster per iteration).  Results do not depend on the mode. */
#define SPCIES_CUDA_TAIL_AUTO   0
#define SPCIES_CUDA_TAIL_SINGLE 1  /* always one launch that runs every instance to its end */
#define SPCIES_CUDA_TAIL_TWO_PHASE 2 /* always park & resume (AUTO does so for batches larger than four waves of lanes) */
#define SPCIES_CUDA_TAIL_CAPS   3  /* iteration-cap rounds (MMA engine; AUTO uses them for large batches): launch r runs every
                                    * instance it holds up to tail_caps[r] iterations and parks the rest for launch r+1, the last
                                    * launch runs what is left to the end.  Slow instances are thus found -- and restarted together --
                                    * early, instead of trailing the launch one by one.  Falls back to TWO_PHASE on other engines. */

/* Kernel engine of the FISTA solvers (FAST arithmetic only; every other solver has one engine and ignores the field).
 * MMA: FP64 tensor-core kernel, 8 instances per warp, iterates in registers, shared matrices as MMA fragments.
 * SCALAR: one thread per instance.  AUTO picks MMA when the problem fits it (nn_ + mm_ <= 8, N <= 12, double). */
#define SPCIES_CUDA_ENGINE_AUTO   0
#define SPCIES_CUDA_ENGINE_SCALAR 1
#define SPCIES_CUDA_ENGINE_MMA    2  /* fail with cudaErrorNotSupported instead of falling back to SCALAR */
#define SPCIES_CUDA_ENGINE_SINGLE 3 /* latency engine of the FISTA solvers: one CTA per instance, dense W^-1 (MPC_FISTA_single.cuh);   */
                                    /* AUTO uses it for host-buffer calls of at most 64 instances (the reference's single-instance    */
                                    /* symbol is one); fails with SPCIES_CUDA_EUNSUPPORTED where it cannot run                        */

/* Options of a batched call.  Zero-initialise, then set what you need; NULL means all defaults. */
typedef struct {
    int device;              /* first CUDA device ordinal to use (default 0) */
    int n_devices;           /* shard the batch contiguously over devices device .. device+n_devices-1 (0 or 1: one device) */
    int arith;               /* SPCIES_CUDA_ARITH_FAST (default) | SPCIES_CUDA_ARITH_EXACT */
    int device_pointers;     /* 1: every array argument is a DEVICE pointer on `device` (no copies; n_devices must be <= 1) */
    const double *LB;        /* optional per-instance bounds [B][nm_] = [LBx; LBu] (the TIME_VARYING LB_in layout,  */
    const double *UB;        /*   code_laxMPC_FISTA_C.c:102-105); NULL: the generated constants                      */
    void *stream;            /* cudaStream_t to run on when device_pointers = 1 (NULL: the library's own stream) */
    int block_threads;       /* 0: default chosen at generation time */
    int grid_blocks;         /* 0: one CTA per SM */
    int tail_mode;           /* SPCIES_CUDA_TAIL_AUTO (default) | _SINGLE | _TWO_PHASE, see above */
    int tail_grace;          /* iterations an instance may still run after the queue ran dry before it is parked (0: 32) */
    int engine;              /* SPCIES_CUDA_ENGINE_AUTO (default) | _SCALAR | _MMA | _SINGLE, see above */
    int tail_caps[3];        /* increasing iteration caps of SPCIES_CUDA_TAIL_CAPS, 0-terminated (all 0: 96, 320) */
    int warm_start;          /* closed loop: 0 = cold start; 1 = start every sampling time from the dual point of the previous one  */
                             /*   (FISTA solvers: the `lambda` argument of platforms/Matlab/spcies_laxMPC_FISTA_solver.m:161-164);  */
                             /*   2 = the same, shifted by one stage (the horizon receded: lambda_l <- lambda_{l+1})                */
    int reserved[1];
    const double *plant_AB;  /* closed loop: plant x+ = [A B] (x; u) as a row-major [nn_][nm_] HOST array; NULL: the prediction model */
} spcies_batch_opts;

/* Measurements of the last batched call (all device times from CUDA events on the launching stream). */
typedef struct {
    double kernel_ms;        /* solver kernel only (max over devices)                */
    double h2d_ms, d2h_ms;   /* host<->device copies (0 with device_pointers)        */
    double total_ms;         /* wall clock of the whole call                         */
    long   launches;         /* number of solver-kernel launches                     */
    long   h2d_bytes, d2h_bytes;
    long   sum_k;            /* sum of iteration counts over the batch               */
    long   n_not_converged;  /* instances with e_flag = -1                           */
    int    block_threads, grid_blocks, smem_bytes, regs_per_thread;
    int    n_devices;
    int    drain_us;         /* kernel start -> first lane finds the instance queue empty (us, device globaltimer) */
    int    span_us;          /* kernel start -> last warp leaves (us); span_us - drain_us = tail of the slowest instances */
    int    parked;           /* instances handed from the first to the second launch (tail handling), 0 with one launch */
    int    reserved[4];
} spcies_batch_info;

/* ---- common symbols exported by every generated library ------------------------------------------- */
int         spcies_cuda_abi_version(void);
const char *spcies_cuda_solver_name(void);     /* "<F>_<method>[_<sub>]", e.g. "laxMPC_FISTA"            */
const char *spcies_cuda_save_name(void);       /* the save_name the solver was generated with             */
const char *spcies_cuda_precision(void);       /* "double" | "float": options.precision = the declared type of the generated constants */
const char *spcies_cuda_arithmetic(void);      /* "double" | "float": arithmetic type of the kernels (the reference computes in double  */
                                               /*   whatever the precision, dec_var.m:16-17; "float" only with float_arithmetic)         */
int         spcies_cuda_dims(int *nn, int *mm, int *NN);
long        spcies_cuda_sol_doubles(void);     /* sizeof(sol_<save_name>) / sizeof(double)                */
int         spcies_cuda_device_count(void);    /* 0 if there is no usable device                          */
int         spcies_cuda_kernel_attributes(int arith, int *regs, int *smem_static, int *smem_dynamic,
                                          int *block_threads, int *local_bytes);
void        spcies_cuda_free(void);            /* release device buffers, streams and events              */
const char *spcies_cuda_last_error(void);

/* ---- closed loop (new) -----------------------------------------------------------------------------
 * `<func>_closed_loop` simulates `steps` sampling times of the closed loop  u_t = MPC(x_t),  x_{t+1} = A x_t + B u_t  for B
 * independent instances (the loop of examples/cl_in_C/main_cl_in_C.c:100-117, batched), with constant references.  The state
 * never returns to the host between sampling times; the FISTA tensor-core engine keeps every instance on chip for the whole
 * run (the successor state is one more MMA), the other solvers launch once per sampling time on device-resident arrays.
 *   x_traj [steps + 1][B][nn_] (x_traj[0] = x0; may be NULL), u_traj [steps][B][mm_], k_traj / e_traj [steps][B]. */

/* ---- per-solver entry points ----------------------------------------------------------------------
 * Written as macros because the `sol_<save_name>` type is generated; `<save_name>.h` expands them.
 * NAME is the reference function name, SOL the generated `sol_<save_name>` type.                      */
#define SPCIES_CUDA_DECLARE_SOLVER(NAME, SOL)                                                              \
    void NAME(double *x0_in, double *xr_in, double *ur_in, double *u_opt, int *k_in, int *e_flag, SOL *sol); \
    int NAME##_batch(long B, const double *x0, const double *xr, const double *ur,                         \
                     double *u_opt, int *k, int *e_flag, SOL *sol /* [B] or NULL */,                       \
                     const spcies_batch_opts *opts /* or NULL */, spcies_batch_info *info /* or NULL */);  \
    int NAME##_closed_loop(long B, int steps, const double *x0, const double *xr, const double *ur,       \
                           double *x_traj, double *u_traj, int *k_traj, int *e_traj,                       \
                           const spcies_batch_opts *opts /* or NULL */, spcies_batch_info *info /* or NULL */)

/* ellipMPC_ADMM_soc takes the size of the terminal ellipsoid at run time (code_ellipMPC_ADMM_soc_C.c:20) */
#define SPCIES_CUDA_DECLARE_SOLVER_R(NAME, SOL)                                                            \
    void NAME(double *x0_in, double *xr_in, double *ur_in, double *r_ellip, double *u_opt, int *k_in,      \
              int *e_flag, SOL *sol);                                                                      \
    int NAME##_batch(long B, const double *x0, const double *xr, const double *ur, const double *r_ellip,  \
                     double *u_opt, int *k, int *e_flag, SOL *sol /* [B] or NULL */,                       \
                     const spcies_batch_opts *opts /* or NULL */, spcies_batch_info *info /* or NULL */);  \
    int NAME##_closed_loop(long B, int steps, const double *x0, const double *xr, const double *ur,       \
                           const double *r_ellip, double *x_traj, double *u_traj, int *k_traj, int *e_traj, \
                           const spcies_batch_opts *opts /* or NULL */, spcies_batch_info *info /* or NULL */)

/* ellipHMPC takes three state and three input references, the constant and the two harmonic components of the reference
 * trajectory (header_ellipHMPC_ADMM_C.h:24; code_ellipHMPC_ADMM_C.c:18) */
#define SPCIES_CUDA_DECLARE_SOLVER_6REF(NAME, SOL)                                                         \
    void NAME(double *x0_in, double *xre_in, double *xrs_in, double *xrc_in, double *ure_in, double *urs_in, \
              double *urc_in, double *u_opt, int *k_in, int *e_flag, SOL *sol);                            \
    int NAME##_batch(long B, const double *x0, const double *xre, const double *xrs, const double *xrc,   \
                     const double *ure, const double *urs, const double *urc, double *u_opt, int *k,      \
                     int *e_flag, SOL *sol /* [B] or NULL */, const spcies_batch_opts *opts /* or NULL */,  \
                     spcies_batch_info *info /* or NULL */)

/* TIME_VARYING solvers (options.time_varying, `#define TIME_VARYING 1`): the model is an argument -- A_in [nn_][nn_] and
 * B_in [nn_][mm_] COLUMN-major (as MATLAB passes them), Q_in [nn_] and R_in [mm_] the diagonals of the weights, LB_in / UB_in [nm_]
 * (code_laxMPC_FISTA_C.c:19, :101-137).  Batched: one model per instance, A [B][nn_*nn_], Bm [B][nn_*mm_], Q [B][nn_], R [B][mm_],
 * LB / UB [B][nm_]; the block-Cholesky factorisation of every instance's W runs on the device (:139-262). */
#define SPCIES_CUDA_DECLARE_SOLVER_TV(NAME, SOL)                                                           \
    void NAME(double *x0_in, double *xr_in, double *ur_in, double *A_in, double *B_in, double *Q_in,       \
              double *R_in, double *LB_in, double *UB_in, double *u_opt, int *k_in, int *e_flag, SOL *sol); \
    int NAME##_batch(long B, const double *x0, const double *xr, const double *ur, const double *A,       \
                     const double *Bm, const double *Q, const double *R, const double *LB, const double *UB, \
                     double *u_opt, int *k, int *e_flag, SOL *sol /* [B] or NULL */,                       \
                     const spcies_batch_opts *opts /* or NULL */, spcies_batch_info *info /* or NULL */)

/* The solver families and the symbols each generated library exports (checked by tests/test_abi.py):
 *   SPCIES_CUDA_SOLVER(laxMPC_FISTA)       laxMPC_FISTA        laxMPC_FISTA_batch
 *   SPCIES_CUDA_SOLVER(laxMPC_ADMM)        laxMPC_ADMM         laxMPC_ADMM_batch
 *   SPCIES_CUDA_SOLVER(equMPC_FISTA)       equMPC_FISTA        equMPC_FISTA_batch
 *   SPCIES_CUDA_SOLVER(equMPC_ADMM)        equMPC_ADMM         equMPC_ADMM_batch
 *   SPCIES_CUDA_SOLVER(ellipMPC_ADMM)      ellipMPC_ADMM       ellipMPC_ADMM_batch
 *   SPCIES_CUDA_SOLVER_R(ellipMPC_ADMM_soc) ellipMPC_ADMM_soc  ellipMPC_ADMM_soc_batch
 *   SPCIES_CUDA_SOLVER(MPCT_EADMM)         MPCT_EADMM          MPCT_EADMM_batch
 *   SPCIES_CUDA_SOLVER(MPCT_ADMM_cs)       MPCT_ADMM_cs        MPCT_ADMM_cs_batch
 *   SPCIES_CUDA_SOLVER(MPCT_ADMM_semiband) MPCT_ADMM_semiband  MPCT_ADMM_semiband_batch
 *   SPCIES_CUDA_SOLVER(HMPC_ADMM)          HMPC_ADMM           HMPC_ADMM_batch   (ADMM, ADMM_split, SADMM_split; ellipHMPC: _6REF)
 */

#ifdef __cplusplus
}
#endif
#endif /* SPCIES_CUDA_H */
