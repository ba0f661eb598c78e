// spcies_entry.cuh -- the extern "C" surface of a generated solver library (include/spcies_cuda.h).
// Included at the end of a kernel header, after SPCIES_TRAITS is defined.  The emitted .cu defines
//   SPCIES_FUNC          reference function name (token), e.g. laxMPC_FISTA
//   SPCIES_SOL_T         generated sol_<save_name> type
//   SPCIES_SOLVER_STR    "<F>_<method>[_<sub>]"
//   SPCIES_SAVE_NAME_STR "<save_name>"
//   SPCIES_PRECISION_STR "double" | "float"   declared type of the generated constants (options.precision)
//   SPCIES_ARITH_STR     "double" | "float"   arithmetic type of the kernels
//   SPCIES_HAS_R         0 | 1
#pragma once

#define SPCIES_CAT_(a, b) a##b
#define SPCIES_CAT(a, b) SPCIES_CAT_(a, b)

namespace spcies {
typedef Runtime<SPCIES_TRAITS> RT;

// The reference's single-instance call = a batch of one.  Timing fields keep the reference's meaning and unit
// (milliseconds, docs/timing.md:9-22): update_time = host->device, solve_time = kernel, polish_time = device->host.
static inline void single_instance(double *x0_in, double *xr_in, double *ur_in, double *r_ellip, double *u_opt,
                                   int *k_in, int *e_flag, SPCIES_SOL_T *sol, const double *const *extra = nullptr,
                                   const spcies_batch_opts *opts = nullptr) {
    spcies_batch_info info;
#ifdef DEBUG
    SPCIES_SOL_T tmp;
#endif
    int rc = RT::get().run(1, x0_in, xr_in, ur_in, r_ellip, u_opt, k_in, e_flag,
#ifdef DEBUG
                           sol ? reinterpret_cast<double *>(&tmp) : nullptr,
#else
                           nullptr,
#endif
                           opts, &info, extra);
    if (rc != 0) {
        if (e_flag) *e_flag = SPCIES_CUDA_EFLAG_DEVICE;
        if (k_in) *k_in = 0;
        fprintf(stderr, "spcies_cuda: %s\n", g_last_error);
        return;
    }
    if (sol) {
#ifdef DEBUG
        *sol = tmp;
#endif
#if MEASURE_TIME == 1
        sol->update_time = info.h2d_ms;
        sol->solve_time = info.kernel_ms;
        sol->polish_time = info.d2h_ms;
        sol->run_time = info.total_ms;
#endif
    }
}
}  // namespace spcies

// Every generated library is built with -fvisibility=hidden -fno-gnu-unique: only the C ABI below is exported, and
// template statics / inline functions are NOT unified across libraries.  (Without this, two generated solvers loaded
// into one process would share one Runtime<> singleton -- and with it the first library's device constants.)
#pragma GCC visibility push(default)
extern "C" {

int spcies_cuda_abi_version(void) { return SPCIES_CUDA_ABI_VERSION; }
const char *spcies_cuda_solver_name(void) { return SPCIES_SOLVER_STR; }
const char *spcies_cuda_save_name(void) { return SPCIES_SAVE_NAME_STR; }
const char *spcies_cuda_precision(void) { return SPCIES_PRECISION_STR; }
const char *spcies_cuda_arithmetic(void) { return SPCIES_ARITH_STR; }
int spcies_cuda_dims(int *nn, int *mm, int *NN) {
    if (nn) *nn = nn_;
    if (mm) *mm = mm_;
    if (NN) *NN = NN_;
    return 0;
}
long spcies_cuda_sol_doubles(void) { return (long)(sizeof(SPCIES_SOL_T) / sizeof(double)); }
int spcies_cuda_device_count(void) { return ::spcies::RT::get().device_count(); }
void spcies_cuda_free(void) { ::spcies::RT::get().free_all(); }
const char *spcies_cuda_last_error(void) { return ::spcies::g_last_error_global; }

int spcies_cuda_kernel_attributes(int arith, int *regs, int *smem_static, int *smem_dynamic, int *block_threads,
                                  int *local_bytes) {
    if (::spcies::RT::get().device_count() <= 0)
        return ::spcies::fail(SPCIES_CUDA_ENODEVICE, "no usable CUDA device (this library has no CPU fallback)");
    cudaFuncAttributes fa;
    cudaError_t e = SPCIES_TRAITS::attributes(arith, false, &fa);
    if (e != cudaSuccess) return ::spcies::fail((int)e, "cudaFuncGetAttributes");
    const int blk = SPCIES_TRAITS::default_block(false);
    if (regs) *regs = fa.numRegs;
    if (smem_static) *smem_static = (int)fa.sharedSizeBytes;
    if (smem_dynamic) *smem_dynamic = (int)SPCIES_TRAITS::smem_bytes(blk, false);
    if (block_threads) *block_threads = blk;
    if (local_bytes) *local_bytes = (int)fa.localSizeBytes;
    return 0;
}

#if defined(TIME_VARYING) && TIME_VARYING == 1
// per-instance model (options.time_varying): the reference signature of code_laxMPC_FISTA_C.c:19 -- A_in, B_in column-major,
// Q_in, R_in the diagonals, LB_in / UB_in = [LBx; LBu] -- and its batched form (one model per instance)
void SPCIES_FUNC(double *x0_in, double *xr_in, double *ur_in, double *A_in, double *B_in, double *Q_in, double *R_in, double *LB_in,
                 double *UB_in, double *u_opt, int *k_in, int *e_flag, SPCIES_SOL_T *sol) {
    const double *extra[4] = {A_in, B_in, Q_in, R_in};
    spcies_batch_opts o;
    memset(&o, 0, sizeof o);
    o.LB = LB_in;
    o.UB = UB_in;
    ::spcies::single_instance(x0_in, xr_in, ur_in, nullptr, u_opt, k_in, e_flag, sol, extra, &o);
}
int SPCIES_CAT(SPCIES_FUNC, _batch)(long B, const double *x0, const double *xr, const double *ur, const double *A, const double *Bm,
                                    const double *Q, const double *R, const double *LB, const double *UB, double *u_opt, int *k,
                                    int *e_flag, SPCIES_SOL_T *sol, const spcies_batch_opts *opts, spcies_batch_info *info) {
    const double *extra[4] = {A, Bm, Q, R};
    spcies_batch_opts o;
    memset(&o, 0, sizeof o);
    if (opts) o = *opts;
    o.LB = LB;
    o.UB = UB;
    return ::spcies::RT::get().run(B, x0, xr, ur, nullptr, u_opt, k, e_flag, reinterpret_cast<double *>(sol), &o, info, extra);
}
#elif defined(SPCIES_NREF) && SPCIES_NREF == 3
// three references (x_re, x_rs, x_rc, u_re, u_rs, u_rc): header_ellipHMPC_ADMM_C.h:24
void SPCIES_FUNC(double *x0_in, double *xre_in, double *xrs_in, double *xrc_in, double *ure_in, double *urs_in, double *urc_in,
                 double *u_opt, int *k_in, int *e_flag, SPCIES_SOL_T *sol) {
    const double *extra[4] = {xrs_in, xrc_in, urs_in, urc_in};
    ::spcies::single_instance(x0_in, xre_in, ure_in, nullptr, u_opt, k_in, e_flag, sol, extra);
}
int SPCIES_CAT(SPCIES_FUNC, _batch)(long B, const double *x0, const double *xre, const double *xrs, const double *xrc,
                                    const double *ure, const double *urs, const double *urc, double *u_opt, int *k, int *e_flag,
                                    SPCIES_SOL_T *sol, const spcies_batch_opts *opts, spcies_batch_info *info) {
    const double *extra[4] = {xrs, xrc, urs, urc};
    return ::spcies::RT::get().run(B, x0, xre, ure, nullptr, u_opt, k, e_flag, reinterpret_cast<double *>(sol), opts, info, extra);
}
#elif SPCIES_HAS_R
void SPCIES_FUNC(double *x0_in, double *xr_in, double *ur_in, double *r_ellip, double *u_opt, int *k_in, int *e_flag,
                 SPCIES_SOL_T *sol) {
    ::spcies::single_instance(x0_in, xr_in, ur_in, r_ellip, u_opt, k_in, e_flag, sol);
}
int SPCIES_CAT(SPCIES_FUNC, _batch)(long B, const double *x0, const double *xr, const double *ur, const double *r_ellip,
                                    double *u_opt, int *k, int *e_flag, SPCIES_SOL_T *sol,
                                    const spcies_batch_opts *opts, spcies_batch_info *info) {
    return ::spcies::RT::get().run(B, x0, xr, ur, r_ellip, u_opt, k, e_flag, reinterpret_cast<double *>(sol), opts, info);
}
int SPCIES_CAT(SPCIES_FUNC, _closed_loop)(long B, int steps, const double *x0, const double *xr, const double *ur, const double *r_ellip,
                                          double *x_traj, double *u_traj, int *k_traj, int *e_traj, const spcies_batch_opts *opts,
                                          spcies_batch_info *info) {
    return ::spcies::RT::get().run_closed_loop(B, steps, x0, xr, ur, r_ellip, x_traj, u_traj, k_traj, e_traj, opts, info, spcies_model_AB);
}
#else
void SPCIES_FUNC(double *x0_in, double *xr_in, double *ur_in, double *u_opt, int *k_in, int *e_flag, SPCIES_SOL_T *sol) {
    ::spcies::single_instance(x0_in, xr_in, ur_in, nullptr, u_opt, k_in, e_flag, sol);
}
int SPCIES_CAT(SPCIES_FUNC, _batch)(long B, const double *x0, const double *xr, const double *ur, double *u_opt, int *k,
                                    int *e_flag, SPCIES_SOL_T *sol, const spcies_batch_opts *opts,
                                    spcies_batch_info *info) {
    return ::spcies::RT::get().run(B, x0, xr, ur, nullptr, u_opt, k, e_flag, reinterpret_cast<double *>(sol), opts, info);
}
int SPCIES_CAT(SPCIES_FUNC, _closed_loop)(long B, int steps, const double *x0, const double *xr, const double *ur, double *x_traj,
                                          double *u_traj, int *k_traj, int *e_traj, const spcies_batch_opts *opts,
                                          spcies_batch_info *info) {
    return ::spcies::RT::get().run_closed_loop(B, steps, x0, xr, ur, nullptr, x_traj, u_traj, k_traj, e_traj, opts, info, spcies_model_AB);
}
#endif

}  // extern "C"
#pragma GCC visibility pop
