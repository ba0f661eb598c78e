/*
 * spcies_cuda_batch_mex.c -- MEX gateway of the batched CUDA entry point.
 *
 *   [u_opt, k, e_flag, info] = <name>_batch(x0, xr, ur)            x0, xr: nn_ x B; ur: mm_ x B
 *   [u_opt, k, e_flag, info] = <name>_batch(x0, xr, ur, r)         ellipMPC_ADMM_soc (r: 1 x B or scalar)
 *   [...] = <name>_batch(..., opts)                                 struct: device, n_devices, exact, LB, UB (nm_ x B)
 *
 * Counterpart of the reference's single-instance gateways (formulations/+laxMPC/struct_laxMPC_FISTA_C_Matlab.c:8-167):
 * same argument checks and message identifiers ("Spcies:<F>:nrhs:..."), k and e_flag returned as doubles (:149-150).
 * One MATLAB column is one instance, so the mxArray data pointers are passed to the C ABI without any copy
 * (include/spcies_cuda.h: arrays are instance-major).
 *
 * Build (MATLAB, after the generator wrote <name>.cu/.h and nvcc built <name>.so):
 *   mex -silent -I<dir> -I<spcies-b200>/include spcies_cuda_batch_mex.c <dir>/<name>.so ...
 *       -DSPCIES_HDR='"<name>.h"' -DSPCIES_FUNC=<func> -DSPCIES_HAS_R=0 -output <name>_batch
 * Not compiled in this repository's environment (no MATLAB / mex.h).
 */
#include "mex.h"
#include SPCIES_HDR
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define BATCH_FN CAT(SPCIES_FUNC, _batch)

static double opt_scalar(const mxArray *s, const char *f, double dflt) {
    const mxArray *v = s ? mxGetField(s, 0, f) : NULL;
    return v ? mxGetScalar(v) : dflt;
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    const int nin = 3 + SPCIES_HAS_R;
    if (nrhs != nin && nrhs != nin + 1)
        mexErrMsgIdAndTxt("Spcies:CUDA:nrhs", "%d inputs required (plus an optional options struct)", nin);
    for (int a = 0; a < nin; a++)
        if (!mxIsDouble(prhs[a]) || mxIsComplex(prhs[a]))
            mexErrMsgIdAndTxt("Spcies:CUDA:notDouble", "Inputs must be real double arrays");
    const mwSize B = mxGetN(prhs[0]);
    if (mxGetM(prhs[0]) != nn_) mexErrMsgIdAndTxt("Spcies:CUDA:nrhs:x0", "x0 must be nn_ x B (nn_ = %d)", nn_);
    if (mxGetM(prhs[1]) != nn_ || mxGetN(prhs[1]) != B) mexErrMsgIdAndTxt("Spcies:CUDA:nrhs:xr", "xr must be nn_ x B");
    if (mxGetM(prhs[2]) != mm_ || mxGetN(prhs[2]) != B) mexErrMsgIdAndTxt("Spcies:CUDA:nrhs:ur", "ur must be mm_ x B");
    const mxArray *mopts = (nrhs == nin + 1) ? prhs[nin] : NULL;
    if (mopts && !mxIsStruct(mopts)) mexErrMsgIdAndTxt("Spcies:CUDA:opts", "the last argument must be a struct");

    spcies_batch_opts opts;
    spcies_batch_info info;
    memset(&opts, 0, sizeof opts);
    opts.device = (int)opt_scalar(mopts, "device", 0);
    opts.n_devices = (int)opt_scalar(mopts, "n_devices", 1);
    opts.arith = opt_scalar(mopts, "exact", 0) != 0 ? SPCIES_CUDA_ARITH_EXACT : SPCIES_CUDA_ARITH_FAST;
    opts.engine = (int)opt_scalar(mopts, "engine", SPCIES_CUDA_ENGINE_AUTO);   /* 0 auto | 1 scalar | 2 tensor-core (MMA) */
    const mxArray *LB = mopts ? mxGetField(mopts, 0, "LB") : NULL, *UB = mopts ? mxGetField(mopts, 0, "UB") : NULL;
    if (LB && UB) {
        if (mxGetM(LB) != nm_ || mxGetN(LB) != B || mxGetM(UB) != nm_ || mxGetN(UB) != B)
            mexErrMsgIdAndTxt("Spcies:CUDA:bounds", "opts.LB and opts.UB must be nm_ x B");
        opts.LB = mxGetPr(LB);
        opts.UB = mxGetPr(UB);
    }
    plhs[0] = mxCreateDoubleMatrix(mm_, B, mxREAL);
    int *k = (int *)mxMalloc(sizeof(int) * (B ? B : 1)), *e = (int *)mxMalloc(sizeof(int) * (B ? B : 1));
#if SPCIES_HAS_R
    double *r = mxGetPr(prhs[3]), *rfull = NULL;
    if (mxGetNumberOfElements(prhs[3]) == 1 && B != 1) {      /* scalar r: broadcast */
        rfull = (double *)mxMalloc(sizeof(double) * B);
        for (mwSize i = 0; i < B; i++) rfull[i] = r[0];
        r = rfull;
    } else if (mxGetNumberOfElements(prhs[3]) != B)
        mexErrMsgIdAndTxt("Spcies:CUDA:nrhs:r", "r must be a scalar or have B elements");
    int rc = BATCH_FN((long)B, mxGetPr(prhs[0]), mxGetPr(prhs[1]), mxGetPr(prhs[2]), r, mxGetPr(plhs[0]), k, e, NULL, &opts, &info);
    if (rfull) mxFree(rfull);
#else
    int rc = BATCH_FN((long)B, mxGetPr(prhs[0]), mxGetPr(prhs[1]), mxGetPr(prhs[2]), mxGetPr(plhs[0]), k, e, NULL, &opts, &info);
#endif
    if (rc != 0) mexErrMsgIdAndTxt("Spcies:CUDA:device", "%s (no CPU fallback)", spcies_cuda_last_error());
    if (nlhs > 1) { plhs[1] = mxCreateDoubleMatrix(1, B, mxREAL); for (mwSize i = 0; i < B; i++) mxGetPr(plhs[1])[i] = (double)k[i]; }
    if (nlhs > 2) { plhs[2] = mxCreateDoubleMatrix(1, B, mxREAL); for (mwSize i = 0; i < B; i++) mxGetPr(plhs[2])[i] = (double)e[i]; }
    if (nlhs > 3) {
        const char *f[] = {"kernel_ms", "h2d_ms", "d2h_ms", "total_ms", "sum_k", "n_not_converged"};
        plhs[3] = mxCreateStructMatrix(1, 1, 6, f);
        mxSetField(plhs[3], 0, "kernel_ms", mxCreateDoubleScalar(info.kernel_ms));
        mxSetField(plhs[3], 0, "h2d_ms", mxCreateDoubleScalar(info.h2d_ms));
        mxSetField(plhs[3], 0, "d2h_ms", mxCreateDoubleScalar(info.d2h_ms));
        mxSetField(plhs[3], 0, "total_ms", mxCreateDoubleScalar(info.total_ms));
        mxSetField(plhs[3], 0, "sum_k", mxCreateDoubleScalar((double)info.sum_k));
        mxSetField(plhs[3], 0, "n_not_converged", mxCreateDoubleScalar((double)info.n_not_converged));
    }
    mxFree(k);
    mxFree(e);
}
