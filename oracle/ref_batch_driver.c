/* TEST INFRASTRUCTURE -- not part of the product path.
 *
 * Batch driver linked next to an instantiated reference solver (oracle/instantiate.py).
 * It only loops over instances and calls the reference's single-instance function with the
 * reference's own signature (e.g. header_laxMPC_FISTA_C.h:26); it contains no solver arithmetic.
 * Optional pthreads split the batch in contiguous slices (one per thread) for the all-cores
 * CPU baseline of SURVEY.md section 8(d).
 *
 * Compile-time parameters (set by instantiate.py):
 *   SPCIES_HDR   "<save_name>.h"      generated header (defines nn_, mm_, ... and sol_<save_name>)
 *   SPCIES_FUNC  e.g. laxMPC_FISTA    the reference solver symbol
 *   SPCIES_SOL   sol_<save_name>
 *   SPCIES_HAS_R 0|1                  1 for ellipMPC_ADMM_soc (extra double *r_ellip input)
 *   SPCIES_NREF  1|3                  3 for ellipHMPC (x_re, x_rs, x_rc, u_re, u_rs, u_rc; header_ellipHMPC_ADMM_C.h:24)
 */
#ifndef SPCIES_NREF
#define SPCIES_NREF 1
#endif
#include SPCIES_HDR
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    long lo, hi;
    const double *x0, *xr, *ur, *r;
    const double *xr2, *xr3, *ur2, *ur3;
    double *u, *sol;
    int *k, *e;
} slice_t;

static const double *g_bounds[2];   /* TIME_VARYING: LB, UB [B][nm_], set by spcies_ref_set_bounds() */
void spcies_ref_set_bounds(const double *LB, const double *UB) { g_bounds[0] = LB; g_bounds[1] = UB; }

long spcies_ref_sol_doubles(void) { return (long)(sizeof(SPCIES_SOL) / sizeof(double)); }

static void *run_slice(void *arg) {
    slice_t *s = (slice_t *)arg;
    const long nsol = spcies_ref_sol_doubles();
    SPCIES_SOL sol;
    double x0[nn_], xr[nn_], ur[mm_], u[mm_];
    for (long i = s->lo; i < s->hi; i++) {
        int k = 0, e = 0;
        memcpy(x0, s->x0 + i * nn_, sizeof x0);
        memcpy(xr, s->xr + i * nn_, sizeof xr);
        memcpy(ur, s->ur + i * mm_, sizeof ur);
        memset(&sol, 0, sizeof sol);
#if TIME_VARYING == 1
        /* per-instance model: A [nn_*nn_], B [nn_*mm_] column-major, Q [nn_], R [mm_], LB / UB [nm_] (code_laxMPC_FISTA_C.c:19) */
        double Ai[nn_ * nn_], Bi[nn_ * mm_], Qi[nn_], Ri[mm_], LBi[nm_], UBi[nm_];
        memcpy(Ai, s->xr2 + i * nn_ * nn_, sizeof Ai);
        memcpy(Bi, s->xr3 + i * nn_ * mm_, sizeof Bi);
        memcpy(Qi, s->ur2 + i * nn_, sizeof Qi);
        memcpy(Ri, s->ur3 + i * mm_, sizeof Ri);
        memcpy(LBi, g_bounds[0] + i * nm_, sizeof LBi);
        memcpy(UBi, g_bounds[1] + i * nm_, sizeof UBi);
        SPCIES_FUNC(x0, xr, ur, Ai, Bi, Qi, Ri, LBi, UBi, u, &k, &e, &sol);
#elif SPCIES_NREF == 3
        double xr2[nn_], xr3[nn_], ur2[mm_], ur3[mm_];
        memcpy(xr2, s->xr2 + i * nn_, sizeof xr2);
        memcpy(xr3, s->xr3 + i * nn_, sizeof xr3);
        memcpy(ur2, s->ur2 + i * mm_, sizeof ur2);
        memcpy(ur3, s->ur3 + i * mm_, sizeof ur3);
        SPCIES_FUNC(x0, xr, xr2, xr3, ur, ur2, ur3, u, &k, &e, &sol);
#elif SPCIES_HAS_R
        double r = s->r[i];
        SPCIES_FUNC(x0, xr, ur, &r, u, &k, &e, &sol);
#else
        SPCIES_FUNC(x0, xr, ur, u, &k, &e, &sol);
#endif
        memcpy(s->u + i * mm_, u, sizeof u);
        s->k[i] = k;
        s->e[i] = e;
        if (s->sol) memcpy(s->sol + i * nsol, &sol, sizeof sol);
    }
    return NULL;
}

static const double *g_extra[4];   /* set by spcies_ref_set_refs() before spcies_ref_batch(): x_rs, x_rc, u_rs, u_rc */
void spcies_ref_set_refs(const double *xr2, const double *xr3, const double *ur2, const double *ur3) {
    g_extra[0] = xr2; g_extra[1] = xr3; g_extra[2] = ur2; g_extra[3] = ur3;
}

int spcies_ref_batch(long B, const double *x0, const double *xr, const double *ur, const double *r,
                     double *u, int *k, int *e, double *sol, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((long)nthreads > B) nthreads = B > 0 ? (int)B : 1;
    pthread_t th[256];
    slice_t sl[256];
    long per = (B + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        long lo = t * per, hi = lo + per;
        if (lo > B) lo = B;
        if (hi > B) hi = B;
        sl[t] = (slice_t){lo, hi, x0, xr, ur, r, g_extra[0], g_extra[1], g_extra[2], g_extra[3], u, sol, k, e};
    }
    if (nthreads == 1) { run_slice(&sl[0]); return 0; }
    for (int t = 0; t < nthreads; t++)
        if (pthread_create(&th[t], NULL, run_slice, &sl[t]) != 0) return -1;
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}

/* Closed loop of reference calls for every instance (examples/cl_in_C/main_cl_in_C.c:100-117): u_t = solver(x_t), then the
 * successor state accumulated exactly like the example (x_aux[i] += AB[i][j] * x[j]; ... += AB[i][nn_+j] * u[j]).
 * x_traj [steps + 1][B][nn_], u_traj [steps][B][mm_], k_traj / e_traj [steps][B]; AB row-major [nn_][nn_ + mm_]. */
#if SPCIES_NREF == 1 && TIME_VARYING != 1
typedef struct {
    long lo, hi, B;
    int steps;
    const double *x0, *xr, *ur, *r, *AB;
    double *xt, *ut;
    int *kt, *et;
} cl_slice_t;

static void *run_cl_slice(void *arg) {
    cl_slice_t *s = (cl_slice_t *)arg;
    SPCIES_SOL sol;
    double x[nn_], xr[nn_], ur[mm_], u[mm_];
    for (long i = s->lo; i < s->hi; i++) {
        memcpy(x, s->x0 + i * nn_, sizeof x);
        memcpy(s->xt + i * nn_, x, sizeof x);
        for (int t = 0; t < s->steps; t++) {
            int k = 0, e = 0;
            double xin[nn_];
            memcpy(xin, x, sizeof x);
            memcpy(xr, s->xr + i * nn_, sizeof xr);
            memcpy(ur, s->ur + i * mm_, sizeof ur);
            memset(&sol, 0, sizeof sol);
#if SPCIES_HAS_R
            double r = s->r[i];
            SPCIES_FUNC(xin, xr, ur, &r, u, &k, &e, &sol);
#else
            SPCIES_FUNC(xin, xr, ur, u, &k, &e, &sol);
#endif
            double x_aux[nn_] = {0.0};
            for (int a = 0; a < nn_; a++) {
                for (int j = 0; j < nn_; j++) x_aux[a] += s->AB[a * (nn_ + mm_) + j] * x[j];
                for (int j = 0; j < mm_; j++) x_aux[a] += s->AB[a * (nn_ + mm_) + nn_ + j] * u[j];
            }
            memcpy(x, x_aux, sizeof x);
            memcpy(s->ut + ((long)t * s->B + i) * mm_, u, sizeof u);
            s->kt[(long)t * s->B + i] = k;
            s->et[(long)t * s->B + i] = e;
            memcpy(s->xt + ((long)(t + 1) * s->B + i) * nn_, x, sizeof x);
        }
    }
    return NULL;
}

int spcies_ref_closed_loop(long B, int steps, const double *x0, const double *xr, const double *ur, const double *r,
                           const double *AB, double *xt, double *ut, int *kt, int *et, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((long)nthreads > B) nthreads = B > 0 ? (int)B : 1;
    pthread_t th[256];
    cl_slice_t sl[256];
    long per = (B + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        long lo = t * per, hi = lo + per;
        if (lo > B) lo = B;
        if (hi > B) hi = B;
        sl[t] = (cl_slice_t){lo, hi, B, steps, x0, xr, ur, r, AB, xt, ut, kt, et};
    }
    if (nthreads == 1) { run_cl_slice(&sl[0]); return 0; }
    for (int t = 0; t < nthreads; t++)
        if (pthread_create(&th[t], NULL, run_cl_slice, &sl[t]) != 0) return -1;
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}
#endif
