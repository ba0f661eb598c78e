/*
 * main_batch.c -- plain-C caller of a generated CUDA solver (the counterpart of the reference's
 * examples/cl_in_C/main_cl_in_C.c:83-131, which calls the generated plain-C solver in a closed loop).
 *
 * Build against any generated solver, e.g. the laxMPC FISTA benchmark configuration:
 *   gcc -O2 harness/main_batch.c -Iinclude -Igenerated_solvers -DSPCIES_HDR='"C2_laxMPC_FISTA.h"' \
 *       -DSPCIES_FUNC=laxMPC_FISTA generated_solvers/C2_laxMPC_FISTA.so -Wl,-rpath,'$ORIGIN/../generated_solvers' \
 *       -lm -o harness/main_batch
 *   ./harness/main_batch [B] [closed-loop steps]
 *
 * 1. single-instance call through the UNCHANGED reference symbol (same lines as main_cl_in_C.c:103), timing p50;
 * 2. batched call: B instances with random x0 around the closed-loop trajectory, solves/s;
 * 3. the batched closed loop on the device: <func>_closed_loop (x+ = A x + B u for every instance between solves,
 *    main_cl_in_C.c:100-117), cold and warm started.
 */
#include SPCIES_HDR
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define BATCH_FN CAT(SPCIES_FUNC, _batch)

static double now_ms(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1e3 * t.tv_sec + 1e-6 * t.tv_nsec;
}
static int cmp_d(const void *a, const void *b) { return (*(const double *)a > *(const double *)b) - (*(const double *)a < *(const double *)b); }

/* discretised 3-mass model, the fixture of examples/cl_in_C/main_cl_in_C.c:96 (rows of [A B]) */
static const double AB[6][8] = {
    {0.921583046607005, 0.038422585681011, 0.000522052604120, 0.194730181566844, 0.002603313047213, 0.000021010412519, 0.019735454526605, 0.000000703027327},
    {0.076845171362023, 0.844737875244983, 0.076845171362023, 0.005206626094427, 0.189523555472418, 0.005206626094427, 0.000264541340786, 0.000264541340786},
    {0.000522052604120, 0.038422585681011, 0.921583046607005, 0.000021010412519, 0.002603313047213, 0.194730181566844, 0.000000703027327, 0.019735454526605},
    {-0.768451713620225, 0.368844000461175, 0.015185150963330, 0.921583046607005, 0.038422585681011, 0.000522052604120, 0.194730181566844, 0.000021010412519},
    {0.737688000922350, -1.506273714321126, 0.737688000922350, 0.076845171362023, 0.844737875244983, 0.076845171362023, 0.005206626094427, 0.005206626094427},
    {0.015185150963330, 0.368844000461175, -0.768451713620225, 0.000522052604120, 0.038422585681011, 0.921583046607005, 0.000021010412519, 0.194730181566844}};

int main(int argc, char **argv) {
    long B = argc > 1 ? atol(argv[1]) : 65536;
    int steps = argc > 2 ? atoi(argv[2]) : 5;
    if (spcies_cuda_device_count() < 1) {
        fprintf(stderr, "no CUDA device: %s has no CPU fallback\n", spcies_cuda_solver_name());
        return 2;
    }
    /* ---- 1. single instance, reference signature */
    double x[nn_] = {0.0}, u[mm_] = {0.0};
    double xr[nn_] = {0.25, 0.25, 0.25, 0.0, 0.0, 0.0}, ur[mm_] = {0.5, 0.5};
    int k, e_flag;
    CAT(sol_, SPCIES_SAVE) sol;
    double lat[200];
    for (int i = 0; i < 200; i++) {
        double t0 = now_ms();
        SPCIES_FUNC(x, xr, ur, u, &k, &e_flag, &sol);
        lat[i] = now_ms() - t0;
    }
    qsort(lat, 200, sizeof(double), cmp_d);
    printf("single solve: u = [%.6f %.6f], k = %d, e_flag = %d, p50 latency = %.1f us\n", u[0], u[1], k, e_flag, 1e3 * lat[100]);

    /* ---- 2. batch */
    double *X = malloc(sizeof(double) * B * nn_), *XR = malloc(sizeof(double) * B * nn_), *UR = malloc(sizeof(double) * B * mm_);
    double *U = malloc(sizeof(double) * B * mm_);
    int *K = malloc(sizeof(int) * B), *E = malloc(sizeof(int) * B);
    srand(1);
    for (long i = 0; i < B; i++) {
        for (int j = 0; j < nn_; j++) { X[i * nn_ + j] = 0.1 * ((double)rand() / RAND_MAX - 0.5); XR[i * nn_ + j] = xr[j]; }
        for (int j = 0; j < mm_; j++) UR[i * mm_ + j] = ur[j];
    }
    spcies_batch_info info;
    int rc = BATCH_FN(B, X, XR, UR, U, K, E, NULL, NULL, &info);   /* warm-up: buffers, constants */
    if (rc) { fprintf(stderr, "batched call failed: %s\n", spcies_cuda_last_error()); return 1; }
    double t0 = now_ms();
    rc = BATCH_FN(B, X, XR, UR, U, K, E, NULL, NULL, &info);
    double dt = now_ms() - t0;
    printf("batch of %ld: %.2f ms end to end (%.2f ms kernel) = %.3g solves/s, mean k = %.1f, not converged = %ld\n", B, dt,
           info.kernel_ms, B / (dt * 1e-3), (double)info.sum_k / B, info.n_not_converged);

    /* ---- 3. batched closed loop ON THE DEVICE: <func>_closed_loop simulates `steps` sampling times of u_t = MPC(x_t),
     *         x_{t+1} = A x_t + B u_t for every instance (main_cl_in_C.c:100-117) without a host round trip per sampling time;
     *         the plant is the generated prediction model (opts.plant_AB = NULL) -- the AB fixture above is the same matrix */
    if (steps > 0) {
        double *XT = malloc(sizeof(double) * (steps + 1) * B * nn_), *UT = malloc(sizeof(double) * steps * B * mm_);
        int *KT = malloc(sizeof(int) * steps * B), *ET = malloc(sizeof(int) * steps * B);
        for (int warm = 0; warm <= 2; warm += 2) {
            spcies_batch_opts o;
            memset(&o, 0, sizeof o);
            o.warm_start = warm;                       /* 0: cold start at every sampling time; 2: shifted dual point of the last one */
            t0 = now_ms();
            rc = CAT(SPCIES_FUNC, _closed_loop)(B, steps, X, XR, UR, XT, UT, KT, ET, &o, &info);
            dt = now_ms() - t0;
            if (rc) { fprintf(stderr, "closed loop failed: %s\n", spcies_cuda_last_error()); return 1; }
            printf("closed loop (%s start), %d steps x %ld instances: %.2f ms (%.2f ms on the device, %ld launch(es)) = %.3g steps/s, mean k = %.1f\n",
                   warm ? "shifted warm" : "cold", steps, B, dt, info.kernel_ms, info.launches, (double)steps * B / (dt * 1e-3),
                   (double)info.sum_k / ((double)steps * B));
        }
        for (int t = 1; t <= steps; t++) {
            double dist = 0.0, err = 0.0;
            long kk = 0;
            for (long i = 0; i < B; i++) {
                const double *xp = XT + ((long)(t - 1) * B + i) * nn_, *xn = XT + ((long)t * B + i) * nn_, *uu = UT + ((long)(t - 1) * B + i) * mm_;
                for (int r = 0; r < nn_; r++) {           /* the device's successor state against the fixture model */
                    double a = 0.0;
                    for (int c = 0; c < nn_; c++) a += AB[r][c] * xp[c];
                    for (int c = 0; c < mm_; c++) a += AB[r][nn_ + c] * uu[c];
                    if (fabs(a - xn[r]) > err) err = fabs(a - xn[r]);
                }
                for (int r = 0; r < 3; r++) dist += fabs(xn[r] - xr[r]);
                kk += KT[(long)(t - 1) * B + i];
            }
            printf("closed loop step %d: mean |x_pos - xr| = %.4f, mean k = %.1f, max |x+ - AB (x; u)| = %.1e\n", t, dist / (3.0 * B),
                   (double)kk / B, err);
        }
        free(XT); free(UT); free(KT); free(ET);
    }
    spcies_cuda_free();
    return 0;
}
