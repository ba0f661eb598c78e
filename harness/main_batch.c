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
 * 3. a batched closed loop: x+ = A x + B u for every instance between solves (main_cl_in_C.c:100-117).
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

    /* ---- 3. batched closed loop */
    for (int t = 1; t <= steps; t++) {
        rc = BATCH_FN(B, X, XR, UR, U, K, E, NULL, NULL, &info);
        if (rc) return 1;
        double xn[nn_], dist = 0.0;
        for (long i = 0; i < B; i++) {
            for (int r = 0; r < nn_; r++) {
                xn[r] = 0.0;
                for (int c = 0; c < nn_; c++) xn[r] += AB[r][c] * X[i * nn_ + c];
                for (int c = 0; c < mm_; c++) xn[r] += AB[r][nn_ + c] * U[i * mm_ + c];
            }
            memcpy(&X[i * nn_], xn, sizeof xn);
            for (int r = 0; r < 3; r++) dist += fabs(xn[r] - xr[r]);
        }
        printf("closed loop step %d: mean |x_pos - xr| = %.4f, mean k = %.1f\n", t, dist / (3.0 * B), (double)info.sum_k / B);
    }
    spcies_cuda_free();
    return 0;
}
