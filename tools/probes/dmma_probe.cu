// dmma_probe.cu -- FP64 tensor-core (DMMA, mma.sync .f64) issue rate and dependent latency on sm_100a.
// Question it answers: is the batched shared-matrix product  D[inst][out] += X[inst][k] * M[k][out]  (8 instances per
// warp, shared matrix as the B fragment in registers) cheaper than 1 thread per instance with LDS-broadcast constants?
// Prints one JSON line: TFMA/s (dense slots, no padding discount) per shape / warps-per-SM / independent chains, and
// the latency of a D->C dependent chain and of a D->A dependent chain (the pattern of the solver's stage recursion).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(2); } } while (0)

__device__ __forceinline__ void mma884(double &d0, double &d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__device__ __forceinline__ void mma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                 "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// SHAPE 0: m8n8k4 (256 FMA), 1: m16n8k4 (512), 2: m16n8k8 (1024), 3: m16n8k16 (2048).  ILP independent accumulators.
template <int SHAPE, int ILP>
__global__ void k_tp(double *out, int iters) {
    double a[8], b[4], d[ILP][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1e-3 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (threadIdx.x * 3 + i);
#pragma unroll
    for (int j = 0; j < ILP; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[j][i] = j + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                if (SHAPE == 0) mma884(d[j][0], d[j][1], a[0], b[0], d[j][0], d[j][1]);
                if (SHAPE == 1) { double aa[2] = {a[0], a[1]}; mma1684(d[j], aa, b[0]); }
                if (SHAPE == 2) { double aa[4] = {a[0], a[1], a[2], a[3]}; double bb[2] = {b[0], b[1]}; mma1688(d[j], aa, bb); }
                if (SHAPE == 3) mma16816(d[j], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += d[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// D -> A dependent chain of m8n8k4 pairs (the solver's  mu_l = c_l - F_l mu_{l-1}  recursion: the two result registers
// of one product are the A fragments of the next two k-steps).
__global__ void k_chain_da(double *out, int iters, long long *cycles) {
    double b0 = 1e-3 * threadIdx.x, b1 = 2e-3 * threadIdx.x, d0 = 1.0, d1 = 2.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double e0, e1;
            mma884(e0, e1, d0, b0, 0.5, 0.25);
            mma884(d0, d1, d1, b1, e0, e1);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int SHAPE, int ILP>
static double run_tp(int warps_per_sm, int sms, double *dout) {
    const int iters = 2000;
    const double fma_per = SHAPE == 0 ? 256 : SHAPE == 1 ? 512 : SHAPE == 2 ? 1024 : 2048;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_tp<SHAPE, ILP><<<sms, warps_per_sm * 32>>>(dout, 10);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_tp<SHAPE, ILP><<<sms, warps_per_sm * 32>>>(dout, iters);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double fma = (double)sms * warps_per_sm * iters * 8.0 * ILP * fma_per;
    return fma / (ms * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double *dout;
    CK(cudaMalloc(&dout, sizeof(double) * sms * 1024));
    long long *dcyc, hcyc;
    CK(cudaMalloc(&dcyc, 8));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    const int W[4] = {4, 8, 16, 32};
    for (int wi = 0; wi < 4; ++wi) {
        int w = W[wi];
        printf(", \"m8n8k4_w%d_ilp1\": %.3f", w, run_tp<0, 1>(w, sms, dout));
        printf(", \"m8n8k4_w%d_ilp2\": %.3f", w, run_tp<0, 2>(w, sms, dout));
        printf(", \"m8n8k4_w%d_ilp4\": %.3f", w, run_tp<0, 4>(w, sms, dout));
        printf(", \"m16n8k4_w%d_ilp1\": %.3f", w, run_tp<1, 1>(w, sms, dout));
        printf(", \"m16n8k4_w%d_ilp4\": %.3f", w, run_tp<1, 4>(w, sms, dout));
        printf(", \"m16n8k8_w%d_ilp1\": %.3f", w, run_tp<2, 1>(w, sms, dout));
        printf(", \"m16n8k8_w%d_ilp4\": %.3f", w, run_tp<2, 4>(w, sms, dout));
        printf(", \"m16n8k16_w%d_ilp1\": %.3f", w, run_tp<3, 1>(w, sms, dout));
        printf(", \"m16n8k16_w%d_ilp4\": %.3f", w, run_tp<3, 4>(w, sms, dout));
    }
    k_chain_da<<<1, 32>>>(dout, 1000, dcyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&hcyc, dcyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"m8n8k4_dependent_latency_cycles\": %.1f", (double)hcyc / (1000.0 * 16.0));
    printf("}\n");
    return 0;
}
