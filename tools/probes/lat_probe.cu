// lat_probe.cu -- dependent-issue latencies that bound the one-CTA-per-instance latency engine (development probe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
__global__ void k_dfma(double *out, double a, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) out[1024] = (double)(t1 - t0) / N_IT;
}
__global__ void k_dadd(double *out, double a) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) x = x + a;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) out[1024] = (double)(t1 - t0) / N_IT;
}
__global__ void k_dsetp(double *out, double a, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) x = (x > a) ? x : b + x * 0.0 + a;   // compare + select (+ junk)
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) out[1024] = (double)(t1 - t0) / N_IT;
}
__global__ void k_clip(double *out, double lo, double hi) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) {
        x = (x > lo) ? x : lo;
        x = (x > hi) ? hi : x;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) out[1024] = (double)(t1 - t0) / N_IT;
}
__global__ void k_shfl(double *out) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) out[1024] = (double)(t1 - t0) / N_IT;
}
__global__ void k_lds(double *out) {
    __shared__ double s[1024];
    s[threadIdx.x] = (double)((threadIdx.x * 7 + 1) % 1024);
    __syncthreads();
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) idx = (int)s[idx & 1023];
    long long t1 = clock64();
    out[threadIdx.x] = idx;
    if (threadIdx.x == 0) out[1024] = (double)(t1 - t0) / N_IT;
}
__global__ void k_bar(double *out) {
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[1024] = (double)(t1 - t0) / N_IT;
}
__global__ void k_bar_or(double *out) {
    int p = threadIdx.x == 1000;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) p = __syncthreads_or(p);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[1024] = (double)(t1 - t0) / N_IT + p * 1e-9;
}
__global__ void k_sts_bar_lds(double *out) {
    __shared__ double s[1024];
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N_IT; ++i) {
        s[threadIdx.x] = x;
        __syncthreads();
        x = s[(threadIdx.x + 33) % blockDim.x] + 1.0;
        __syncthreads();
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) out[1024] = (double)(t1 - t0) / N_IT;
}
int main() {
    double *d, h;
    cudaMalloc(&d, 1025 * sizeof(double));
    cudaMemset(d, 0, 1025 * sizeof(double));
    const int blocks[3] = {32, 128, 320};
    for (int b = 0; b < 3; ++b) {
        const int B = blocks[b];
#define RUN(name, ...)                                                   \
    name<<<1, B>>>(__VA_ARGS__);                                          \
    name<<<1, B>>>(__VA_ARGS__);                                          \
    cudaDeviceSynchronize();                                             \
    cudaMemcpy(&h, d + 1024, sizeof h, cudaMemcpyDeviceToHost);          \
    printf("%-16s block %4d: %7.1f cycles per dependent op\n", #name, B, h);
        RUN(k_dfma, d, 1.0000001, 1e-9)
        RUN(k_dadd, d, 1e-9)
        RUN(k_dsetp, d, 0.5, 0.25)
        RUN(k_clip, d, -0.5, 0.5)
        RUN(k_shfl, d)
        RUN(k_lds, d)
        RUN(k_bar, d)
        RUN(k_bar_or, d)
        RUN(k_sts_bar_lds, d)
    }
    return 0;
}
