// uldc_probe.cu -- how fast can warp-uniform FP64 constants reach DFMA when they are NOT direct c[bank][imm] operands
// and NOT shared-memory loads?  Variants:
//   ldc     runtime-uniform offset into __constant__ memory (compiler emits LDC / ULDC + DFMA with R / UR operand)
//   lds     LDS.128 broadcast from shared memory (the production path), for reference
//   ldg     uniform global loads through L1 (LDG.E.128 .CONSTANT)
// Each variant sweeps a 10-stage x 144-double table (11.5 KB, the FistaDerived footprint).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(2); } } while (0)

constexpr int STAGES = 10, PER = 144;
__constant__ double ctab[STAGES * PER];

__global__ void k_ldc(double *out, int iters, int stride) {
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll 1
        for (int l = 0; l < STAGES; ++l) {
            const double *c = ctab + l * stride;
#pragma unroll
            for (int k = 0; k < PER; ++k) acc[k % 8] = fma(acc[k % 8], c[k], 1e-3);
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_lds(double *out, int iters, int stride, const double *g) {
    __shared__ __align__(16) double sc[STAGES * PER];
    for (int i = threadIdx.x; i < STAGES * PER; i += blockDim.x) sc[i] = g[i];
    __syncthreads();
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll 1
        for (int l = 0; l < STAGES; ++l) {
            const double2 *c = reinterpret_cast<const double2 *>(sc + l * stride);
#pragma unroll
            for (int k = 0; k < PER / 2; ++k) {
                double2 v = c[k];
                acc[(2 * k) % 8] = fma(acc[(2 * k) % 8], v.x, 1e-3);
                acc[(2 * k + 1) % 8] = fma(acc[(2 * k + 1) % 8], v.y, 1e-3);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ldg(double *out, int iters, int stride, const double *__restrict__ g) {
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll 1
        for (int l = 0; l < STAGES; ++l) {
            const double2 *c = reinterpret_cast<const double2 *>(g + l * stride);
#pragma unroll
            for (int k = 0; k < PER / 2; ++k) {
                double2 v = __ldg(c + k);
                acc[(2 * k) % 8] = fma(acc[(2 * k) % 8], v.x, 1e-3);
                acc[(2 * k + 1) % 8] = fma(acc[(2 * k + 1) % 8], v.y, 1e-3);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// half of the constants through LDS.128 (shared-memory pipe, one per SM), half through LDCU (uniform datapath, one per
// SM sub-partition): do the two delivery paths add up?
template <int NLDC>
__global__ void k_mix(double *out, int iters, int stride, const double *g) {
    __shared__ __align__(16) double sc[STAGES * PER];
    for (int i = threadIdx.x; i < STAGES * PER; i += blockDim.x) sc[i] = g[i];
    __syncthreads();
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll 1
        for (int l = 0; l < STAGES; ++l) {
            const double2 *c = reinterpret_cast<const double2 *>(sc + l * stride);
            const double *cc = ctab + l * stride;
#pragma unroll
            for (int k = 0; k < PER / 2; ++k) {
                if ((k % 4) < NLDC) {
                    acc[(2 * k) % 8] = fma(acc[(2 * k) % 8], cc[2 * k], 1e-3);
                    acc[(2 * k + 1) % 8] = fma(acc[(2 * k + 1) % 8], cc[2 * k + 1], 1e-3);
                } else {
                    double2 v = c[k];
                    acc[(2 * k) % 8] = fma(acc[(2 * k) % 8], v.x, 1e-3);
                    acc[(2 * k + 1) % 8] = fma(acc[(2 * k + 1) % 8], v.y, 1e-3);
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> static double time_ms(F launch) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    std::vector<double> h(STAGES * PER);
    for (size_t i = 0; i < h.size(); ++i) h[i] = 1.0 + 1e-9 * i;
    CK(cudaMemcpyToSymbol(ctab, h.data(), h.size() * 8));
    double *g, *out;
    CK(cudaMalloc(&g, h.size() * 8)); CK(cudaMemcpy(g, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, sizeof(double) * sms * 1024));
    const int iters = 200;
    printf("{");
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int block = warps * 32;
        const double fl = (double)sms * block * iters * STAGES * PER;
        double ms = time_ms([&] { k_ldc<<<sms, block>>>(out, iters, PER); });
        printf("\"ldc_%dw\": %.3f, ", warps, fl / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_lds<<<sms, block>>>(out, iters, PER, g); });
        printf("\"lds_%dw\": %.3f, ", warps, fl / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_ldg<<<sms, block>>>(out, iters, PER, g); });
        printf("\"ldg_%dw\": %.3f, ", warps, fl / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_mix<1><<<sms, block>>>(out, iters, PER, g); });
        printf("\"mix25_%dw\": %.3f, ", warps, fl / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_mix<2><<<sms, block>>>(out, iters, PER, g); });
        printf("\"mix50_%dw\": %.3f, ", warps, fl / (ms * 1e-3) / 1e12);
    }
    printf("\"unit\": \"TFMA/s\"}\n");
    return 0;
}
