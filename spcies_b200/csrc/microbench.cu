// microbench.cu -- measured on-chip ceilings for the roofline of the solver kernels.
//
// MEASURED_PEAKS.json holds HBM GB/s and bf16 tensor TFLOP/s only; the batched MPC solvers are bound by
// FP64 (or FP32) FMA issue and by shared-memory bandwidth, so those ceilings are
// measured here, on the same GPU, by the same bench run:
//   fp64_fma / fp32_fma   register-only dependent-chain FMA throughput, full chip
//   fp64_fma_lds          DFMA whose second operand comes from an LDS.128 broadcast (constants in smem)
//   fp64_fma_cbank_<KB>   DFMA whose operand is a constant-bank reference sweeping <KB> of __constant__ memory
//   smem_bw               per-thread 64-bit loads of an [element][thread] array (the iterate layout)
// Output: one JSON object on stdout.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(2); } } while (0)

template <typename T, int CHAINS>
__global__ void fma_regs(T *out, int iters, T a, T b) {
    T acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = (T)(threadIdx.x + c);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) acc[c] = acc[c] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA with one operand from shared memory (same address for the whole warp -> broadcast LDS.128)
template <int CHAINS>
__global__ void fma_lds(double *out, int iters, int nconst) {
    extern __shared__ double2 sc[];
    for (int i = threadIdx.x; i < nconst / 2; i += blockDim.x) sc[i] = make_double2(1.0 + 1e-9 * i, 1.0 - 1e-9 * i);
    __syncthreads();
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
        for (int base = 0; base + CHAINS / 2 <= nconst / 2; base += CHAINS / 2) {
#pragma unroll
            for (int c = 0; c < CHAINS / 2; ++c) {
                double2 k = sc[base + c];
                acc[2 * c] = fma(acc[2 * c], k.x, 1e-3);
                acc[2 * c + 1] = fma(acc[2 * c + 1], k.y, 1e-3);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__constant__ double cbank[7680];   // 60 KB

template <int NCONST, int CHAINS>
__global__ void fma_cbank(double *out, int iters) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NCONST; ++k) acc[k % CHAINS] = fma(acc[k % CHAINS], cbank[k], 1e-3);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ELEMS>
__global__ void smem_stream(double *out, int iters) {
    extern __shared__ double st[];
    for (int e = 0; e < ELEMS; ++e) st[e * blockDim.x + threadIdx.x] = e + threadIdx.x;
    __syncthreads();
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll 8
        for (int e = 0; e < ELEMS; e += 4) {
            s0 += st[(e + 0) * blockDim.x + threadIdx.x];
            s1 += st[(e + 1) * blockDim.x + threadIdx.x];
            s2 += st[(e + 2) * blockDim.x + threadIdx.x];
            s3 += st[(e + 3) * blockDim.x + threadIdx.x];
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

// FP64 tensor-core issue rate: m8n8k4 MMA (SASS DMMA.8x8x4, 256 FMA per warp instruction), 4 independent accumulators
__global__ void dmma_regs(double *out, int iters) {
    double a = 1e-3 * threadIdx.x, b = 2e-3 * threadIdx.x, d[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j][0] = d[j][1] = j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(d[j][0]), "+d"(d[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += d[j][0] + d[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> static double time_ms(F launch, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    launch();
    launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

template <int NCONST> static double run_cbank(double *d_out, int sms, int block, int iters) {
    double ms = time_ms([&] { fma_cbank<NCONST, 8><<<sms, block>>>(d_out, iters); });
    return (double)sms * block * iters * NCONST / (ms * 1e-3) / 1e12;
}

int main(int argc, char **argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    const int sms = p.multiProcessorCount;
    double *d_out;
    CK(cudaMalloc(&d_out, sizeof(double) * sms * 8 * 1024));
    std::vector<double> hc(7680);
    for (size_t i = 0; i < hc.size(); ++i) hc[i] = 1.0 + 1e-9 * i;
    CK(cudaMemcpyToSymbol(cbank, hc.data(), sizeof(double) * hc.size()));

    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    {   // FP64 / FP32 register-only FMA, 1024 threads/SM x 8 CTAs/SM worth of grid
        const int block = 256, grid = sms * 8, iters = 4096;
        double ms = time_ms([&] { fma_regs<double, 8><<<grid, block>>>(d_out, iters, 1.0000001, 1e-7); });
        double tf = (double)grid * block * iters * 64.0 / (ms * 1e-3) / 1e12;
        printf(", \"fp64_tfma_per_s\": %.3f, \"fp64_tflops\": %.3f", tf, 2 * tf);
        ms = time_ms([&] { fma_regs<float, 8><<<grid, block>>>((float *)d_out, iters, 1.0000001f, 1e-7f); });
        tf = (double)grid * block * iters * 64.0 / (ms * 1e-3) / 1e12;
        printf(", \"fp32_tfma_per_s\": %.3f, \"fp32_tflops\": %.3f", tf, 2 * tf);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {   // occupancy the solvers actually run at
        const int block = warps * 32, grid = sms, iters = 8192;
        double ms = time_ms([&] { fma_regs<double, 8><<<grid, block>>>(d_out, iters, 1.0000001, 1e-7); });
        printf(", \"fp64_tfma_per_s_%dwarps\": %.3f", warps, (double)grid * block * iters * 64.0 / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { fma_regs<double, 2><<<grid, block>>>(d_out, iters, 1.0000001, 1e-7); });
        printf(", \"fp64_tfma_per_s_%dwarps_ilp2\": %.3f", warps, (double)grid * block * iters * 16.0 / (ms * 1e-3) / 1e12);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int block = warps * 32, grid = sms, iters = 64, nconst = 768;
        double ms = time_ms([&] { fma_lds<8><<<grid, block, nconst * 8>>>(d_out, iters, nconst); });
        printf(", \"fp64_tfma_per_s_lds_%dwarps\": %.3f", warps,
               (double)grid * block * iters * (nconst / 8 * 8) / (ms * 1e-3) / 1e12);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int block = warps * 32, iters = 64;
        printf(", \"fp64_tfma_per_s_cbank2KB_%dwarps\": %.3f", warps, run_cbank<256>(d_out, sms, block, iters * 4));
        printf(", \"fp64_tfma_per_s_cbank6KB_%dwarps\": %.3f", warps, run_cbank<768>(d_out, sms, block, iters));
        printf(", \"fp64_tfma_per_s_cbank13KB_%dwarps\": %.3f", warps, run_cbank<1664>(d_out, sms, block, iters));
        printf(", \"fp64_tfma_per_s_cbank40KB_%dwarps\": %.3f", warps, run_cbank<5120>(d_out, sms, block, iters / 2));
    }
    {   // DMMA: the ceiling of the tensor-core engine (MPC_FISTA_mma.cuh); shares the FP64 datapath with DFMA
        const int block = 256, iters = 2000;
        double ms = time_ms([&] { dmma_regs<<<sms, block>>>(d_out, iters); });
        printf(", \"fp64_dmma_tfma_per_s\": %.3f", (double)sms * (block / 32) * iters * 32.0 * 256.0 / (ms * 1e-3) / 1e12);
    }
    {   // shared-memory streaming in the iterate layout: 128 threads x 200 doubles (the FISTA N=10 footprint)
        const int block = 128, elems = 200, iters = 2000;
        auto k = smem_stream<200>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, block * elems * 8));
        double ms = time_ms([&] { k<<<sms, block, block * elems * 8>>>(d_out, iters); });
        printf(", \"smem_read_GBps\": %.1f", (double)sms * block * elems * 8.0 * iters / (ms * 1e-3) / 1e9);
    }
    int clk = 0;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    printf(", \"sm_clock_khz_nominal\": %d}\n", clk);
    return 0;
}
