// tmem_probe.cu -- measures Tensor Memory (TMEM) used as per-lane private scratch (no MMA involved).
//
// The FISTA kernel keeps the per-instance iterates y / lambda in TMEM so that shared memory only has to hold the
// W-solve workspace: 8 instead of 4 resident warps per SM at N = 10.  The access pattern is
// tcgen05.{ld,st}.32x32b.xN: lane i of warp w touches TMEM lane 32*(w%4)+i, columns [c, c+N) -- a private row.
// This probe checks that pattern for 8 warps (two column halves per lane quadrant) and measures
//   * load-to-use latency of tcgen05.ld (+ wait::ld), store->load round trip,
//   * aggregate ld / st throughput per SM at 4 and 8 warps,
// so that the kernel design rests on measured numbers (the vendor text only quotes MMA-side bandwidth).
// Output: one JSON object on stdout.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t *slot) {
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
__device__ __forceinline__ void tmem_free_all(uint32_t base) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

__device__ __forceinline__ void tld2(uint32_t a, double &x0, double &x1) {   // 4 columns = 2 doubles
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    x0 = __hiloint2double((int)r1, (int)r0);
    x1 = __hiloint2double((int)r3, (int)r2);
}
__device__ __forceinline__ void tst2(uint32_t a, double x0, double x1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(__double2loint(x0)), "r"(__double2hiint(x0)),
                 "r"(__double2loint(x1)), "r"(__double2hiint(x1)) : "memory");
}
__device__ __forceinline__ void tld4_nowait(uint32_t a, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tst4(uint32_t a, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// private row of this thread: lane quadrant of the warp, column half by warp group
__device__ __forceinline__ uint32_t my_row(uint32_t base, int cols_per_thread) {
    const int w = threadIdx.x >> 5;
    return base + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)((w >> 2) * cols_per_thread);
}

// ---- 1. correctness: every thread writes 2*ND doubles to its row, all read back with another chunking
template <int ND>
__global__ void probe_check(unsigned long long *errors) {
    __shared__ uint32_t slot;
    const uint32_t base = tmem_alloc_all(&slot);
    const uint32_t row = my_row(base, 2 * ND);
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    for (int e = 0; e < ND; e += 2) tst2(row + 2 * e, (double)gt * 1000.0 + e, (double)gt * 1000.0 + e + 1 + 0.5);
    wait_st();
    __syncthreads();
    unsigned long long bad = 0;
    for (int e = 0; e < ND; e += 4) {
        uint32_t r[8];
        tld4_nowait(row + 2 * e, r);
        wait_ld();
        for (int q = 0; q < 4; ++q) {
            double v = __hiloint2double((int)r[2 * q + 1], (int)r[2 * q]);
            double want = (double)gt * 1000.0 + e + q + ((q & 1) ? 0.5 : 0.0);
            bad += (v != want);
        }
    }
    if (bad) atomicAdd(errors, bad);
    tmem_free_all(base);
}

// ---- 2. latency: dependent chain of loads (address from the loaded value), single warp
__global__ void probe_latency(long long *out, int iters) {
    __shared__ uint32_t slot;
    const uint32_t base = tmem_alloc_all(&slot);
    const uint32_t row = my_row(base, 256);
    for (int e = 0; e < 128; e += 2) tst2(row + 2 * e, 0.0, 0.0);
    wait_st();
    __syncthreads();
    double a, b;
    uint32_t off = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        tld2(row + off, a, b);
        off = (uint32_t)(__double2loint(a) & 0xfc);   // always 0, but data dependent
    }
    long long t1 = clock64();
    // store -> load round trip on the same columns
    double x = 1.0;
    long long t2 = clock64();
    for (int i = 0; i < iters; ++i) {
        tst2(row + 8, x, x);
        wait_st();
        tld2(row + 8, a, b);
        x = a + 1.0;
    }
    long long t3 = clock64();
    if (threadIdx.x == 0) {
        out[0] = (t1 - t0) / iters;
        out[1] = (t3 - t2) / iters;
        out[2] = (long long)(off + (x > 0));
    }
    tmem_free_all(base);
}

// ---- 3. throughput: every warp streams its row with x8 loads (G loads in flight per wait) / x8 stores
template <int G>
__global__ void probe_bw(long long *out, int iters, int do_store) {
    __shared__ uint32_t slot;
    const uint32_t base = tmem_alloc_all(&slot);
    const int cpt = (blockDim.x > 128) ? 256 : 512;
    const uint32_t row = my_row(base, cpt);
    uint32_t acc = 0;
    uint32_t r[G][8];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int q = 0; q < 8; ++q) r[g][q] = threadIdx.x + g + q;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (do_store) {
#pragma unroll
            for (int g = 0; g < G; ++g) tst4(row + ((i * G + g) * 8) % cpt, r[g]);
            wait_st();
        } else {
#pragma unroll
            for (int g = 0; g < G; ++g) tld4_nowait(row + ((i * G + g) * 8) % cpt, r[g]);
            wait_ld();
#pragma unroll
            for (int g = 0; g < G; ++g) acc += r[g][0] ^ r[g][7];
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = acc;
    }
    tmem_free_all(base);
}

int main() {
    CK(cudaSetDevice(0));
    long long *d_out, h_out[4];
    unsigned long long *d_err, h_err = 0;
    CK(cudaMalloc(&d_out, 64));
    CK(cudaMalloc(&d_err, 8));
    CK(cudaMemset(d_err, 0, 8));
    printf("{");
    probe_check<128><<<148, 256>>>(d_err);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h_err, d_err, 8, cudaMemcpyDeviceToHost));
    printf("\"check_8warps_128doubles_errors\": %llu", h_err);
    CK(cudaMemset(d_err, 0, 8));
    probe_check<256><<<148, 128>>>(d_err);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h_err, d_err, 8, cudaMemcpyDeviceToHost));
    printf(", \"check_4warps_256doubles_errors\": %llu", h_err);

    probe_latency<<<1, 32>>>(d_out, 2000);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_out, d_out, 32, cudaMemcpyDeviceToHost));
    printf(", \"ld_x4_latency_cycles\": %lld, \"st_wait_ld_roundtrip_cycles\": %lld", h_out[0], h_out[1]);

    const int iters = 4000;
    for (int warps = 4; warps <= 8; warps *= 2) {
        for (int st = 0; st <= 1; ++st) {
            probe_bw<1><<<148, warps * 32>>>(d_out, iters, st);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h_out, d_out, 16, cudaMemcpyDeviceToHost));
            printf(", \"%s_x8_g1_%dwarps_bytes_per_clk_sm\": %.1f", st ? "st" : "ld", warps, (double)warps * 32 * 32 * iters / (double)h_out[0]);
            probe_bw<4><<<148, warps * 32>>>(d_out, iters, st);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h_out, d_out, 16, cudaMemcpyDeviceToHost));
            printf(", \"%s_x8_g4_%dwarps_bytes_per_clk_sm\": %.1f", st ? "st" : "ld", warps, (double)warps * 32 * 32 * 4 * iters / (double)h_out[0]);
        }
    }
    printf("}\n");
    return 0;
}
