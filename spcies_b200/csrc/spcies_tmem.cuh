// spcies_tmem.cuh -- Tensor Memory (TMEM, 256 KB per SM on sm_100a) used as per-lane private iterate storage.
//
// No MMA is involved: the solvers are thread-per-instance FP64 code whose residency is capped by the
// per-instance iterates (1.6 KB at N = 10 in shared memory => 4 warps per SM).  TMEM is the only other large
// on-chip store; `tcgen05.{ld,st}.32x32b.xK` moves K consecutive 32-bit columns of ONE TMEM lane to / from K
// registers of ONE thread (lane i of warp w <-> TMEM lane 32*(w%4)+i), i.e. a private row per thread:
//     128 threads: 512 columns = 2 KB per thread;   256 threads: warps w and w+4 split the columns, 1 KB each.
// Measured on B200 (spcies_b200/csrc/tmem_probe.cu, profiles/r1_tmem_probe.json): load-to-use 36 cycles (x4),
// store->wait->load 43 cycles, ~8 cycles per column per warp, 113 B/clk/SM (ld) and 150 B/clk/SM (st) at 8 warps
// -- about the shared-memory figure (108 B/clk/SM), on a separate datapath.
//
// Rules these helpers encode:
//   * every tcgen05 instruction is warp-collective (.sync.aligned): call them from converged code only, with a
//     warp-uniform address (ptxas keeps it in a uniform register: `LDTM.x8 R4, tmem[UR6+0x18]`);
//   * loads are asynchronous: registers are valid after `wait_ld`, which also takes the destination words as
//     read-write operands so that no use can be scheduled above the wait;
//   * stores are asynchronous: `wait_st` before the same columns are loaded again.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spcies {
namespace tmem {

// Allocate all 512 columns for this CTA (one CTA per SM in the persistent kernels).  Contains a __syncthreads.
__device__ __forceinline__ uint32_t alloc_all(uint32_t *smem_slot) {
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(smem_slot))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *smem_slot;
}
// All threads of the CTA must call it (contains a __syncthreads).
__device__ __forceinline__ void free_all(uint32_t base) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

// Address of column 0 of this thread's private row; `cols_per_thread` columns are reserved per warp group.
__device__ __forceinline__ uint32_t my_row(uint32_t base, int cols_per_thread) {
    const uint32_t w = threadIdx.x >> 5;
    return base + (((w & 3u) * 32u) << 16) + (w >> 2) * (uint32_t)cols_per_thread;
}

__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- K-word loads / stores, K in {1, 2, 4, 8, 16}
template <int K> struct Op;
template <> struct Op<1> {
    static __device__ __forceinline__ void ld(uint32_t a, uint32_t *r) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(a) : "memory");
    }
    static __device__ __forceinline__ void st(uint32_t a, const uint32_t *r) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a), "r"(r[0]) : "memory");
    }
};
template <> struct Op<2> {
    static __device__ __forceinline__ void ld(uint32_t a, uint32_t *r) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
    }
    static __device__ __forceinline__ void st(uint32_t a, const uint32_t *r) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(a), "r"(r[0]), "r"(r[1]) : "memory");
    }
};
template <> struct Op<4> {
    static __device__ __forceinline__ void ld(uint32_t a, uint32_t *r) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(a)
                     : "memory");
    }
    static __device__ __forceinline__ void st(uint32_t a, const uint32_t *r) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                     "r"(r[3])
                     : "memory");
    }
};
template <> struct Op<8> {
    static __device__ __forceinline__ void ld(uint32_t a, uint32_t *r) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(a)
                     : "memory");
    }
    static __device__ __forceinline__ void st(uint32_t a, const uint32_t *r) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(r[0]), "r"(r[1]),
                     "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                     : "memory");
    }
};
template <> struct Op<16> {
    static __device__ __forceinline__ void ld(uint32_t a, uint32_t *r) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(a)
            : "memory");
    }
    static __device__ __forceinline__ void st(uint32_t a, const uint32_t *r) {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
            "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
            : "memory");
    }
};

// W consecutive words starting at column address `a`, split greedily into 16/8/4/2/1-word instructions
template <int W> __device__ __forceinline__ void ld_words(uint32_t a, uint32_t *r) {
    if constexpr (W > 0) {
        constexpr int K = W >= 16 ? 16 : W >= 8 ? 8 : W >= 4 ? 4 : W >= 2 ? 2 : 1;
        Op<K>::ld(a, r);
        ld_words<W - K>(a + K, r + K);
    }
}
template <int W> __device__ __forceinline__ void st_words(uint32_t a, const uint32_t *r) {
    if constexpr (W > 0) {
        constexpr int K = W >= 16 ? 16 : W >= 8 ? 8 : W >= 4 ? 4 : W >= 2 ? 2 : 1;
        Op<K>::st(a, r);
        st_words<W - K>(a + K, r + K);
    }
}
// makes every later use of r[0..W) depend on a point after the preceding wait_ld()
template <int W> __device__ __forceinline__ void tie(uint32_t *r) {
#pragma unroll
    for (int i = 0; i < W; ++i) asm volatile("" : "+r"(r[i]));
}

// ---- typed vectors of ND reals (double = 2 words, float = 1 word)
template <typename T> struct Words;
template <> struct Words<double> {
    static constexpr int PER = 2;
    static __device__ __forceinline__ double get(const uint32_t *r, int i) { return __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]); }
    static __device__ __forceinline__ void put(uint32_t *r, int i, double v) {
        r[2 * i] = (uint32_t)__double2loint(v);
        r[2 * i + 1] = (uint32_t)__double2hiint(v);
    }
};
template <> struct Words<float> {
    static constexpr int PER = 1;
    static __device__ __forceinline__ float get(const uint32_t *r, int i) { return __uint_as_float(r[i]); }
    static __device__ __forceinline__ void put(uint32_t *r, int i, float v) { r[i] = __float_as_uint(v); }
};

// An in-flight load of ND reals: issue() ... wait_ld() ... get(v)
template <typename T, int ND> struct Vec {
    static constexpr int W = ND * Words<T>::PER;
    uint32_t r[W];
    __device__ __forceinline__ void issue(uint32_t a) { ld_words<W>(a, r); }
    __device__ __forceinline__ void get(T (&v)[ND]) {   // call after wait_ld()
        tie<W>(r);
#pragma unroll
        for (int i = 0; i < ND; ++i) v[i] = Words<T>::get(r, i);
    }
    static __device__ __forceinline__ void store(uint32_t a, const T (&v)[ND]) {
        uint32_t s[W];
#pragma unroll
        for (int i = 0; i < ND; ++i) Words<T>::put(s, i, v[i]);
        st_words<W>(a, s);
    }
};

}  // namespace tmem
}  // namespace spcies
