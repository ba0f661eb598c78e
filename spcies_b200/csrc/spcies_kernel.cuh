// spcies_kernel.cuh -- the persistent, batch-parallel kernel skeleton shared by the ADMM-family solvers.
//
// One thread owns one MPC instance at a time and pulls the next one from the global queue when it
// finishes (WorkQueue, spcies_common.cuh).  A solver is a policy class `S` that provides
//
//   typedef ... real;                       arithmetic type (double | float)
//   static constexpr int STATE;             per-instance persistent elements (shared or global memory)
//   static constexpr int STATE_VARB;        same with per-instance bounds
//   static constexpr bool HAS_VARB;
//   template <class A, bool VARB, class ST> struct Ctx {            // per-lane context
//       __device__ void init(long long inst);                       // read inputs, zero the iterates
//       __device__ bool iterate(int k);                             // one iteration; true = exit test satisfied
//       __device__ void finish(long long inst, int k, int e_flag);  // write u_opt / k / e_flag (/ sol)
//   };
//
// Where the state lives is decided at compile time from its size:
//   * shared memory, [element][thread], when at least MIN_SMEM_THREADS instances fit next to the constants;
//   * otherwise a global scratch array [element][grid thread] (coalesced; it stays L2 resident for the
//     handful of MB a launch needs) -- the long-horizon (N = 50) configurations.
// The problem constants are staged once per CTA into shared memory with one bulk async copy when they
// fit (<= MAX_SMEM_CONSTS); larger sets (the dense HMPC matrices) are read through the read-only path.
#pragma once
#include "spcies_host.cuh"

namespace spcies {

constexpr size_t SMEM_MAX_BYTES = 227 * 1024;
constexpr size_t MAX_SMEM_CONSTS = 64 * 1024;
constexpr int MIN_SMEM_THREADS = 64;
#ifndef SPCIES_GLOBAL_STATE_BLOCK
#define SPCIES_GLOBAL_STATE_BLOCK 128
#endif
constexpr int GLOBAL_STATE_BLOCK = SPCIES_GLOBAL_STATE_BLOCK;   // threads per CTA when the iterates live in the global scratch

template <typename T, int STRIDE_CT> struct StateRef {
    // STRIDE_CT > 0: shared memory with a compile-time stride (= block size); 0: global memory, run-time stride
    T *base;
    int stride_rt;
    __device__ __forceinline__ int stride() const { return STRIDE_CT > 0 ? STRIDE_CT : stride_rt; }
    __device__ __forceinline__ T ld(int e) const { return base[(size_t)e * stride()]; }
    __device__ __forceinline__ void st(int e, T v) const { base[(size_t)e * stride()] = v; }
};

template <class S> struct KernelPlan {
    typedef typename S::real real;
    static constexpr size_t CONSTS_BYTES = (sizeof(spcies_consts) + 15) / 16 * 16;
    static constexpr bool GCONST = CONSTS_BYTES > MAX_SMEM_CONSTS;
    static constexpr size_t SMEM_FOR_STATE = SMEM_MAX_BYTES - 64 - (GCONST ? 0 : CONSTS_BYTES);
    static constexpr int fit(int elems) { return (int)(SMEM_FOR_STATE / ((size_t)elems * sizeof(real))) / 32 * 32; }
    static constexpr int cap(int t) { return t > 256 ? 256 : t; }
    static constexpr bool GSTATE_FIXED = fit(S::STATE) < MIN_SMEM_THREADS;
    static constexpr bool GSTATE_VARB = fit(S::STATE_VARB) < MIN_SMEM_THREADS;
    static constexpr int BLOCK_FIXED = GSTATE_FIXED ? GLOBAL_STATE_BLOCK : cap(fit(S::STATE));
    static constexpr int BLOCK_VARB = GSTATE_VARB ? GLOBAL_STATE_BLOCK : cap(fit(S::STATE_VARB));
};

template <class S, bool EXACT, bool VARB, int BLOCK, bool GSTATE, bool GCONST>
__global__ void __launch_bounds__(BLOCK, 1)
    persistent_kernel(const BatchIO io, const spcies_consts *__restrict__ g_consts, typename S::real *g_state) {
    typedef typename S::real real;
    typedef Arith<real, EXACT> A;
    typedef StateRef<real, GSTATE ? 0 : BLOCK> ST;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t mbar;
    const spcies_consts *C = g_consts;
    size_t off = 0;
    if (!GCONST) {
        stage_constants(smem_raw, g_consts, (uint32_t)KernelPlan<S>::CONSTS_BYTES, &mbar);
        C = reinterpret_cast<const spcies_consts *>(smem_raw);
        off = KernelPlan<S>::CONSTS_BYTES;
    }
    ST st;
    if (GSTATE) {
        st.base = g_state + (size_t)blockIdx.x * BLOCK + threadIdx.x;
        st.stride_rt = (int)(gridDim.x * BLOCK);
    } else {
        st.base = reinterpret_cast<real *>(smem_raw + off) + threadIdx.x;
        st.stride_rt = BLOCK;
    }
    typename S::template Ctx<A, VARB, ST> ctx(C, st, io);

    WorkQueue wq{io.queue, io.B, io.ready};
    wq.mark_start();
    unsigned long long stat_k = 0;
    unsigned int stat_nc = 0;
    long long inst = -1;
    int k = 0;
    for (;;) {
        if (inst < 0) {
            inst = wq.next();
            if (inst < 0) {
                wq.mark_drained();
                break;
            }
            ctx.init(inst);
            k = 0;
        }
        k += 1;
        const bool conv = ctx.iterate(k);
        const int ef = conv ? 1 : ((k >= k_max) ? -1 : 0);   // `e_flag = 1` wins over k_max, code_equMPC_ADMM_C.c:535-542
        if (ef != 0) {
            ctx.finish(inst, k, ef);
            stat_k += (unsigned long long)k;
            stat_nc += (ef < 0);
            inst = -1;
        }
    }
    flush_stats(io.queue, stat_k, stat_nc);
    wq.mark_end();
}

// Host-side traits for a policy solver
template <class S> struct PolicyTraits {
    typedef KernelPlan<S> P;
    typedef typename S::real real;
    static constexpr int NN = nn_, MM = mm_, NMM = nm_;
    static constexpr bool HAS_R = (SPCIES_HAS_R != 0);
#ifdef SPCIES_NREF
    static constexpr int NREF = SPCIES_NREF;
#else
    static constexpr int NREF = 1;
#endif
#if defined(TIME_VARYING) && TIME_VARYING == 1
    static constexpr bool TV = true;       // per-instance model: extra inputs A, B, Q, R (and LB, UB through opts)
#else
    static constexpr bool TV = false;
#endif
    // widths of the extra per-instance inputs BatchIO::ex[i] (0: unused)
    static constexpr int extra_width(int i) {
        if (TV) return i == 0 ? nn_ * nn_ : (i == 1 ? nn_ * mm_ : (i == 2 ? nn_ : mm_));
        if (NREF == 3) return i < 2 ? nn_ : mm_;
        return 0;
    }
    static constexpr bool HAS_VARB = S::HAS_VARB;
    static constexpr int SOL_DOUBLES = (int)(sizeof(SPCIES_SOL_T) / sizeof(double));
    typedef spcies_consts Consts;
    static const Consts &host_consts() { return spcies_h_consts; }
    static constexpr bool HAS_PARK = false;
    static constexpr int PARK_DOUBLES = 0;
    static size_t blob_bytes() { return sizeof(spcies_consts); }
    static void fill_blob(void *dst) { memcpy(dst, &spcies_h_consts, sizeof spcies_h_consts); }
    static int resume_block(bool varb) { return default_block(varb); }
    static void engine_shape(int, const BatchIO &, int &block, size_t &, int &ipb) { ipb = block; }
    static bool caps_engine(int, const BatchIO &) { return false; }
    static bool cl_engine(int, const BatchIO &) { return false; }     // no in-kernel closed loop: one launch per sampling time
    static bool single_engine(int, const BatchIO &) { return false; } // no one-CTA-per-instance latency engine
    static constexpr bool HAS_SERVER = false;                         // ... and no lingering server for the single-instance symbol
    static cudaError_t launch_server(cudaStream_t, const void *, void *, unsigned int, unsigned long long, int &, size_t &) {
        return cudaErrorNotSupported;
    }
    static cudaError_t launch_single(bool, int, cudaStream_t, const BatchIO &, const void *, int &, size_t &, const double *, const double *,
                                     const double *) {
        return cudaErrorNotSupported;
    }
    static bool uses_scratch(int, const BatchIO &) { return true; }   // the global per-instance state of the scalar kernels
    static size_t engine_scratch_bytes(int, const BatchIO &, int) { return 0; }   // global scratch of a tensor-core engine, if any
    static constexpr int K_MAX = 0;
    static cudaError_t init_device_symbols() { return cudaSuccess; }
    static int default_block(bool varb) { return varb ? P::BLOCK_VARB : P::BLOCK_FIXED; }
    static bool gstate(bool varb) { return varb ? P::GSTATE_VARB : P::GSTATE_FIXED; }
    static size_t smem_bytes(int block, bool varb) {
        size_t b = P::GCONST ? 0 : P::CONSTS_BYTES;
        if (!gstate(varb)) b += (size_t)(varb ? S::STATE_VARB : S::STATE) * block * sizeof(real);
        return b;
    }
    static size_t scratch_bytes(int grid, int block, bool varb) {
        return gstate(varb) ? (size_t)(varb ? S::STATE_VARB : S::STATE) * grid * block * sizeof(real) : 0;
    }
    template <bool EXACT, bool VARB>
    static cudaError_t launch_t(int grid, size_t smem, cudaStream_t s, const BatchIO &io, const void *dc, void *scratch) {
        constexpr int BLOCK = VARB ? P::BLOCK_VARB : P::BLOCK_FIXED;
        constexpr bool GS = VARB ? P::GSTATE_VARB : P::GSTATE_FIXED;
        auto kern = persistent_kernel<S, EXACT, VARB, BLOCK, GS, P::GCONST>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, BLOCK, smem, s>>>(io, (const spcies_consts *)dc, (real *)scratch);
        return cudaGetLastError();
    }
    static cudaError_t launch(int arith, bool varb, int grid, int block, size_t smem, cudaStream_t s, const BatchIO &io,
                              const void *dc, void *scratch) {
        if (io.engine == SPCIES_CUDA_ENGINE_MMA) return cudaErrorNotSupported;   // this skeleton is the scalar engine
        if (block != default_block(varb)) return cudaErrorInvalidConfiguration;
        const bool ex = arith == SPCIES_CUDA_ARITH_EXACT;
        if (varb) {
            if (!HAS_VARB) return cudaErrorNotSupported;
            return ex ? launch_t<true, S::HAS_VARB>(grid, smem, s, io, dc, scratch)
                      : launch_t<false, S::HAS_VARB>(grid, smem, s, io, dc, scratch);
        }
        return ex ? launch_t<true, false>(grid, smem, s, io, dc, scratch) : launch_t<false, false>(grid, smem, s, io, dc, scratch);
    }
    static cudaError_t attributes(int arith, bool varb, cudaFuncAttributes *a) {
        const bool ex = arith == SPCIES_CUDA_ARITH_EXACT;
        constexpr bool V = S::HAS_VARB;
        if (varb && V)
            return ex ? cudaFuncGetAttributes(a, persistent_kernel<S, true, V, P::BLOCK_VARB, P::GSTATE_VARB, P::GCONST>)
                      : cudaFuncGetAttributes(a, persistent_kernel<S, false, V, P::BLOCK_VARB, P::GSTATE_VARB, P::GCONST>);
        return ex ? cudaFuncGetAttributes(a, persistent_kernel<S, true, false, P::BLOCK_FIXED, P::GSTATE_FIXED, P::GCONST>)
                  : cudaFuncGetAttributes(a, persistent_kernel<S, false, false, P::BLOCK_FIXED, P::GSTATE_FIXED, P::GCONST>);
    }
};

}  // namespace spcies
