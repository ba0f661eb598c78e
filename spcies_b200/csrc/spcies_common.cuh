// spcies_common.cuh -- device-side building blocks shared by every solver kernel.
//
//  * Arith<T, EXACT>   arithmetic policy.  EXACT = same IEEE operations in the same order as the
//                      reference C compiled by gcc -O3 for x86-64 (no FMA contraction, true division,
//                      correctly rounded sqrt) -> bit-identical iterates.  FAST = fused multiply-add.
//  * clip()            the reference's saturation, `(v > LB) ? v : LB` then `(v > UB) ? UB : v`
//                      (code_laxMPC_FISTA_C.c:488-489) -- written with the same ternaries so NaN and
//                      LB > UB behave identically.
//  * State<T, BLOCK>   per-instance iterates in shared memory, laid out [element][thread] so that a
//                      warp touching element e of its 32 instances hits 32 consecutive words
//                      (conflict-free; a 64-bit access is the minimum two wavefronts).
//  * stage_constants() one bulk asynchronous copy (cp.async.bulk, the 1-D TMA path; SASS UBLKCP) of the
//                      problem-constant blob global -> shared per CTA, completion on an mbarrier.
//  * WorkQueue         persistent-kernel instance queue: one 64-bit atomic counter per launch; a lane
//                      that finishes its instance pulls the next index (ptxas aggregates the atomics of
//                      a warp), so the heavy-tailed iteration counts do not idle lanes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spcies {

// ------------------------------------------------------------------------------------------------
// Arithmetic policies
// ------------------------------------------------------------------------------------------------
template <typename T, bool EXACT> struct Arith;

template <> struct Arith<double, false> {
    typedef double T;
    static __device__ __forceinline__ T add(T a, T b) { return a + b; }
    static __device__ __forceinline__ T sub(T a, T b) { return a - b; }
    static __device__ __forceinline__ T mul(T a, T b) { return a * b; }
    static __device__ __forceinline__ T div(T a, T b) { return a / b; }
    static __device__ __forceinline__ T madd(T a, T b, T c) { return fma(b, c, a); }    // a + b*c
    static __device__ __forceinline__ T nmsub(T a, T b, T c) { return fma(-b, c, a); }  // a - b*c
    static __device__ __forceinline__ T sqrt(T a) { return ::sqrt(a); }
};
template <> struct Arith<double, true> {
    typedef double T;
    static __device__ __forceinline__ T add(T a, T b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ T sub(T a, T b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ T mul(T a, T b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ T div(T a, T b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ T madd(T a, T b, T c) { return __dadd_rn(a, __dmul_rn(b, c)); }
    static __device__ __forceinline__ T nmsub(T a, T b, T c) { return __dsub_rn(a, __dmul_rn(b, c)); }
    static __device__ __forceinline__ T sqrt(T a) { return __dsqrt_rn(a); }
};
template <> struct Arith<float, false> {
    typedef float T;
    static __device__ __forceinline__ T add(T a, T b) { return a + b; }
    static __device__ __forceinline__ T sub(T a, T b) { return a - b; }
    static __device__ __forceinline__ T mul(T a, T b) { return a * b; }
    static __device__ __forceinline__ T div(T a, T b) { return a / b; }
    static __device__ __forceinline__ T madd(T a, T b, T c) { return fmaf(b, c, a); }
    static __device__ __forceinline__ T nmsub(T a, T b, T c) { return fmaf(-b, c, a); }
    static __device__ __forceinline__ T sqrt(T a) { return ::sqrtf(a); }
};
template <> struct Arith<float, true> {
    typedef float T;
    static __device__ __forceinline__ T add(T a, T b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ T sub(T a, T b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ T mul(T a, T b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ T div(T a, T b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ T madd(T a, T b, T c) { return __fadd_rn(a, __fmul_rn(b, c)); }
    static __device__ __forceinline__ T nmsub(T a, T b, T c) { return __fsub_rn(a, __fmul_rn(b, c)); }
    static __device__ __forceinline__ T sqrt(T a) { return __fsqrt_rn(a); }
};

// Product of two generated constants as the C templates evaluate it: in the constants' own type, then promoted.  With
// precision = 'float' the constants are `const static float` and `AB[j][i]*Hi[l-1][i]*z[l-1][i]` (code_equMPC_ADMM_C.c:340) is
// a float product times a double; with double constants this is the plain product of the arithmetic policy.
template <class A> __device__ __forceinline__ typename A::T cprod(double a, double b) { return A::mul((typename A::T)a, (typename A::T)b); }
template <class A> __device__ __forceinline__ typename A::T cprod(float a, float b) { return (typename A::T)__fmul_rn(a, b); }
static inline double cprod_host(double a, double b) { return a * b; }
static inline double cprod_host(float a, float b) {
    volatile float p = a * b;
    return (double)p;
}

template <typename T> __device__ __forceinline__ T clip(T v, T lb, T ub) {
    v = (v > lb) ? v : lb;   // maximum between v and the lower bound
    v = (v > ub) ? ub : v;   // minimum between v and the upper bound
    return v;
}

// |x| > tol exactly as `res = (res > 0.0) ? res : -res; if (res > tol)` (code_laxMPC_FISTA_C.c:343-344)
template <typename T> __device__ __forceinline__ bool exceeds(T x, T tol_) {
    T a = (x > T(0)) ? x : -x;
    return a > tol_;
}

// ------------------------------------------------------------------------------------------------
// Engineering units (options.in_engineering): inputs are scaled as `scaling_x[i]*( x0_in[i] - OpPoint_x[i] )`, the control
// action is returned as `u[j]*scaling_i_u[j] + OpPoint_u[j]` (code_laxMPC_FISTA_C.c:75-82, :398-402; the same block in every
// template).  Individually rounded operations in both arithmetic modes, as gcc -O3 emits them for x86-64.  `C` is the generated
// constant struct: it has the five scaling arrays only when the solver was generated with in_engineering (the `#define` decides).
// ------------------------------------------------------------------------------------------------
#if defined(in_engineering) && in_engineering == 1
template <class CT> __device__ __forceinline__ double eng_x(const CT *C, const double *a, long long inst, int n_, int i) {
    return __dmul_rn((double)C->scaling_x[i], __dsub_rn(a[inst * n_ + i], (double)C->OpPoint_x[i]));
}
template <class CT> __device__ __forceinline__ double eng_u(const CT *C, const double *a, long long inst, int m_, int i) {
    return __dmul_rn((double)C->scaling_u[i], __dsub_rn(a[inst * m_ + i], (double)C->OpPoint_u[i]));
}
template <class CT> __device__ __forceinline__ double eng_u_out(const CT *C, double u, int j) {
    return __dadd_rn(__dmul_rn(u, (double)C->scaling_i_u[j]), (double)C->OpPoint_u[j]);
}
template <class CT> __device__ __forceinline__ double eng_x_value(const CT *C, double v, int i) {
    return __dmul_rn((double)C->scaling_x[i], __dsub_rn(v, (double)C->OpPoint_x[i]));
}
template <class CT> __device__ __forceinline__ double eng_u_value(const CT *C, double v, int i) {
    return __dmul_rn((double)C->scaling_u[i], __dsub_rn(v, (double)C->OpPoint_u[i]));
}
#else
template <class CT> __device__ __forceinline__ double eng_x(const CT *, const double *a, long long inst, int n_, int i) {
    return a[inst * n_ + i];
}
template <class CT> __device__ __forceinline__ double eng_u(const CT *, const double *a, long long inst, int m_, int i) {
    return a[inst * m_ + i];
}
template <class CT> __device__ __forceinline__ double eng_u_out(const CT *, double u, int) { return u; }
template <class CT> __device__ __forceinline__ double eng_x_value(const CT *, double v, int) { return v; }
template <class CT> __device__ __forceinline__ double eng_u_value(const CT *, double v, int) { return v; }
#endif

// ------------------------------------------------------------------------------------------------
// Per-instance state in shared memory: element-major, thread-minor
// ------------------------------------------------------------------------------------------------
template <typename T, int BLOCK> struct State {
    T *base;   // already offset by threadIdx.x
    __device__ __forceinline__ explicit State(T *smem_state) : base(smem_state + threadIdx.x) {}
    __device__ __forceinline__ T ld(int e) const { return base[e * BLOCK]; }
    __device__ __forceinline__ void st(int e, T v) const { base[e * BLOCK] = v; }
};

// ------------------------------------------------------------------------------------------------
// Constant blob: global -> shared, one bulk async copy per CTA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// bytes must be a multiple of 16; dst/src 16-byte aligned.  All threads of the CTA must call it.
__device__ __forceinline__ void stage_constants(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                                uint64_t *mbar /* in shared memory */) {
    const uint32_t bar = smem_u32(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(smem_dst)),
            "l"(gmem_src), "r"(bytes), "r"(bar)
            : "memory");
    }
    // every thread waits for phase 0 of the barrier (try_wait suspends in hardware, it is not a spin)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar)
        : "memory");
}

// ------------------------------------------------------------------------------------------------
// Device-side view of one batched call
// ------------------------------------------------------------------------------------------------
struct BatchIO {
    long long B;
    const double *x0, *xr, *ur, *r;   // [B][nn_], [B][nn_], [B][mm_], [B] (r: solvers with r_ellip only)
    // up to four extra per-instance inputs (widths Traits::extra_width(i)), else nullptr:
    //   ellipHMPC (three references): ex[0] = x_rs, ex[1] = x_rc [B][nn_], ex[2] = u_rs, ex[3] = u_rc [B][mm_]
    //   TIME_VARYING solvers: ex[0] = A [B][nn_*nn_], ex[1] = B [B][nn_*mm_] (column-major per instance), ex[2] = Q [B][nn_], ex[3] = R [B][mm_]
    const double *ex[4];
    const double *LB, *UB;            // optional per-instance bounds [B][nm_], or nullptr
    double *u;                        // [B][mm_]
    int *k, *e;                       // [B]
    double *sol;                      // optional [B][sol_doubles] (sol_<name> structs), or nullptr
    unsigned long long *queue;        // [0] next instance index, [1] sum of k, [2] number of e_flag = -1,
                                      // [3] ~(first time a lane found the queue empty), [4] ~(kernel start), [5] kernel end
                                      //     (globaltimer ns; min kept as max of the complement so that memset(0) initialises)
                                      // [6] number of parked instances, [7] next parked record (phase 2)
                                      // [9] number of records to resume (copy of [6] made between two launches), [10] park events
    // tail handling of kernels that support it (Traits::HAS_PARK): see MPC_FISTA.cuh
    double *park;                     // [PARK_DOUBLES][park_cap] records of parked instances, element-major
    long long park_cap;
    int phase;                        // 0: one launch runs every instance to its end
                                      // 1: park the instances still running `grace` iterations after the queue ran dry
                                      // 2: resume the parked instances (B is read from queue[6])
    int grace;
    const double *park_in;            // phase 2 of kernels with iteration caps: the records to resume (count in queue[9]) ...
    long long park_in_cap;            // ... while `park` receives the instances that reach this launch's cap
    int cap;                          // > 0: park an instance once it has done `cap` iterations (MPC_FISTA_mma.cuh)
    int engine;                       // spcies_batch_opts.engine (kernels with more than one engine, MPC_FISTA.cuh)
    const unsigned long long *ready;  // optional: instances [0, *ready) have their inputs in device memory (the host->device
                                      // copies of a host-buffer call run on a second stream, chunk by chunk, under the kernel)    // closed-loop runs of engines that keep an instance on chip across sampling times (Traits::cl_engine, MPC_FISTA_mma.cuh):
    // u / k / e are trajectories [cl_steps][cl_ld][..], x0 is sampling time 0, cl_x (optional) receives x at times 1..cl_steps
    int cl_steps;                     // 0: an ordinary batched call
    int cl_warm;                      // 1: the dual point of the previous sampling time is the starting point of the next
    long long cl_ld;                  // instances per sampling time in the trajectory arrays
    double *cl_x;                     // [cl_steps + 1][cl_ld][nn_] or nullptr
    // latency engines (Traits::single_engine): per-instance completion flags in mapped host memory, or nullptr
    unsigned int *done;               // [B]: set to done_seq (after a system-wide fence) once u / k / e of the instance are written
    unsigned int done_seq;
};
// Mailbox of a lingering single-instance server kernel (Traits::HAS_SERVER, MPC_FISTA_single.cuh), in mapped pinned host memory.
// Host -> device: 64-byte lines of seven doubles + (seq, cmd); the payload is (x0[NN], xr[NN], ur[MM]) and a request is complete
// when every line carries the new sequence number (the host writes a line's doubles, then -- after a store fence -- its seq; a
// PCIe read returns a consistent snapshot of a 64-byte line).  cmd of line 0 != 0 tells the server to exit.
// Device -> host: one line with the results and the sequence number they answer (written after a system-wide fence); `alive` is
// set by the host before a launch and cleared by the kernel as its last action.
template <int NN, int MM> struct alignas(64) SingleMailbox {
    static constexpr int NL = (2 * NN + MM + 6) / 7;
    struct alignas(64) Line {
        double v[7];
        unsigned int seq, cmd;
    } in[NL];
    struct alignas(64) Out {
        double u[MM < 5 ? 5 : MM];
        int k, e;
        unsigned int seq, pad;
    } out;
    alignas(64) unsigned int alive;
};

constexpr int QUEUE_WORDS = 16;   // [8]: number of instances whose inputs have arrived (pipelined host->device copies)

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct WorkQueue {
    unsigned long long *q;
    long long B;
    const unsigned long long *ready = nullptr;
    __device__ __forceinline__ long long next() const {
        unsigned long long i = atomicAdd(q, 1ULL);
        if (i >= (unsigned long long)B) return -1LL;
        if (ready) {   // inputs still in flight: wait for the watermark (chunk boundaries never share a cache line)
            while (*(const volatile unsigned long long *)ready <= i) __nanosleep(500);
        }
        return (long long)i;
    }
    // launch-shape telemetry (three atomics per warp and launch): when the kernel started, when the queue ran dry (the
    // start of the tail in which only the slowest instances are still iterating) and when the last warp left
    __device__ __forceinline__ void mark_start() const {
        if ((threadIdx.x & 31) == 0) atomicMax(q + 4, ~globaltimer_ns());
    }
    __device__ __forceinline__ void mark_drained() const { atomicMax(q + 3, ~globaltimer_ns()); }
    __device__ __forceinline__ void mark_end() const {
        if ((threadIdx.x & 31) == 0) atomicMax(q + 5, globaltimer_ns());
    }
};

// lane-local statistics -> two atomics per warp at kernel exit
__device__ __forceinline__ void flush_stats(unsigned long long *q, unsigned long long sum_k, unsigned int n_nc) {
    for (int o = 16; o > 0; o >>= 1) {
        sum_k += __shfl_down_sync(0xffffffffu, sum_k, o);
        n_nc += __shfl_down_sync(0xffffffffu, n_nc, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (sum_k) atomicAdd(q + 1, sum_k);
        if (n_nc) atomicAdd(q + 2, (unsigned long long)n_nc);
    }
}

}  // namespace spcies
