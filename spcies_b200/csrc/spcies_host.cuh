// spcies_host.cuh -- host side of a generated solver library: device contexts, buffers, the batched
// call (H2D -> persistent kernel -> D2H), contiguous sharding over several GPUs (one host thread per
// device, no collective: instances are independent), and the reference's
// single-instance symbol implemented as a batch of one.
//
// A solver kernel header provides a `Traits` type:
//   static constexpr int NN, MM, NMM;            // nn_, mm_, nm_
//   static constexpr bool HAS_R;                 // extra r_ellip input
//   static constexpr int SOL_DOUBLES;            // sizeof(sol_<name>)/8
//   static constexpr bool HAS_VARB;              // per-instance bounds supported by the kernel
//   typedef ... Consts;  static const Consts &host_consts();
//   static int default_block();  static size_t smem_bytes(int block, bool varb);
//   static size_t scratch_bytes(int grid, int block, bool varb);   // global per-instance state, 0 if none
//   static cudaError_t launch(int arith, bool varb, int grid, int block, size_t smem, cudaStream_t s,
//                             const BatchIO &io, const void *d_consts, void *d_scratch);
//   static cudaError_t attributes(int arith, bool varb, cudaFuncAttributes *a);
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "spcies_common.cuh"
#include "spcies_cuda.h"

namespace spcies {

static thread_local char g_last_error[512] = "";
static char g_last_error_global[512] = "";

static inline int fail(int code, const char *what) {
    snprintf(g_last_error, sizeof g_last_error, "%s (code %d%s%s)", what, code,
             (code > 0 && code < 1000) ? ": " : "", (code > 0 && code < 1000) ? cudaGetErrorString((cudaError_t)code) : "");
    memcpy(g_last_error_global, g_last_error, sizeof g_last_error);
    return code;
}
#define SPCIES_CK(call)                                           \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) return ::spcies::fail((int)_e, #call); \
    } while (0)

constexpr int MAX_CHUNKS = 32;

struct DeviceCtx {
    int dev = -1;
    bool ready = false;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;          // host->device input chunks, overlapped with the solver kernel
    unsigned long long *h_ready = nullptr;       // pinned: watermark values copied to d_queue[8] after every chunk
    double *h_stage = nullptr;                   // pinned, device-mapped staging of the small-batch path (zero-copy in / out)
    double *d_stage = nullptr;                   // its device address
    unsigned int small_seq = 0;                  // call counter of the latency engine (value of its completion flags)
    void *h_mb = nullptr, *d_mb = nullptr;       // mailbox of the lingering single-instance server (mapped pinned host memory)
    unsigned int srv_seq = 0;                    // sequence number of the last request posted to it
    cudaStream_t srv_stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void *d_consts = nullptr;
    unsigned long long *d_queue = nullptr;
    // device staging buffers (grown on demand)
    long long cap = 0, cap_sol = 0, cap_b = 0;
    double *d_x0 = nullptr, *d_xr = nullptr, *d_ur = nullptr, *d_r = nullptr, *d_LB = nullptr, *d_UB = nullptr;
    double *d_ex[4] = {nullptr, nullptr, nullptr, nullptr};   // extra per-instance inputs (Traits::extra_width: ellipHMPC, TIME_VARYING)
    double *d_u = nullptr, *d_sol = nullptr;
    int *d_k = nullptr, *d_e = nullptr;
    void *d_scratch = nullptr;   // per-instance state of solvers whose iterates do not fit shared memory
    size_t cap_scratch = 0;
    double *d_park = nullptr;    // parked-instance records (tail handling, Traits::HAS_PARK)
    double *d_park2 = nullptr;   // second record buffer of the iteration-cap rounds (a round resumes from one, parks into the other)
    long long cap_park = 0;
    // closed-loop trajectories (run_closed_loop): x [steps + 1][B][nn_], u [steps][B][mm_], k / e [steps][B]; the plant [A B]
    double *d_clx = nullptr, *d_clu = nullptr, *d_plant = nullptr;
    int *d_clk = nullptr, *d_cle = nullptr;
    long long cap_cl = 0;        // capacity in (steps + 1) * B units
};

// x+ = [A B] (x; u) for every instance, accumulated in the order of examples/cl_in_C/main_cl_in_C.c:104-113 with individually
// rounded operations (what gcc -O3 emits for x86-64): the trajectories of a closed-loop run in EXACT arithmetic are bit-identical
// to a loop of reference calls.
template <int NN, int MM>
__global__ void cl_plant_kernel(long long B, const double *__restrict__ AB, const double *__restrict__ x, const double *__restrict__ u,
                                double *__restrict__ xn) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double xv[NN], uv[MM];
#pragma unroll
    for (int j = 0; j < NN; ++j) xv[j] = x[i * NN + j];
#pragma unroll
    for (int j = 0; j < MM; ++j) uv[j] = u[i * MM + j];
#pragma unroll
    for (int r = 0; r < NN; ++r) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < NN; ++j) a = __dadd_rn(a, __dmul_rn(AB[r * (NN + MM) + j], xv[j]));
#pragma unroll
        for (int j = 0; j < MM; ++j) a = __dadd_rn(a, __dmul_rn(AB[r * (NN + MM) + NN + j], uv[j]));
        xn[i * NN + r] = a;
    }
}

template <class Traits> struct Runtime {
    std::vector<DeviceCtx> ctx;
    std::mutex mu;

    static Runtime &get() {
        static Runtime r;
        return r;
    }

    int device_count() {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        return n;
    }

    int init_device(int dev, DeviceCtx **out) {
        int n = device_count();
        if (n <= 0) return fail(SPCIES_CUDA_ENODEVICE, "no usable CUDA device (this library has no CPU fallback)");
        if (dev < 0 || dev >= n) return fail(SPCIES_CUDA_ENODEVICE, "requested CUDA device does not exist");
        {
            std::lock_guard<std::mutex> lk(mu);
            if ((int)ctx.size() < n) ctx.resize(n);
        }
        DeviceCtx &c = ctx[dev];
        SPCIES_CK(cudaSetDevice(dev));
        if (!c.ready) {
            c.dev = dev;
            SPCIES_CK(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, dev));
            SPCIES_CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
            SPCIES_CK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
            SPCIES_CK(cudaHostAlloc((void **)&c.h_ready, MAX_CHUNKS * sizeof(unsigned long long), cudaHostAllocDefault));
            for (auto &e : c.ev) SPCIES_CK(cudaEventCreate(&e));
            std::vector<unsigned char> blob(Traits::blob_bytes());
            Traits::fill_blob(blob.data());
            SPCIES_CK(cudaMalloc(&c.d_consts, blob.size()));
            SPCIES_CK(cudaMemcpy(c.d_consts, blob.data(), blob.size(), cudaMemcpyHostToDevice));
            SPCIES_CK(Traits::init_device_symbols());
            SPCIES_CK(cudaMalloc(&c.d_queue, QUEUE_WORDS * sizeof(unsigned long long)));
            c.ready = true;
        }
        *out = &c;
        return 0;
    }

    template <typename P> static int grow(P **p, long long count) {
        if (*p) SPCIES_CK(cudaFree(*p));
        *p = nullptr;
        SPCIES_CK(cudaMalloc((void **)p, (size_t)count * sizeof(P)));
        return 0;
    }
    // a failed allocation leaves *p == nullptr: the capacities are cleared first so that no later, smaller call trusts them
    static void drop_caps(DeviceCtx &c) { c.cap = c.cap_b = c.cap_sol = c.cap_cl = 0; }

    int reserve(DeviceCtx &c, long long B, bool varb, bool sol) {
        if (B > c.cap) {
            long long n = B + B / 8 + 64;
            int rc;
            c.cap = 0;
            if ((rc = grow(&c.d_x0, n * Traits::NN))) return rc;
            if ((rc = grow(&c.d_xr, n * Traits::NN))) return rc;
            if ((rc = grow(&c.d_ur, n * Traits::MM))) return rc;
            if (Traits::HAS_R && (rc = grow(&c.d_r, n))) return rc;
            for (int i = 0; i < 4; ++i)
                if (Traits::extra_width(i) > 0 && (rc = grow(&c.d_ex[i], n * Traits::extra_width(i)))) return rc;
            if ((rc = grow(&c.d_u, n * Traits::MM))) return rc;
            if ((rc = grow(&c.d_k, n))) return rc;
            if ((rc = grow(&c.d_e, n))) return rc;
            c.cap = n;
        }
        if (varb && B > c.cap_b) {
            int rc;
            c.cap_b = 0;
            if ((rc = grow(&c.d_LB, B * Traits::NMM))) return rc;
            if ((rc = grow(&c.d_UB, B * Traits::NMM))) return rc;
            c.cap_b = B;
        }
        if (sol && B > c.cap_sol) {
            int rc;
            c.cap_sol = 0;
            if ((rc = grow(&c.d_sol, B * (long long)Traits::SOL_DOUBLES))) return rc;
            c.cap_sol = B;
        }
        return 0;
    }

    void free_all() {
        for (auto &c : ctx) {
            if (!c.ready) continue;
            cudaSetDevice(c.dev);
            cudaFree(c.d_consts); cudaFree(c.d_queue);
            cudaFree(c.d_x0); cudaFree(c.d_xr); cudaFree(c.d_ur); cudaFree(c.d_r); cudaFree(c.d_LB); cudaFree(c.d_UB);
            for (auto p_ : c.d_ex) cudaFree(p_);
            cudaFree(c.d_clx); cudaFree(c.d_clu); cudaFree(c.d_clk); cudaFree(c.d_cle); cudaFree(c.d_plant);
            cudaFree(c.d_u); cudaFree(c.d_sol); cudaFree(c.d_k); cudaFree(c.d_e); cudaFree(c.d_scratch); cudaFree(c.d_park); cudaFree(c.d_park2);
            for (auto &e : c.ev) cudaEventDestroy(e);
            cudaStreamDestroy(c.stream);
            cudaStreamDestroy(c.copy_stream);
            cudaFreeHost(c.h_ready);
            if (c.h_stage) cudaFreeHost(c.h_stage);
            if (c.h_mb) {
                stop_server(c, true);
                cudaStreamDestroy(c.srv_stream);
                cudaFreeHost(c.h_mb);
            }
            c = DeviceCtx();
        }
    }

    struct Call {
        long long B;
        const double *x0, *xr, *ur, *r, *LB, *UB;
        const double *ex[4] = {nullptr, nullptr, nullptr, nullptr};
        double *u;
        int *k, *e;
        double *sol;
        int arith, block, grid;
        int tail_mode, tail_grace, engine;
        int tail_caps[3];
        bool device_pointers;
        cudaStream_t user_stream;
    };
    struct Result {
        int rc = 0;
        double kernel_ms = 0, h2d_ms = 0, d2h_ms = 0;
        long long sum_k = 0, n_nc = 0;
        int block = 0, grid = 0, smem = 0;
        int drain_us = 0, span_us = 0;
        int launches = 0;
        long long parked = 0;
    };

    static bool direct_results() {
        static const bool on = [] {
            const char *v = getenv("SPCIES_CUDA_DIRECT_RESULTS");   // set to 0 to force the device->host copies
            return v == nullptr || v[0] != '0';
        }();
        return on;
    }

    // ---- lingering server of the single-instance symbol (Traits::HAS_SERVER) --------------------------------------------------
    typedef SingleMailbox<Traits::NN, Traits::MM> Mailbox;
    static long long server_linger_us() {
        static const long long us = [] {
            const char *v = getenv("SPCIES_CUDA_SERVER_LINGER_US");    // 0: no server, every single-instance call is a launch
            return v ? atoll(v) : 200LL;
        }();
        return us;
    }
    // tell a running server to exit (it also exits by itself once idle for the linger time); wait = until it has
    static void stop_server(DeviceCtx &c, bool wait) {
        if (!c.h_mb) return;
        Mailbox *mb = static_cast<Mailbox *>(c.h_mb);
        volatile unsigned int *alive = &mb->alive;
        if (*alive) {
            *reinterpret_cast<volatile unsigned int *>(&mb->in[0].cmd) = 1u;
            if (wait) cudaStreamSynchronize(c.srv_stream);
        }
    }
    int run_server(DeviceCtx &c, const Call &cl, Result &res) {
        if (!c.h_mb) {
            SPCIES_CK(cudaHostAlloc(&c.h_mb, sizeof(Mailbox), cudaHostAllocMapped));
            memset(c.h_mb, 0, sizeof(Mailbox));
            SPCIES_CK(cudaHostGetDevicePointer(&c.d_mb, c.h_mb, 0));
            SPCIES_CK(cudaStreamCreateWithFlags(&c.srv_stream, cudaStreamNonBlocking));
        }
        Mailbox *mb = static_cast<Mailbox *>(c.h_mb);
        volatile unsigned int *alive = &mb->alive, *cmd = &mb->in[0].cmd, *oseq = &mb->out.seq;
        if (*cmd != 0) {                                 // a stop request is pending: let that server go first
            if (*alive) SPCIES_CK(cudaStreamSynchronize(c.srv_stream));
            *cmd = 0;
        }
        const unsigned int seq = ++c.srv_seq;
        double pay[Mailbox::NL * 7] = {};
        memcpy(pay, cl.x0, Traits::NN * 8);
        memcpy(pay + Traits::NN, cl.xr, Traits::NN * 8);
        memcpy(pay + 2 * Traits::NN, cl.ur, Traits::MM * 8);
        const auto t0 = std::chrono::steady_clock::now();
        for (int l = 0; l < Mailbox::NL; ++l)
            for (int j = 0; j < 7; ++j) *reinterpret_cast<volatile double *>(&mb->in[l].v[j]) = pay[l * 7 + j];
        std::atomic_thread_fence(std::memory_order_release);
        for (int l = 0; l < Mailbox::NL; ++l) *reinterpret_cast<volatile unsigned int *>(&mb->in[l].seq) = seq;
        std::atomic_thread_fence(std::memory_order_seq_cst);
        int block = 0;
        size_t smem = 0;
        for (unsigned long long spin = 1;; ++spin) {
            if (*oseq == seq) break;
            if (!*alive) {                               // no server (never started, or it has just lingered out): start one
                *alive = 1u;
                std::atomic_thread_fence(std::memory_order_seq_cst);
                const cudaError_t le = Traits::launch_server(c.srv_stream, c.d_consts, c.d_mb, seq - 1,
                                                             (unsigned long long)server_linger_us() * 1000ULL, block, smem);
                if (le != cudaSuccess) {
                    *alive = 0u;
                    SPCIES_CK(le);
                }
                res.launches += 1;
            }
            if ((spin & 0xfff) == 0 && std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > 2000.0) {
                cudaError_t e = cudaStreamQuery(c.srv_stream);
                if (e != cudaSuccess && e != cudaErrorNotReady) SPCIES_CK(e);
                return fail((int)cudaErrorLaunchTimeout, "the single-instance server did not answer");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        res.kernel_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();     // request to answer
        for (int j = 0; j < Traits::MM; ++j) cl.u[j] = *reinterpret_cast<volatile double *>(&mb->out.u[j]);
        cl.k[0] = *reinterpret_cast<volatile int *>(&mb->out.k);
        cl.e[0] = *reinterpret_cast<volatile int *>(&mb->out.e);
        res.sum_k = cl.k[0];
        res.n_nc = cl.e[0] < 0;
        res.block = block; res.grid = 1; res.smem = (int)smem;
        return 0;
    }

    // Small host-buffer batches (the reference's single-instance call is a batch of one): latency is API calls, not bytes.
    // Inputs are memcpy'd into a pinned, device-mapped staging block that the kernel reads directly, results come back the same
    // way, the statistics are summed on the host: one memset, one launch, two event records and one synchronisation instead of
    // seven copies.  Same kernels, same results.
    static constexpr long long SMALL_B = 64;
    static constexpr size_t stage_doubles() {
        return (size_t)SMALL_B * (2 * Traits::NN + 2 * Traits::MM + 1 + 2 * Traits::NMM + 1 + Traits::extra_width(0) + Traits::extra_width(1) +
                                  Traits::extra_width(2) + Traits::extra_width(3) + 1);
    }
    int run_small(DeviceCtx &c, const Call &cl, Result &res) {
        const long long B = cl.B;
        const bool varb = cl.LB != nullptr && cl.UB != nullptr;
        if (!c.h_stage) {
            SPCIES_CK(cudaHostAlloc((void **)&c.h_stage, stage_doubles() * sizeof(double), cudaHostAllocMapped));
            SPCIES_CK(cudaHostGetDevicePointer((void **)&c.d_stage, c.h_stage, 0));
        }
        // layout (doubles): x0 | xr | ur | r | LB | UB | u | (k, e as ints)
        double *h = c.h_stage, *d = c.d_stage;
        const size_t o_x0 = 0, o_xr = o_x0 + SMALL_B * Traits::NN, o_ur = o_xr + SMALL_B * Traits::NN, o_r = o_ur + SMALL_B * Traits::MM,
                     o_lb = o_r + SMALL_B, o_ub = o_lb + SMALL_B * Traits::NMM, o_u = o_ub + SMALL_B * Traits::NMM,
                     o_ke = o_u + SMALL_B * Traits::MM;
        size_t o_ex[4];
        o_ex[0] = o_ke + SMALL_B;
        for (int i = 1; i < 4; ++i) o_ex[i] = o_ex[i - 1] + SMALL_B * Traits::extra_width(i - 1);
        const size_t o_done = o_ex[3] + SMALL_B * Traits::extra_width(3);        // completion flags of the latency engine (one word per instance)
        memcpy(h + o_x0, cl.x0, (size_t)B * Traits::NN * 8);
        memcpy(h + o_xr, cl.xr, (size_t)B * Traits::NN * 8);
        memcpy(h + o_ur, cl.ur, (size_t)B * Traits::MM * 8);
        if (Traits::HAS_R) memcpy(h + o_r, cl.r, (size_t)B * 8);
        for (int i = 0; i < 4; ++i)
            if (Traits::extra_width(i) > 0) memcpy(h + o_ex[i], cl.ex[i], (size_t)B * Traits::extra_width(i) * 8);
        if (varb) {
            memcpy(h + o_lb, cl.LB, (size_t)B * Traits::NMM * 8);
            memcpy(h + o_ub, cl.UB, (size_t)B * Traits::NMM * 8);
        }
        BatchIO io;
        memset(&io, 0, sizeof io);
        io.B = B;
        io.queue = c.d_queue;
        io.x0 = d + o_x0; io.xr = d + o_xr; io.ur = d + o_ur; io.r = d + o_r;
        for (int i = 0; i < 4; ++i) io.ex[i] = Traits::extra_width(i) > 0 ? d + o_ex[i] : nullptr;
        io.LB = varb ? d + o_lb : nullptr; io.UB = varb ? d + o_ub : nullptr;
        io.u = d + o_u;
        io.k = reinterpret_cast<int *>(d + o_ke);
        io.e = io.k + SMALL_B;
        io.engine = cl.engine;
        if (Traits::HAS_SERVER && B == 1 && !varb && cl.engine == SPCIES_CUDA_ENGINE_AUTO && server_linger_us() > 0 &&
            Traits::single_engine(cl.arith, io))
            return run_server(c, cl, res);
        if (Traits::single_engine(cl.arith, io) && (!varb || Traits::HAS_VARB)) {
            // latency engine: one CTA per instance; no queue, no events, and the host waits on per-instance completion flags the
            // kernel sets in this (mapped) block after a system-wide fence -- a launch and a few PCIe round trips per call
            volatile unsigned int *hd = reinterpret_cast<volatile unsigned int *>(h + o_done);
            unsigned int seq = ++c.small_seq;
            if (seq == 0) seq = ++c.small_seq;
            io.done = reinterpret_cast<unsigned int *>(d + o_done);
            io.done_seq = seq;
            int block = 0;
            size_t smem = 0;
            const auto t0 = std::chrono::steady_clock::now();
            SPCIES_CK(Traits::launch_single(varb, (int)B, c.stream, io, c.d_consts, block, smem, cl.x0, cl.xr, cl.ur));
            bool done = false;
            for (unsigned long long spin = 1; !done; ++spin) {
                done = true;
                for (long long i = 0; i < B; ++i)
                    if (hd[i] != seq) {
                        done = false;
                        break;
                    }
                if (!done && (spin & 0xfff) == 0 &&
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > 5.0)
                    break;                                  // long solve (or a failed launch): wait the ordinary way
            }
            if (!done) {
                SPCIES_CK(cudaStreamSynchronize(c.stream));
                SPCIES_CK(cudaGetLastError());
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            res.kernel_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();   // launch to flag
            memcpy(cl.u, h + o_u, (size_t)B * Traits::MM * 8);
            const int *hk = reinterpret_cast<const int *>(h + o_ke), *he = hk + SMALL_B;
            memcpy(cl.k, hk, (size_t)B * 4);
            memcpy(cl.e, he, (size_t)B * 4);
            for (long long i = 0; i < B; ++i) {
                res.sum_k += hk[i];
                res.n_nc += (he[i] < 0);
            }
            res.launches = 1;
            res.block = block; res.grid = (int)B; res.smem = (int)smem;
            return 0;
        }
        if (cl.engine == SPCIES_CUDA_ENGINE_SINGLE) return fail(SPCIES_CUDA_EUNSUPPORTED, "the latency engine (one CTA per instance) does not take this call");
        int block = cl.block > 0 ? cl.block : Traits::default_block(varb), ipb = 0;
        size_t smem = Traits::smem_bytes(block, varb);
        Traits::engine_shape(cl.arith, io, block, smem, ipb);
        int grid = (int)((B + ipb - 1) / ipb);
        if (cl.grid > 0 && cl.grid < grid) grid = cl.grid;
        if (grid > c.sm_count) grid = c.sm_count;
        const size_t need_scratch = std::max<size_t>(Traits::uses_scratch(cl.arith, io) ? Traits::scratch_bytes(grid, block, varb) : 0,
                                                     Traits::engine_scratch_bytes(cl.arith, io, grid));
        if (need_scratch > c.cap_scratch) {
            c.cap_scratch = 0;
            if (c.d_scratch) SPCIES_CK(cudaFree(c.d_scratch));
            c.d_scratch = nullptr;
            SPCIES_CK(cudaMalloc(&c.d_scratch, need_scratch));
            c.cap_scratch = need_scratch;
        }
        cudaStream_t s = c.stream;
        SPCIES_CK(cudaMemsetAsync(c.d_queue, 0, QUEUE_WORDS * sizeof(unsigned long long), s));
        SPCIES_CK(cudaEventRecord(c.ev[1], s));
        SPCIES_CK(Traits::launch(cl.arith, varb, grid, block, smem, s, io, c.d_consts, c.d_scratch));
        SPCIES_CK(cudaEventRecord(c.ev[2], s));
        SPCIES_CK(cudaStreamSynchronize(s));
        SPCIES_CK(cudaGetLastError());
        memcpy(cl.u, h + o_u, (size_t)B * Traits::MM * 8);
        const int *hk = reinterpret_cast<const int *>(h + o_ke), *he = hk + SMALL_B;
        memcpy(cl.k, hk, (size_t)B * 4);
        memcpy(cl.e, he, (size_t)B * 4);
        float ms = 0;
        SPCIES_CK(cudaEventElapsedTime(&ms, c.ev[1], c.ev[2]));
        res.kernel_ms = ms;
        for (long long i = 0; i < B; ++i) {
            res.sum_k += hk[i];
            res.n_nc += (he[i] < 0);
        }
        res.launches = 1;
        res.block = block; res.grid = grid; res.smem = (int)smem;
        return 0;
    }

    // one device, one contiguous slice [off, off+B) of the caller's arrays
    int run_on_device(int dev, const Call &cl, Result &res) {
        DeviceCtx *pc;
        int rc = init_device(dev, &pc);
        if (rc) return rc;
        DeviceCtx &c = *pc;
        const long long B = cl.B;
        const bool varb = cl.LB != nullptr && cl.UB != nullptr;
        if (varb && !Traits::HAS_VARB) return fail(SPCIES_CUDA_EUNSUPPORTED, "this solver was generated without per-instance bounds");
        if (!cl.device_pointers && cl.sol == nullptr && B > 0 && B <= SMALL_B && cl.tail_mode != SPCIES_CUDA_TAIL_TWO_PHASE &&
            cl.tail_mode != SPCIES_CUDA_TAIL_CAPS)
            return run_small(c, cl, res);
        stop_server(c, false);       // a lingering single-instance server gives its SM back (it would exit by itself a moment later)
        cudaStream_t s = (cl.device_pointers && cl.user_stream) ? cl.user_stream : c.stream;
        bool direct_out = false;
        BatchIO io;
        memset(&io, 0, sizeof io);
        io.B = B;
        io.queue = c.d_queue;
        if (cl.device_pointers) {
            io.x0 = cl.x0; io.xr = cl.xr; io.ur = cl.ur; io.r = cl.r; io.LB = cl.LB; io.UB = cl.UB;
            for (int i = 0; i < 4; ++i) io.ex[i] = cl.ex[i];
            io.u = cl.u; io.k = cl.k; io.e = cl.e; io.sol = cl.sol;
        } else {
            if ((rc = reserve(c, B, varb, cl.sol != nullptr))) return rc;
            io.x0 = c.d_x0; io.xr = c.d_xr; io.ur = c.d_ur; io.r = c.d_r;
            for (int i = 0; i < 4; ++i) io.ex[i] = Traits::extra_width(i) > 0 ? c.d_ex[i] : nullptr;
            io.LB = varb ? c.d_LB : nullptr; io.UB = varb ? c.d_UB : nullptr;
            io.u = c.d_u; io.k = c.d_k; io.e = c.d_e; io.sol = cl.sol ? c.d_sol : nullptr;
            // results straight into the caller's arrays when those are pinned (device-mapped) host memory: 24 bytes per instance
            // posted over PCIe under the running kernel instead of three device->host copies after it
            if (direct_results()) {
                cudaPointerAttributes au, ak, ae;
                if (cudaPointerGetAttributes(&au, cl.u) == cudaSuccess && cudaPointerGetAttributes(&ak, cl.k) == cudaSuccess &&
                    cudaPointerGetAttributes(&ae, cl.e) == cudaSuccess && au.type == cudaMemoryTypeHost && ak.type == cudaMemoryTypeHost &&
                    ae.type == cudaMemoryTypeHost && au.devicePointer && ak.devicePointer && ae.devicePointer) {
                    io.u = (double *)au.devicePointer; io.k = (int *)ak.devicePointer; io.e = (int *)ae.devicePointer;
                    direct_out = true;
                }
                cudaGetLastError();   // cudaPointerGetAttributes on pageable memory may leave a sticky-free error code
            }
        }
        io.engine = cl.engine;
        int block = cl.block > 0 ? cl.block : Traits::default_block(varb);
        size_t smem = Traits::smem_bytes(block, varb);
        int ipb = block;                                 // instances resident per CTA
        Traits::engine_shape(cl.arith, io, block, smem, ipb);
        long long want = (B + ipb - 1) / ipb;
        int grid = cl.grid > 0 ? cl.grid : c.sm_count;   // persistent: one CTA per SM
        if (want < grid) grid = (int)(want > 0 ? want : 1);
        const size_t need_scratch = std::max<size_t>(Traits::uses_scratch(cl.arith, io) ? Traits::scratch_bytes(grid, block, varb) : 0,
                                                     Traits::engine_scratch_bytes(cl.arith, io, grid));
        if (need_scratch > c.cap_scratch) {
            c.cap_scratch = 0;
            if (c.d_scratch) SPCIES_CK(cudaFree(c.d_scratch));
            c.d_scratch = nullptr;
            SPCIES_CK(cudaMalloc(&c.d_scratch, need_scratch));
            c.cap_scratch = need_scratch;
        }
        // tail handling (kernels that can park & resume instances, Traits::HAS_PARK), batches larger than four waves:
        //   two_phase  park the instances still running shortly after the queue ran dry, resume them in a second launch
        //   caps       iteration-cap rounds (engines with Traits::caps_engine): launch r runs every instance it holds up to
        //              caps[r] iterations and parks the rest, so that the slow instances are all known -- and all start -- early
        //              instead of trailing the launch one by one; the last round has no cap
        bool two_phase = false;
        int caps[4] = {0, 0, 0, 0}, ncaps = 0;
        if constexpr (Traits::HAS_PARK) {
            const bool big = B > 4LL * grid * ipb;
            const bool caps_ok = Traits::caps_engine(cl.arith, io);
            int mode = cl.tail_mode;
            if (mode == SPCIES_CUDA_TAIL_AUTO) mode = !big ? SPCIES_CUDA_TAIL_SINGLE : (caps_ok ? SPCIES_CUDA_TAIL_CAPS : SPCIES_CUDA_TAIL_TWO_PHASE);
            if (mode == SPCIES_CUDA_TAIL_CAPS && !caps_ok) mode = SPCIES_CUDA_TAIL_TWO_PHASE;
            if (!Traits::park_engine(cl.arith, io)) mode = SPCIES_CUDA_TAIL_SINGLE;      // this call's engine cannot park instances
            two_phase = mode == SPCIES_CUDA_TAIL_TWO_PHASE;
            if (mode == SPCIES_CUDA_TAIL_CAPS) {
                const int dflt[3] = {96, 320, 0};
                const int *want_caps = cl.tail_caps[0] > 0 ? cl.tail_caps : dflt;
                for (int i = 0; i < 3 && want_caps[i] > 0; ++i)
                    if (want_caps[i] < Traits::K_MAX && (ncaps == 0 || want_caps[i] > caps[ncaps - 1])) caps[ncaps++] = want_caps[i];
            }
            if (two_phase || ncaps > 0) {
                long long need = (long long)grid * block;
                if (ncaps > 0) {
                    // records of the instances that outlive a cap (C2: 1.5 % at 96 iterations, 25 % at 32); instances that find the
                    // park full run on to the end in place.  SPCIES_CUDA_PARK_DIV: capacity = B / div (development knob)
                    static const long long div = [] {
                        const char *e = getenv("SPCIES_CUDA_PARK_DIV");
                        const long long v = e ? atoll(e) : 0;
                        return v >= 1 ? v : 16;
                    }();
                    need = std::max<long long>(4LL * grid * ipb, B / div + 1024);
                }
                if (need > c.cap_park) {
                    if (c.d_park) SPCIES_CK(cudaFree(c.d_park));
                    if (c.d_park2) SPCIES_CK(cudaFree(c.d_park2));
                    c.d_park = c.d_park2 = nullptr;
                    c.cap_park = 0;
                    SPCIES_CK(cudaMalloc((void **)&c.d_park, (size_t)need * Traits::PARK_DOUBLES * sizeof(double)));
                    c.cap_park = need;
                }
                if (ncaps > 1 && !c.d_park2)
                    SPCIES_CK(cudaMalloc((void **)&c.d_park2, (size_t)c.cap_park * Traits::PARK_DOUBLES * sizeof(double)));
                io.park = c.d_park;
                io.park_cap = c.cap_park;
                io.grace = cl.tail_grace > 0 ? cl.tail_grace : 32;
            }
        }
        SPCIES_CK(cudaMemsetAsync(c.d_queue, 0, QUEUE_WORDS * sizeof(unsigned long long), s));
        bool pipelined = false;
        if (!cl.device_pointers) {
            // host -> device.  Large batches: chunk by chunk on the copy stream, under the kernel; a watermark word tells
            // the kernel how many instances have arrived (first chunk = two waves of lanes, so the kernel starts at once).
            // Chunks are multiples of 32 instances: no 128-byte line is shared between two chunks.
            pipelined = B >= 8LL * grid * block;
            cudaStream_t cs = pipelined ? c.copy_stream : s;
            if (pipelined) {
                SPCIES_CK(cudaEventRecord(c.ev[5], s));
                SPCIES_CK(cudaStreamWaitEvent(cs, c.ev[5], 0));
            }
            SPCIES_CK(cudaEventRecord(c.ev[0], cs));
            long long first = pipelined ? ((2LL * grid * block + 31) / 32 * 32) : B;
            long long rest = pipelined ? (((B - first) / (MAX_CHUNKS / 2 - 1) + 31) / 32 * 32) : 0;
            if (rest < 32) rest = 32;
            int nchunk = 0;
            for (long long lo = 0; lo < B; ++nchunk) {
                const long long hi = std::min<long long>(B, lo + (nchunk == 0 ? first : rest));
                const size_t cnt = (size_t)(hi - lo);
                SPCIES_CK(cudaMemcpyAsync(c.d_x0 + lo * Traits::NN, cl.x0 + lo * Traits::NN, cnt * Traits::NN * 8, cudaMemcpyHostToDevice, cs));
                SPCIES_CK(cudaMemcpyAsync(c.d_xr + lo * Traits::NN, cl.xr + lo * Traits::NN, cnt * Traits::NN * 8, cudaMemcpyHostToDevice, cs));
                SPCIES_CK(cudaMemcpyAsync(c.d_ur + lo * Traits::MM, cl.ur + lo * Traits::MM, cnt * Traits::MM * 8, cudaMemcpyHostToDevice, cs));
                if (Traits::HAS_R) SPCIES_CK(cudaMemcpyAsync(c.d_r + lo, cl.r + lo, cnt * 8, cudaMemcpyHostToDevice, cs));
                for (int i = 0; i < 4; ++i) {
                    const int w = Traits::extra_width(i);
                    if (w > 0) SPCIES_CK(cudaMemcpyAsync(c.d_ex[i] + lo * w, cl.ex[i] + lo * w, cnt * w * 8, cudaMemcpyHostToDevice, cs));
                }
                if (varb) {
                    SPCIES_CK(cudaMemcpyAsync(c.d_LB + lo * Traits::NMM, cl.LB + lo * Traits::NMM, cnt * Traits::NMM * 8, cudaMemcpyHostToDevice, cs));
                    SPCIES_CK(cudaMemcpyAsync(c.d_UB + lo * Traits::NMM, cl.UB + lo * Traits::NMM, cnt * Traits::NMM * 8, cudaMemcpyHostToDevice, cs));
                }
                if (pipelined) {
                    c.h_ready[nchunk] = (unsigned long long)hi;
                    SPCIES_CK(cudaMemcpyAsync(c.d_queue + 8, c.h_ready + nchunk, sizeof(unsigned long long), cudaMemcpyHostToDevice, cs));
                }
                lo = hi;
            }
            SPCIES_CK(cudaEventRecord(c.ev[4], cs));
            if (pipelined) io.ready = c.d_queue + 8;
        }
        SPCIES_CK(cudaEventRecord(c.ev[1], s));
        if (B > 0) {
            io.phase = (two_phase || ncaps > 0) ? 1 : 0;
            io.grace = two_phase ? io.grace : (1 << 30);
            io.cap = ncaps > 0 ? caps[0] : 0;
            SPCIES_CK(Traits::launch(cl.arith, varb, grid, block, smem, s, io, c.d_consts, c.d_scratch));
            res.launches = 1;
            if constexpr (Traits::HAS_PARK) {
                const int rounds = two_phase ? 1 : ncaps;          // resume launches
                for (int r = 0; r < rounds; ++r) {
                    // the records parked by the previous launch become this launch's queue
                    SPCIES_CK(cudaMemcpyAsync(c.d_queue + 9, c.d_queue + 6, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
                    if (ncaps > 0) SPCIES_CK(cudaMemsetAsync(c.d_queue + 6, 0, 2 * sizeof(unsigned long long), s));
                    io.phase = 2;
                    io.park_in = (r % 2 == 0) ? c.d_park : c.d_park2;
                    io.park_in_cap = c.cap_park;
                    io.park = (r + 1 < rounds) ? ((r % 2 == 0) ? c.d_park2 : c.d_park) : (two_phase ? c.d_park : nullptr);
                    io.cap = (r + 1 < rounds) ? caps[r + 1] : 0;
                    io.ready = nullptr;
                    int block2 = Traits::resume_block(varb), ipb2 = 0;
                    size_t smem2 = Traits::smem_bytes(block2, varb);
                    Traits::engine_shape(cl.arith, io, block2, smem2, ipb2);
                    SPCIES_CK(Traits::launch(cl.arith, varb, c.sm_count, block2, smem2, s, io, c.d_consts, c.d_scratch));
                    res.launches += 1;
                }
            }
        }
        SPCIES_CK(cudaEventRecord(c.ev[2], s));
        if (pipelined) SPCIES_CK(cudaStreamWaitEvent(s, c.ev[4], 0));
        unsigned long long stats[QUEUE_WORDS] = {0};
        SPCIES_CK(cudaMemcpyAsync(stats, c.d_queue, sizeof stats, cudaMemcpyDeviceToHost, s));
        if (!cl.device_pointers) {
            if (!direct_out) {
                SPCIES_CK(cudaMemcpyAsync(cl.u, c.d_u, (size_t)B * Traits::MM * 8, cudaMemcpyDeviceToHost, s));
                SPCIES_CK(cudaMemcpyAsync(cl.k, c.d_k, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
                SPCIES_CK(cudaMemcpyAsync(cl.e, c.d_e, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
            }
            if (cl.sol)
                SPCIES_CK(cudaMemcpyAsync(cl.sol, c.d_sol, (size_t)B * Traits::SOL_DOUBLES * 8, cudaMemcpyDeviceToHost, s));
        }
        SPCIES_CK(cudaEventRecord(c.ev[3], s));
        SPCIES_CK(cudaStreamSynchronize(s));
        SPCIES_CK(cudaGetLastError());
        float ms = 0;
        SPCIES_CK(cudaEventElapsedTime(&ms, c.ev[1], c.ev[2]));
        res.kernel_ms = ms;
        if (!cl.device_pointers) {
            SPCIES_CK(cudaEventElapsedTime(&ms, c.ev[0], c.ev[4]));
            res.h2d_ms = ms;
            SPCIES_CK(cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]));
            res.d2h_ms = ms;
        }
        res.sum_k = (long long)stats[1];
        res.n_nc = (long long)stats[2];
        res.parked = ncaps > 0 ? (long long)stats[10] : (long long)stats[6];
        if (stats[4] && stats[5]) {
            const unsigned long long t_start = ~stats[4], t_drain = stats[3] ? ~stats[3] : stats[5];
            res.span_us = (int)((stats[5] - t_start) / 1000ULL);
            res.drain_us = (int)((t_drain - t_start) / 1000ULL);
        }
        res.block = block; res.grid = grid; res.smem = (int)smem;
        return 0;
    }

    // ---- closed loop -------------------------------------------------------------------------------------------------------------
    struct ClCall {
        long long B, ld;                       // instances of this device's slice; instances per sampling time in the caller's arrays
        int steps;
        const double *x0, *xr, *ur, *r, *plant;
        double *x_traj, *u_traj;
        int *k_traj, *e_traj;
        int arith, engine, warm;
        bool model_plant;
    };
    int cl_on_device(int dev, const ClCall &cl, Result &res) {
        DeviceCtx *pc;
        int rc = init_device(dev, &pc);
        if (rc) return rc;
        DeviceCtx &c = *pc;
        const long long B = cl.B;
        const int steps = cl.steps;
        if ((rc = reserve(c, B, false, false))) return rc;
        const long long need = (long long)(steps + 1) * B;
        if (need > c.cap_cl) {
            c.cap_cl = 0;
            if ((rc = grow(&c.d_clx, need * Traits::NN))) return rc;
            if ((rc = grow(&c.d_clu, need * Traits::MM))) return rc;
            if ((rc = grow(&c.d_clk, need))) return rc;
            if ((rc = grow(&c.d_cle, need))) return rc;
            c.cap_cl = need;
        }
        if (!c.d_plant) SPCIES_CK(cudaMalloc((void **)&c.d_plant, sizeof(double) * Traits::NN * Traits::NMM));
        cudaStream_t s = c.stream;
        SPCIES_CK(cudaEventRecord(c.ev[0], s));
        SPCIES_CK(cudaMemcpyAsync(c.d_plant, cl.plant, sizeof(double) * Traits::NN * Traits::NMM, cudaMemcpyHostToDevice, s));
        SPCIES_CK(cudaMemcpyAsync(c.d_clx, cl.x0, (size_t)B * Traits::NN * 8, cudaMemcpyHostToDevice, s));
        SPCIES_CK(cudaMemcpyAsync(c.d_xr, cl.xr, (size_t)B * Traits::NN * 8, cudaMemcpyHostToDevice, s));
        SPCIES_CK(cudaMemcpyAsync(c.d_ur, cl.ur, (size_t)B * Traits::MM * 8, cudaMemcpyHostToDevice, s));
        if (Traits::HAS_R) SPCIES_CK(cudaMemcpyAsync(c.d_r, cl.r, (size_t)B * 8, cudaMemcpyHostToDevice, s));
        BatchIO io;
        memset(&io, 0, sizeof io);
        io.B = B;
        io.queue = c.d_queue;
        io.xr = c.d_xr; io.ur = c.d_ur; io.r = c.d_r;
        io.engine = cl.engine;
        io.grace = 1 << 30;
        int block = Traits::default_block(false), ipb = block;
        size_t smem = Traits::smem_bytes(block, false);
        Traits::engine_shape(cl.arith, io, block, smem, ipb);
        long long want = (B + ipb - 1) / ipb;
        int grid = c.sm_count;
        if (want < grid) grid = (int)(want > 0 ? want : 1);
        const size_t need_scratch = std::max<size_t>(Traits::uses_scratch(cl.arith, io) ? Traits::scratch_bytes(grid, block, false) : 0,
                                                     Traits::engine_scratch_bytes(cl.arith, io, grid));
        if (need_scratch > c.cap_scratch) {
            c.cap_scratch = 0;
            if (c.d_scratch) SPCIES_CK(cudaFree(c.d_scratch));
            c.d_scratch = nullptr;
            SPCIES_CK(cudaMalloc(&c.d_scratch, need_scratch));
            c.cap_scratch = need_scratch;
        }
        std::vector<unsigned long long> stats((size_t)QUEUE_WORDS * (steps > 0 ? steps : 1), 0ULL);
        SPCIES_CK(cudaEventRecord(c.ev[1], s));
        const bool in_kernel = cl.model_plant && Traits::cl_engine(cl.arith, io) && B > 0 && steps > 0;
        if (in_kernel) {
            // the engine keeps every instance on chip for the whole run: one launch
            io.x0 = c.d_clx; io.u = c.d_clu; io.k = c.d_clk; io.e = c.d_cle;
            io.cl_steps = steps; io.cl_warm = cl.warm; io.cl_ld = B; io.cl_x = c.d_clx;
            SPCIES_CK(cudaMemsetAsync(c.d_queue, 0, QUEUE_WORDS * sizeof(unsigned long long), s));
            SPCIES_CK(Traits::launch(cl.arith, false, grid, block, smem, s, io, c.d_consts, c.d_scratch));
            SPCIES_CK(cudaMemcpyAsync(stats.data(), c.d_queue, QUEUE_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            res.launches = 1;
        } else if (B > 0) {
            for (int t = 0; t < steps; ++t) {
                io.x0 = c.d_clx + (size_t)t * B * Traits::NN;
                io.u = c.d_clu + (size_t)t * B * Traits::MM;
                io.k = c.d_clk + (size_t)t * B;
                io.e = c.d_cle + (size_t)t * B;
                SPCIES_CK(cudaMemsetAsync(c.d_queue, 0, QUEUE_WORDS * sizeof(unsigned long long), s));
                SPCIES_CK(Traits::launch(cl.arith, false, grid, block, smem, s, io, c.d_consts, c.d_scratch));
                SPCIES_CK(cudaMemcpyAsync(stats.data() + (size_t)t * QUEUE_WORDS, c.d_queue, QUEUE_WORDS * sizeof(unsigned long long),
                                          cudaMemcpyDeviceToHost, s));
                cl_plant_kernel<Traits::NN, Traits::MM><<<(unsigned)((B + 127) / 128), 128, 0, s>>>(
                    B, c.d_plant, io.x0, io.u, c.d_clx + (size_t)(t + 1) * B * Traits::NN);
                SPCIES_CK(cudaGetLastError());
                res.launches += 2;
            }
        }
        SPCIES_CK(cudaEventRecord(c.ev[2], s));
        if (B > 0 && steps > 0) {
            const size_t w8 = (size_t)B * 8, w4 = (size_t)B * 4;
            if (cl.x_traj)
                SPCIES_CK(cudaMemcpy2DAsync(cl.x_traj, (size_t)cl.ld * Traits::NN * 8, c.d_clx, w8 * Traits::NN, w8 * Traits::NN, steps + 1,
                                            cudaMemcpyDeviceToHost, s));
            SPCIES_CK(cudaMemcpy2DAsync(cl.u_traj, (size_t)cl.ld * Traits::MM * 8, c.d_clu, w8 * Traits::MM, w8 * Traits::MM, steps,
                                        cudaMemcpyDeviceToHost, s));
            SPCIES_CK(cudaMemcpy2DAsync(cl.k_traj, (size_t)cl.ld * 4, c.d_clk, w4, w4, steps, cudaMemcpyDeviceToHost, s));
            SPCIES_CK(cudaMemcpy2DAsync(cl.e_traj, (size_t)cl.ld * 4, c.d_cle, w4, w4, steps, cudaMemcpyDeviceToHost, s));
        }
        SPCIES_CK(cudaEventRecord(c.ev[3], s));
        SPCIES_CK(cudaStreamSynchronize(s));
        SPCIES_CK(cudaGetLastError());
        float ms = 0;
        SPCIES_CK(cudaEventElapsedTime(&ms, c.ev[1], c.ev[2]));
        res.kernel_ms = ms;
        SPCIES_CK(cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]));
        res.h2d_ms = ms;
        SPCIES_CK(cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]));
        res.d2h_ms = ms;
        for (int t = 0; t < (in_kernel ? 1 : steps); ++t) {
            res.sum_k += (long long)stats[(size_t)t * QUEUE_WORDS + 1];
            res.n_nc += (long long)stats[(size_t)t * QUEUE_WORDS + 2];
        }
        res.block = block; res.grid = grid; res.smem = (int)smem;
        return 0;
    }

    int run_closed_loop(long long B, int steps, const double *x0, const double *xr, const double *ur, const double *r, double *x_traj,
                        double *u_traj, int *k_traj, int *e_traj, const spcies_batch_opts *opts, spcies_batch_info *info,
                        const double *model_AB) {
        auto t0 = std::chrono::steady_clock::now();
        spcies_batch_opts o;
        memset(&o, 0, sizeof o);
        if (opts) o = *opts;
        if (B < 0 || steps < 0) return fail(SPCIES_CUDA_EINVAL, "B < 0 or steps < 0");
        if (o.warm_start < 0 || o.warm_start > 2) return fail(SPCIES_CUDA_EINVAL, "warm_start must be 0, 1 or 2");
        if (Traits::extra_width(0) > 0) return fail(SPCIES_CUDA_EUNSUPPORTED, "closed loop: not available for solvers with extra inputs");
        if (B > 0 && steps > 0 && (!x0 || !xr || !ur || !u_traj || !k_traj || !e_traj || (Traits::HAS_R && !r)))
            return fail(SPCIES_CUDA_EINVAL, "NULL array argument");
        if (o.device_pointers || o.LB || o.UB) return fail(SPCIES_CUDA_EUNSUPPORTED, "closed loop: host arrays, generated bounds");
#if defined(in_engineering) && in_engineering == 1
        // the prediction model acts on scaled, incremental variables; the states of the simulation are in engineering units
        if (o.plant_AB == nullptr)
            return fail(SPCIES_CUDA_EUNSUPPORTED, "closed loop of an in_engineering solver: give the plant in engineering units (opts.plant_AB)");
#endif
        if (o.arith != SPCIES_CUDA_ARITH_FAST && o.arith != SPCIES_CUDA_ARITH_EXACT) return fail(SPCIES_CUDA_EINVAL, "unknown arith mode");
        int ndev = o.n_devices > 1 ? o.n_devices : 1;
        int avail = device_count();
        if (avail <= 0) return fail(SPCIES_CUDA_ENODEVICE, "no usable CUDA device (this library has no CPU fallback)");
        if (o.device < 0 || o.device + ndev > avail) return fail(SPCIES_CUDA_ENODEVICE, "requested CUDA devices do not exist");
        {
            std::lock_guard<std::mutex> lk(mu);
            if ((int)ctx.size() < avail) ctx.resize(avail);
        }
        std::vector<Result> res(ndev);
        std::vector<ClCall> calls(ndev);
        const long long per = (B + ndev - 1) / ndev;
        for (int d = 0; d < ndev; ++d) {
            long long lo = std::min<long long>(B, d * per), hi = std::min<long long>(B, lo + per);
            ClCall &c = calls[d];
            c.B = hi - lo; c.ld = B; c.steps = steps;
            c.x0 = x0 + lo * Traits::NN; c.xr = xr + lo * Traits::NN; c.ur = ur + lo * Traits::MM; c.r = r ? r + lo : nullptr;
            c.plant = o.plant_AB ? o.plant_AB : model_AB;
            c.model_plant = o.plant_AB == nullptr;
            c.x_traj = x_traj ? x_traj + lo * Traits::NN : nullptr;
            c.u_traj = u_traj + lo * Traits::MM; c.k_traj = k_traj + lo; c.e_traj = e_traj + lo;
            c.arith = o.arith; c.engine = o.engine; c.warm = o.warm_start;
        }
        if (o.warm_start) {
            BatchIO probe;
            memset(&probe, 0, sizeof probe);
            probe.engine = o.engine;
            if (!Traits::cl_engine(o.arith, probe) || o.plant_AB)
                return fail(SPCIES_CUDA_EUNSUPPORTED, "warm start needs the FISTA tensor-core engine (FAST arithmetic, model plant)");
        }
        if (ndev == 1) {
            res[0].rc = cl_on_device(o.device, calls[0], res[0]);
        } else {
            std::vector<std::thread> th;
            std::vector<std::string> errs(ndev);
            for (int d = 0; d < ndev; ++d)
                th.emplace_back([&, d] {
                    res[d].rc = cl_on_device(o.device + d, calls[d], res[d]);
                    if (res[d].rc) errs[d] = g_last_error;
                });
            for (auto &t : th) t.join();
            for (int d = 0; d < ndev; ++d)
                if (res[d].rc) snprintf(g_last_error, sizeof g_last_error, "device %d: %s", o.device + d, errs[d].c_str());
        }
        for (int d = 0; d < ndev; ++d)
            if (res[d].rc) return res[d].rc;
        if (info) {
            memset(info, 0, sizeof *info);
            for (int d = 0; d < ndev; ++d) {
                info->kernel_ms = std::max(info->kernel_ms, res[d].kernel_ms);
                info->h2d_ms = std::max(info->h2d_ms, res[d].h2d_ms);
                info->d2h_ms = std::max(info->d2h_ms, res[d].d2h_ms);
                info->sum_k += res[d].sum_k;
                info->n_not_converged += res[d].n_nc;
                info->launches += res[d].launches;
            }
            info->block_threads = res[0].block; info->grid_blocks = res[0].grid; info->smem_bytes = res[0].smem;
            info->n_devices = ndev;
            info->h2d_bytes = B * 8 * (2 * Traits::NN + Traits::MM + (Traits::HAS_R ? 1 : 0));
            info->d2h_bytes = B * (long long)steps * (8 * Traits::MM + 8) + (x_traj ? B * (long long)(steps + 1) * 8 * Traits::NN : 0);
            info->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        return 0;
    }

    // extra: the four additional per-instance inputs of solvers with Traits::extra_width(i) > 0 (ellipHMPC: x_rs, x_rc, u_rs, u_rc;
    // TIME_VARYING: A, B, Q, R), else nullptr
    int run(long long B, const double *x0, const double *xr, const double *ur, const double *r, double *u, int *k,
            int *e, double *sol, const spcies_batch_opts *opts, spcies_batch_info *info, const double *const *extra = nullptr) {
        auto t0 = std::chrono::steady_clock::now();
        spcies_batch_opts o;
        memset(&o, 0, sizeof o);
        if (opts) o = *opts;
        if (B < 0) return fail(SPCIES_CUDA_EINVAL, "B < 0");
        if (B > 0 && (!x0 || !xr || !ur || !u || !k || !e || (Traits::HAS_R && !r)))
            return fail(SPCIES_CUDA_EINVAL, "NULL array argument");
        if (B > 0 && Traits::extra_width(0) > 0 && (!extra || !extra[0] || !extra[1] || !extra[2] || !extra[3]))
            return fail(SPCIES_CUDA_EINVAL, "NULL array argument (references / model)");
        if (B > 0 && Traits::TV && (!o.LB || !o.UB)) return fail(SPCIES_CUDA_EINVAL, "a TIME_VARYING solver needs LB and UB");
        if ((o.LB == nullptr) != (o.UB == nullptr)) return fail(SPCIES_CUDA_EINVAL, "LB and UB must be given together");
        if (o.arith != SPCIES_CUDA_ARITH_FAST && o.arith != SPCIES_CUDA_ARITH_EXACT)
            return fail(SPCIES_CUDA_EINVAL, "unknown arith mode");
        int ndev = o.n_devices > 1 ? o.n_devices : 1;
        if (o.device_pointers && ndev > 1) return fail(SPCIES_CUDA_EINVAL, "device_pointers needs n_devices <= 1");
        int avail = device_count();
        if (avail <= 0) return fail(SPCIES_CUDA_ENODEVICE, "no usable CUDA device (this library has no CPU fallback)");
        if (o.device < 0 || o.device + ndev > avail) return fail(SPCIES_CUDA_ENODEVICE, "requested CUDA devices do not exist");
        {
            std::lock_guard<std::mutex> lk(mu);
            if ((int)ctx.size() < avail) ctx.resize(avail);
        }
        std::vector<Result> res(ndev);
        std::vector<Call> calls(ndev);
        const long long per = (B + ndev - 1) / ndev;
        for (int d = 0; d < ndev; ++d) {
            long long lo = std::min<long long>(B, d * per), hi = std::min<long long>(B, lo + per);
            Call &c = calls[d];
            c.B = hi - lo;
            c.x0 = x0 + lo * Traits::NN; c.xr = xr + lo * Traits::NN; c.ur = ur + lo * Traits::MM;
            c.r = r ? r + lo : nullptr;
            for (int i = 0; i < 4; ++i)
                if (Traits::extra_width(i) > 0 && extra) c.ex[i] = extra[i] + lo * Traits::extra_width(i);
            c.LB = o.LB ? o.LB + lo * Traits::NMM : nullptr; c.UB = o.UB ? o.UB + lo * Traits::NMM : nullptr;
            c.u = u + lo * Traits::MM; c.k = k + lo; c.e = e + lo;
            c.sol = sol ? sol + lo * (long long)Traits::SOL_DOUBLES : nullptr;
            c.arith = o.arith; c.block = o.block_threads; c.grid = o.grid_blocks;
            c.tail_mode = o.tail_mode; c.tail_grace = o.tail_grace; c.engine = o.engine;
            for (int i = 0; i < 3; ++i) c.tail_caps[i] = o.tail_caps[i];
            c.device_pointers = o.device_pointers != 0;
            c.user_stream = (cudaStream_t)o.stream;
        }
        if (ndev == 1) {
            res[0].rc = run_on_device(o.device, calls[0], res[0]);
        } else {
            std::vector<std::thread> th;
            std::vector<std::string> errs(ndev);
            for (int d = 0; d < ndev; ++d)
                th.emplace_back([&, d] {
                    res[d].rc = run_on_device(o.device + d, calls[d], res[d]);
                    if (res[d].rc) errs[d] = g_last_error;
                });
            for (auto &t : th) t.join();
            for (int d = 0; d < ndev; ++d)
                if (res[d].rc) snprintf(g_last_error, sizeof g_last_error, "device %d: %s", o.device + d, errs[d].c_str());
        }
        for (int d = 0; d < ndev; ++d)
            if (res[d].rc) return res[d].rc;
        if (info) {
            memset(info, 0, sizeof *info);
            for (int d = 0; d < ndev; ++d) {
                info->kernel_ms = std::max(info->kernel_ms, res[d].kernel_ms);
                info->h2d_ms = std::max(info->h2d_ms, res[d].h2d_ms);
                info->d2h_ms = std::max(info->d2h_ms, res[d].d2h_ms);
                info->sum_k += res[d].sum_k;
                info->n_not_converged += res[d].n_nc;
                info->launches += res[d].launches;
                info->parked += (int)res[d].parked;
                info->drain_us = std::max(info->drain_us, res[d].drain_us);
                info->span_us = std::max(info->span_us, res[d].span_us);
            }
            info->block_threads = res[0].block; info->grid_blocks = res[0].grid; info->smem_bytes = res[0].smem;
            info->n_devices = ndev;
            if (!o.device_pointers) {
                const long long varb = (o.LB ? 2LL * Traits::NMM : 0);
                info->h2d_bytes = B * 8 * (2 * Traits::NN + Traits::MM + (Traits::HAS_R ? 1 : 0) + varb + Traits::extra_width(0) + Traits::extra_width(1) +
                                           Traits::extra_width(2) + Traits::extra_width(3));
                info->d2h_bytes = B * (8 * Traits::MM + 8) + (sol ? B * 8LL * Traits::SOL_DOUBLES : 0);
            }
            static int regs_cache[2][2] = {{0, 0}, {0, 0}};          // cudaFuncGetAttributes costs microseconds: once per variant
            int &rcache = regs_cache[o.arith == SPCIES_CUDA_ARITH_EXACT ? 1 : 0][o.LB != nullptr ? 1 : 0];
            if (rcache == 0) {
                cudaFuncAttributes fa;
                if (Traits::attributes(o.arith, o.LB != nullptr, &fa) == cudaSuccess) rcache = fa.numRegs;
            }
            info->regs_per_thread = rcache;
            info->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        return 0;
    }
};

}  // namespace spcies
