// spcies_dense_mma.cuh -- generic FP64 tensor-core engine for the ADMM solvers whose equality-constrained QP step is one
// linear map shared by the whole batch:
//
//     y = F [w ; c],      then a component-wise / cone-wise update of the iterates from y
//
// (w: NW tiles recomputed every iteration from the iterates, c: NC tiles fixed per instance -- the parts of the right-hand side
// that carry x0 / xr / ur --, y: NO tiles).  The reference evaluates F as a chain of sparse products with run-time index arrays
// (CSR mat-vec, CSC L D L' solve, CSR mat-vec: code_MPCT_ADMM_cs_C.c:111-165, code_HMPC_ADMM_C.c:122-160); for a batch that
// shares the model the chain is a GEMM with the batch as M.
//
// Mapping (spcies_mma.cuh): 8 instances = the 8 rows of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) = one *group*; a vector is a
// sequence of tiles of 8 columns, lane (g, t) of a group holds columns 2t, 2t+1 of instance g (the C/D and, one register per
// k-step, the A fragment layout).  F is a table of B fragments.
//
// Data movement, sm_100a:
//   * the fragment table is laid out in consumption order, [pass][input chunk][input tile in chunk][output tile in pass][lane],
//     so one iteration reads it front to back.  A *producer warp* streams it L2 -> shared memory with bulk asynchronous copies
//     (cp.async.bulk, SASS UBLKCP) into a ring of NSTAGE stages, completion on `full` mbarriers (expect_tx); consumer warps
//     release a stage on its `empty` mbarrier.  One stream feeds every consumer warp of the CTA (the per-warp rings of round 1
//     read the table once per warp).  A table that fits beside the iterates is loaded once and stays resident (RESIDENT).
//   * the iterates of a group live in shared memory, [tile][lane] double2 (one conflict-free LDS.128 / STS.128 per lane).
//   * long horizons leave room for only two groups per SM: a group can be served by a *team* of TEAM warps (one per SM
//     scheduler) that split the output passes and the component-wise work and meet at a named barrier.
//
// A solver provides an engine policy E (see MPCT_ADMM_cs_mma.cuh for the smallest one):
//   constants  NO, NW, NC, NSTATE (iterate tiles per group), NB (output tiles per pass = independent accumulators),
//              TEAM, OK (shape supported), struct Small (per-solver table staged in shared memory), struct Lane (per-lane
//              registers of the instance: its q entries, ...)
//   host       static void fill(const spcies_consts &, Small &, long double *F /* [NO*8][(NW+NC)*8] row-major, tile coords */)
//   device     init(Lane &, C, Small, io, inst, st, t4)      read the inputs, zero the iterates, write the c tiles
//              make_w(Lane &, C, Small, t, st) -> double2    input tile t
//              update(Lane &, C, Small, t0, acc[NB][2], st, over)   consume the NB output tiles of the pass that starts at t0
//              finish(Lane &, C, io, inst, st, t4)           write u_opt
// Arithmetic: FAST only (explicit F, FMA, MMA accumulation order); EXACT, float and the debug payload use the scalar kernel.
#pragma once
#include <type_traits>
#include <utility>
#include "spcies_kernel.cuh"
#include "spcies_mma.cuh"

namespace spcies {
namespace dense {

// Producers of the fragment stream: PRODUCER_WARPS warps x PRODUCER_LANES lanes, producer p feeding the ring slots p, p + P, ...
// One thread's loop (wait for a free slot, expect_tx, bulk copy: a chain of shared-memory / mbarrier round trips, ~400 cycles per
// 4 KB stage) was what bounded the streamed engines -- not L2, not the ring depth.  HMPC N = 50, 16 Ki instances
// (tools/engine_variants.py): 1 thread 219.6 ms; 2 / 4 lanes of one warp 147.9 / 154.4 ms (the lanes' spin loops serialise);
// 2 / 4 / 8 warps with one lane each 129.9 / 111.1 / 111.1 ms; 2 warps x 2 lanes 115.6 ms.  Four warps: the consumers' MMAs bound it.
#ifndef SPCIES_DENSE_PRODUCER_LANES
#define SPCIES_DENSE_PRODUCER_LANES 1
#endif
constexpr int PRODUCER_LANES = SPCIES_DENSE_PRODUCER_LANES;       // 1, 2, 4 or 8: divides the 8 slots of the ring
#ifndef SPCIES_DENSE_PRODUCER_WARPS
#define SPCIES_DENSE_PRODUCER_WARPS 4
#endif
constexpr int PRODUCER_WARPS = SPCIES_DENSE_PRODUCER_WARPS;       // producers = warps x lanes, each feeding every (warps x lanes)-th slot
constexpr int PRODUCERS = PRODUCER_WARPS * PRODUCER_LANES;
static_assert(8 % PRODUCERS == 0, "the producers split the 8 slots of the ring evenly");
constexpr size_t SMEM_LIMIT = 227 * 1024 - 256;
constexpr int TILE_BYTES = 32 * sizeof(double2);   // 512

// Optional members of an engine policy (detected):
//   E::BLOB_OFFSET    position of (Small | fragments) in the device constant blob; default: right behind spcies_consts
//   E::WARMUP         passes that come before the first counted iteration and have no exit test (FISTA: 1); default 0
//   E::Lane::pass     if present, the engine keeps it equal to the index of the current pass of the instance (0, 1, ...) --
//                     for policies whose step depends on the iteration number
template <class E, class = void> struct BlobOffset {
    static constexpr size_t value = (sizeof(spcies_consts) + 15) / 16 * 16;
};
template <class E> struct BlobOffset<E, std::void_t<decltype(E::BLOB_OFFSET)>> {
    static constexpr size_t value = E::BLOB_OFFSET;
};
template <class E, class = void> struct Warmup {
    static constexpr int value = 0;
};
template <class E> struct Warmup<E, std::void_t<decltype(E::WARMUP)>> {
    static constexpr int value = E::WARMUP;
};
template <class L, class = void> struct HasPass : std::false_type {};
template <class L> struct HasPass<L, std::void_t<decltype(std::declval<L &>().pass)>> : std::true_type {};

template <class E> struct Plan {
    static constexpr int NB = E::NB;
    static constexpr int KI = 2;                                     // input tiles per stage
    static constexpr int NIN = E::NW + E::NC;
    static constexpr int NCH = (NIN + 2 * KI - 1) / (2 * KI) * 2;    // chunks (stages) per pass: even, the consumer loop handles them in pairs
    static constexpr int NIN_PAD = NCH * KI;
    static constexpr int NPASS = (E::NO + NB - 1) / NB;
    static constexpr int STAGE_TILES = KI * NB;
    static constexpr size_t STAGE_BYTES = (size_t)STAGE_TILES * TILE_BYTES;
    static constexpr int TEAM = E::TEAM;
    // the warps of a team take the passes round-robin; their stages interleave in the stream (round, chunk, rank) so that the
    // ranks consume concurrently and a warp's next stage is always TEAM stages ahead (well inside the ring: no dead lock)
    static constexpr int NROUND = (NPASS + TEAM - 1) / TEAM;
    static constexpr int STAGES_PER_ITER = NROUND * NCH * TEAM;
    static constexpr size_t FRAG_BYTES = (size_t)STAGES_PER_ITER * STAGE_BYTES;
    static constexpr int GROUP_TILES = E::NSTATE + NIN_PAD;          // iterates, then the input vector [w ; c ; padding]
    static constexpr size_t GROUP_BYTES = (size_t)GROUP_TILES * TILE_BYTES;
    static constexpr size_t SMALL_BYTES = (sizeof(typename E::Small) + 15) / 16 * 16;
    static constexpr size_t CTRL_BYTES = 1536;                       // mbarriers, flags, team mailboxes
    static constexpr size_t FIXED = SMALL_BYTES + CTRL_BYTES;
    // resident table: at least 4 groups must fit beside it
    static constexpr bool RESIDENT = FIXED + FRAG_BYTES + 4 * GROUP_BYTES <= SMEM_LIMIT && STAGES_PER_ITER <= 96;
    // a streamed table goes through a ring of 8 stages (a deeper ring was measured on HMPC N = 50: 11 stages, no gain -- the
    // stream is not bound by the bytes in flight); Ctrl::empty has one barrier per ring slot
    static constexpr int NSTAGE = RESIDENT ? STAGES_PER_ITER : 8;
    static constexpr size_t RING_BYTES = (size_t)NSTAGE * STAGE_BYTES;
    static constexpr int GROUPS_RAW = FIXED + RING_BYTES >= SMEM_LIMIT ? 0 : (int)((SMEM_LIMIT - FIXED - RING_BYTES) / GROUP_BYTES);
    static constexpr int GROUPS_MAX = 8 / TEAM;
    static constexpr int GROUPS = GROUPS_RAW > GROUPS_MAX ? GROUPS_MAX : GROUPS_RAW;
    static constexpr int CONSUMER_WARPS = GROUPS * TEAM;
    static constexpr int BLOCK = (CONSUMER_WARPS + PRODUCER_WARPS) * 32;   // + the producer warp(s)
    static constexpr int IPB = GROUPS * 8;
    static constexpr size_t SMEM = FIXED + RING_BYTES + (size_t)GROUPS * GROUP_BYTES;
    static constexpr size_t OFF_SMALL = BlobOffset<E>::value;        // blob: spcies_consts [| other tables] | Small | fragments
    static constexpr size_t OFF_FRAG = OFF_SMALL + SMALL_BYTES;
    static constexpr size_t BLOB_BYTES = OFF_FRAG + FRAG_BYTES;
    static constexpr bool HAS = E::OK && sizeof(SPCIES_REAL) == 8 && GROUPS >= 1 && NSTAGE <= 96;
};

// ---- mbarrier helpers (shared::cta) --------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *b, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity) {
    while (!mbar_try_wait(b, parity)) {
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct Ctrl {                       // first bytes of the dynamic shared memory
    uint64_t full[96];
    uint64_t empty[8];
    int alive[4];                   // alive[iter % 3]: some consumer warp still has work after iteration iter
    int stop;                       // set by the consumers when the CTA is done
    int final_stages;               // stages consumed in total (valid once stop is set)
    long long slot[8][8];           // team mailbox: instance index pulled by rank 0 for (group, instance in group)
    unsigned over[8][2];            // team mailbox: exit-test ballots
};
static_assert(sizeof(Ctrl) <= 1536, "Ctrl");
static_assert(sizeof(((Ctrl *)nullptr)->empty) / sizeof(uint64_t) == 8, "one `empty` barrier per slot of the streaming ring");

// ---- host: dense map -> fragment table in consumption order --------------------------------------------------------------------
template <class E> static inline void fill_fragments(const long double *F, double2 *frag) {
    typedef Plan<E> P;
    constexpr int NINC = P::NIN * 8;
    for (int round = 0; round < P::NROUND; ++round)
        for (int ch = 0; ch < P::NCH; ++ch)
            for (int r = 0; r < P::TEAM; ++r)
                for (int ki = 0; ki < P::KI; ++ki)
                    for (int b = 0; b < P::NB; ++b)
                        for (int lane = 0; lane < 32; ++lane) {
                            const int pass = round * P::TEAM + r;           // passes beyond NPASS are zero padding
                            const int ot = pass * P::NB + b, it = ch * P::KI + ki, o = lane / 4, t = lane % 4;
                            double v[2] = {0.0, 0.0};
                            if (pass < P::NPASS && ot < E::NO && it < P::NIN)
                                for (int i = 0; i < 2; ++i) v[i] = (double)F[(size_t)(ot * 8 + o) * NINC + it * 8 + 2 * t + i];
                            const size_t stage = ((size_t)round * P::NCH + ch) * P::TEAM + r;
                            frag[((stage * P::KI + ki) * P::NB + b) * 32 + lane] = make_double2(v[0], v[1]);
                        }
}

// ---- the kernel ----------------------------------------------------------------------------------------------------------------
template <class E> __global__ void __launch_bounds__(Plan<E>::BLOCK, 1) dense_mma_kernel(const BatchIO io, const unsigned char *__restrict__ g_blob) {
    typedef Plan<E> P;
    using mma::dmma;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NB = P::NB, KI = P::KI, TEAM = P::TEAM;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Ctrl *ctrl = reinterpret_cast<Ctrl *>(smem_raw);
    const typename E::Small *S = reinterpret_cast<const typename E::Small *>(smem_raw + P::CTRL_BYTES);
    unsigned char *ring = smem_raw + P::FIXED;
    unsigned char *groups = ring + P::RING_BYTES;
    const spcies_consts *C = reinterpret_cast<const spcies_consts *>(g_blob);
    const unsigned char *g_frag = g_blob + P::OFF_FRAG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // every CTA starts the (circular) fragment stream at a different round: the order of the output passes within an iteration is
    // free (every pass reads the same w), and 148 SMs walking one table in step would all hit the same L2 lines at the same time
    const int r0 = P::RESIDENT ? 0 : (int)(blockIdx.x % P::NROUND);

    if (threadIdx.x == 0) {
        for (int s = 0; s < P::NSTAGE; ++s) mbar_init(&ctrl->full[s], 1);
        for (int s = 0; s < (P::RESIDENT ? 0 : P::NSTAGE); ++s) mbar_init(&ctrl->empty[s], P::GROUPS);
        ctrl->alive[0] = ctrl->alive[1] = ctrl->alive[2] = 0;
        ctrl->stop = 0;
        ctrl->final_stages = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the per-solver table: plain loads (it is a few KB)
    for (unsigned i = threadIdx.x; i < P::SMALL_BYTES / 16; i += blockDim.x)
        reinterpret_cast<int4 *>(smem_raw + P::CTRL_BYTES)[i] = reinterpret_cast<const int4 *>(g_blob + P::OFF_SMALL)[i];
    __syncthreads();

    // =============================================== producer warp ===============================================
    if (warp >= P::CONSUMER_WARPS) {
        const int pid = (warp - P::CONSUMER_WARPS) * PRODUCER_LANES + lane;     // producer index of this lane
        if (P::RESIDENT) {
            if (pid == 0 && lane == 0) {
                for (int s = 0; s < P::NSTAGE; ++s) {
                    mbar_expect_tx(&ctrl->full[s], (unsigned)P::STAGE_BYTES);
                    bulk_g2s(ring + (size_t)s * P::STAGE_BYTES, g_frag + (size_t)s * P::STAGE_BYTES, (unsigned)P::STAGE_BYTES, &ctrl->full[s]);
                }
                for (int s = 0; s < P::NSTAGE; ++s) mbar_wait(&ctrl->full[s], 0);     // never leave with copies in flight
            }
        } else if (lane < PRODUCER_LANES) {
            // lane l feeds the ring slots l, l + PRODUCER_LANES, ...: the wait for a free slot, the expect_tx and the copy of a stage are
            // a chain of shared-memory round trips, several lanes keep several of them in flight
            int issued = 0, slot = pid, idx = (r0 * P::NCH * TEAM + pid) % P::STAGES_PER_ITER;   // idx: stage of the table (the stream starts at round r0)
            unsigned par = 1;                          // parity of the `empty` phase to wait for (1: passes on a fresh barrier)
            bool stopped = false;
            for (;;) {
                while (!mbar_try_wait(&ctrl->empty[slot], par)) {
                    if (*(volatile int *)&ctrl->stop) {
                        stopped = true;
                        break;
                    }
                }
                if (stopped || *(volatile int *)&ctrl->stop) break;
                mbar_expect_tx(&ctrl->full[slot], (unsigned)P::STAGE_BYTES);
                bulk_g2s(ring + (size_t)slot * P::STAGE_BYTES, g_frag + (size_t)idx * P::STAGE_BYTES, (unsigned)P::STAGE_BYTES,
                         &ctrl->full[slot]);
                ++issued;
                idx += PRODUCERS;
                if (idx >= P::STAGES_PER_ITER) idx -= P::STAGES_PER_ITER;
                slot += PRODUCERS;
                if (slot >= P::NSTAGE) {
                    slot -= P::NSTAGE;
                    par ^= 1u;
                }
            }
            // drain: the stages issued beyond what the consumers took must land before the CTA may exit
            const int consumed = *(volatile int *)&ctrl->final_stages;
            int q = consumed > pid ? (consumed - pid + PRODUCERS - 1) / PRODUCERS : 0;
            for (; q < issued; ++q) {
                const int g = pid + PRODUCERS * q;
                mbar_wait(&ctrl->full[g % P::NSTAGE], (unsigned)((g / P::NSTAGE) & 1));
            }
        }
        return;
    }

    // =============================================== consumer warps ==============================================
    const int grp = warp / TEAM, rank = warp % TEAM;
    const int g = lane >> 2, t4 = lane & 3;
    const unsigned gmask = 0xFu << (4 * g);
    const bool leader = t4 == 0;
    double2 *st = reinterpret_cast<double2 *>(groups + (size_t)grp * P::GROUP_BYTES) + lane;
    double2 *win = st + E::NSTATE * 32;                 // input vector of the product
    constexpr int CONS_THREADS = P::CONSUMER_WARPS * 32;
    auto team_sync = [&]() {
        if (TEAM > 1) named_barrier(2 + grp, TEAM * 32);
        else __syncwarp();
    };

    const WorkQueue wq{io.queue, io.B, io.ready};
    wq.mark_start();
    unsigned long long stat_k = 0;
    unsigned int stat_nc = 0;
    long long inst = -1;
    int k = 0;
    bool live = false, drained = false;
    typename E::Lane L;
    E::lane_reset(L);
    for (int t = E::NW + E::NC; t < P::NIN_PAD; ++t) win[t * 32] = make_double2(0.0, 0.0);   // padding tiles meet zero fragments

    int iter = 0;
    for (;; ++iter) {
        // ---- refill (rank 0 pulls, the team shares the index)
        const bool need = !live && !drained;
        if (__any_sync(FULL, need)) {
            long long slot = -1;
            if (rank == 0) {
                if (need && leader) slot = wq.next();
                if (TEAM > 1 && leader) ctrl->slot[grp][g] = need ? slot : -2;
            }
            if (TEAM > 1) {
                team_sync();
                slot = ctrl->slot[grp][g];
                team_sync();
            } else {
                slot = __shfl_sync(FULL, slot, lane & ~3);
            }
            if (need) {
                if (slot < 0) {
                    drained = true;
                    if (leader && rank == 0) wq.mark_drained();
                } else {
                    inst = slot;
                    E::init(L, C, S, io, inst, st, win + E::NW * 32, t4, rank);
                    k = -Warmup<E>::value;
                    live = true;
                }
            }
            if (TEAM > 1) team_sync();      // the ranks of a team initialise different tiles
        }
        // ---- input vector w (tiles split over the team)
        if constexpr (HasPass<typename E::Lane>::value) L.pass = k + Warmup<E>::value;
#pragma unroll 1
        for (int t = rank; t < E::NW; t += TEAM) win[t * 32] = E::make_w(L, C, S, t, st, t4);
        team_sync();

        // ---- y = F [w ; c], NB output tiles per pass, fused with the update of those tiles
        bool over = false;
        {
            const int gs0 = iter * P::STAGES_PER_ITER;       // global index of this iteration's first stage
            double2 fc[KI][NB], vc[KI];
            auto wait_load = [&](int idx, int ch, double2 (&f)[KI][NB], double2 (&v)[KI]) {   // ch: the chunk of its pass (input tiles ch KI ..)
                const int gsi = gs0 + idx, slot = P::RESIDENT ? idx : gsi % P::NSTAGE;
                mbar_wait(&ctrl->full[slot], P::RESIDENT ? 0u : (unsigned)((gsi / P::NSTAGE) & 1));
                const double2 *sp = reinterpret_cast<const double2 *>(ring + (size_t)slot * P::STAGE_BYTES) + lane;
                const int it0 = ch * KI;
#pragma unroll
                for (int ki = 0; ki < KI; ++ki) {
                    v[ki] = win[(it0 + ki) * 32];
#pragma unroll
                    for (int b = 0; b < NB; ++b) f[ki][b] = sp[(ki * NB + b) * 32];
                }
            };
            auto release = [&](int idx) {
                if (!P::RESIDENT) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ctrl->empty[(gs0 + idx) % P::NSTAGE]);
                }
            };
            // this warp's stages: rank, rank + TEAM, rank + 2 TEAM, ...; NCH consecutive ones make a pass
            if (!__any_sync(FULL, live)) {
                // nothing left in this warp: keep the stream in lock step (wait for a stage, hand it back), no loads, no MMAs
#pragma unroll 1
                for (int idx = rank; idx < P::STAGES_PER_ITER; idx += TEAM) {
                    const int gsi = gs0 + idx, slot = P::RESIDENT ? idx : gsi % P::NSTAGE;
                    mbar_wait(&ctrl->full[slot], P::RESIDENT ? 0u : (unsigned)((gsi / P::NSTAGE) & 1));
                    release(idx);
                }
            } else {
                // Two operand buffers, stages taken in pairs: the operands of the next stage are read from shared memory while the
                // MMAs of the current one issue, and the hand-over is a renaming (no register moves, no predicates).  The chunk
                // loop is fully unrolled, so the input tiles are read at constant offsets.
                double2 fb[KI][NB], vb[KI];
                auto mmas = [&](double (&acc)[NB][2], const double2 (&f)[KI][NB], const double2 (&v)[KI]) {
#pragma unroll
                    for (int ki = 0; ki < KI; ++ki) {
#pragma unroll
                        for (int b = 0; b < NB; ++b) dmma(acc[b][0], acc[b][1], v[ki].x, f[ki][b].x, acc[b][0], acc[b][1]);
#pragma unroll
                        for (int b = 0; b < NB; ++b) dmma(acc[b][0], acc[b][1], v[ki].y, f[ki][b].y, acc[b][0], acc[b][1]);
                    }
                };
                wait_load(rank, 0, fc, vc);
#pragma unroll 1
                for (int round = 0; round < P::NROUND; ++round) {
                    double acc[NB][2];
#pragma unroll
                    for (int b = 0; b < NB; ++b) acc[b][0] = acc[b][1] = 0.0;
                    const int base = round * P::NCH * TEAM + rank;
#pragma unroll
                    for (int ch = 0; ch < P::NCH; ch += 2) {
                        const int idx = base + ch * TEAM;
                        wait_load(idx + TEAM, ch + 1, fb, vb);                  // stage ch + 1 of this pass
                        mmas(acc, fc, vc);
                        release(idx);
                        if (ch + 2 < P::NCH || round + 1 < P::NROUND) wait_load(idx + 2 * TEAM, (ch + 2) % P::NCH, fc, vc);   // the stage after it
                        mmas(acc, fb, vb);
                        release(idx + TEAM);
                    }
                    const int pass = ((round + r0) % P::NROUND) * TEAM + rank;
                    if (pass < P::NPASS) E::update(L, C, S, pass * NB, acc, st, t4, over);
                }
            }
        }

        // ---- exit condition
        unsigned ob = __ballot_sync(FULL, over);
        if (TEAM > 1) {
            if (lane == 0) ctrl->over[grp][rank] = ob;
            team_sync();
            ob = 0;
#pragma unroll
            for (int r = 0; r < TEAM; ++r) ob |= ctrl->over[grp][r];
        }
        if (live) k += 1;
        const bool gover = (ob & gmask) != 0u;
        if (live) {
            const int ef = (k >= 1 && !gover) ? 1 : ((k >= k_max) ? -1 : 0);      // (k < 1: a warm-up pass, no exit test)
            if (ef != 0) {
                if (rank == 0) {
                    E::finish(L, C, io, inst, st, t4);
                    if (leader) {
                        io.k[inst] = k;
                        io.e[inst] = ef;
                        stat_k += (unsigned long long)k;
                        stat_nc += (ef < 0);
                    }
                }
                live = false;
            }
        }
        // ---- CTA-wide: does any warp go on?  (the fragment stream is consumed in lock step)
        const bool warp_alive = __any_sync(FULL, live || !drained);
        if (lane == 0) {
            if (warp_alive) ctrl->alive[iter % 3] = 1;
            if (warp == 0) ctrl->alive[(iter + 1) % 3] = 0;
        }
        named_barrier(1, CONS_THREADS);
        if (*(volatile int *)&ctrl->alive[iter % 3] == 0) break;
    }
    if (threadIdx.x == 0) {
        ctrl->final_stages = (iter + 1) * P::STAGES_PER_ITER;
        __threadfence_block();
        ctrl->stop = 1;
    }
    if (rank == 0) flush_stats(io.queue, stat_k, stat_nc);
    wq.mark_end();
}

// ---- host-side traits: the scalar skeleton of `Solver` plus the engine E --------------------------------------------------------
template <class Solver, class E> struct DenseTraits : PolicyTraits<Solver> {
    typedef PolicyTraits<Solver> Base;
    typedef Plan<E> P;
    static constexpr bool HAS_MMA = P::HAS;
    static size_t blob_bytes() { return HAS_MMA ? P::BLOB_BYTES : Base::blob_bytes(); }
    static void fill_blob(void *dst) {
        memset(dst, 0, blob_bytes());
        Base::fill_blob(dst);
        if constexpr (HAS_MMA) {
            typename E::Small *S = new typename E::Small;
            memset(S, 0, sizeof *S);
            long double *F = new long double[(size_t)E::NO * 8 * P::NIN * 8]();
            E::fill(spcies_h_consts, *S, F);
            fill_fragments<E>(F, reinterpret_cast<double2 *>((char *)dst + P::OFF_FRAG));
            memcpy((char *)dst + P::OFF_SMALL, S, sizeof *S);
            delete[] F;
            delete S;
        }
    }
    static bool use_mma(int arith, const BatchIO &io) {
        if constexpr (!HAS_MMA) return false;
        return arith != SPCIES_CUDA_ARITH_EXACT && io.sol == nullptr && io.engine != SPCIES_CUDA_ENGINE_SCALAR && io.LB == nullptr;
    }
    static bool uses_scratch(int arith, const BatchIO &io) { return !use_mma(arith, io); }
    static void engine_shape(int arith, const BatchIO &io, int &block, size_t &smem, int &ipb) {
        ipb = block;
        if (use_mma(arith, io)) {
            block = P::BLOCK;
            smem = P::SMEM;
            ipb = P::IPB;
        }
    }
    static cudaError_t launch(int arith, bool varb, int grid, int block, size_t smem, cudaStream_t s, const BatchIO &io,
                              const void *dc, void *scratch) {
        if (io.engine == SPCIES_CUDA_ENGINE_MMA && !use_mma(arith, io)) return cudaErrorNotSupported;
        if constexpr (HAS_MMA) {
            if (use_mma(arith, io)) {
                cudaError_t e = cudaFuncSetAttribute(dense_mma_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM);
                if (e != cudaSuccess) return e;
                dense_mma_kernel<E><<<grid, P::BLOCK, P::SMEM, s>>>(io, (const unsigned char *)dc);
                return cudaGetLastError();
            }
        }
        return Base::launch(arith, varb, grid, block, smem, s, io, dc, scratch);
    }
    static cudaError_t attributes(int arith, bool varb, cudaFuncAttributes *a) {
        if constexpr (HAS_MMA) {
            if (arith != SPCIES_CUDA_ARITH_EXACT) return cudaFuncGetAttributes(a, dense_mma_kernel<E>);
        }
        return Base::attributes(arith, varb, a);
    }
};

}  // namespace dense
}  // namespace spcies
