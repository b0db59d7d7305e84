// puckswarm_b200 -- the engine: CUDA kernels (sm_100a) and the C-ABI of include/puckswarm_b200.h.
//
// One CTA steps one arena at a time (persistent grid: CTA i takes arenas i, i + grid, ...). Per-arena state
// lives in HBM in the ps_state_view layout; a CTA stages the bodies into shared memory, runs World.step there
// (ps_physics.cuh) and writes them back. Sensing runs one CTA per robot (ps_sense.cuh). There is no CPU path:
// without a CUDA device every entry point fails with PS_E_NODEVICE / PS_E_CUDA.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <map>
#include <mutex>
#include <dlfcn.h>
#include <climits>
#include "ps_host.hpp"
#include "ps_sense_dev.cuh"
#include "ps_pipeline.cuh"
#include "ps_island_lanes.cuh"

using namespace ps;

static thread_local std::string g_lastError;
static int setError(int code, const std::string& msg) {
    g_lastError = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                             \
    do {                                                                                                           \
        cudaError_t err__ = (expr);                                                                                \
        if (err__ != cudaSuccess) return setError(PS_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
    } while (0)

struct KParams {
    PhysEnv env;
    StatePtrs sp;
    char* slow;              // per-slot working memory in HBM
    size_t slowStride;
    size_t fastCap;          // dynamic shared memory handed to the carver
    int nArenas;
    int a0;                  // first arena of the launch (arena groups run on their own streams)
    int stepInterval;
    const int64_t* seeds;
    const ps_robot_state* robotInit;
    int nInformed;
    int puckShiftStep, puckShuffleStep;   // Arena.puckShiftStep / puckShuffleStep (-1 = never)
    int* stepStats;          // [nArenas][8] diagnostics of the last World.step
    // three-stage pipeline (ps_pipeline.cuh): per-arena scratch blocks in HBM + the batch-wide island queue
    PhysWork workOff;        // byte offsets of the working arrays inside one arena's scratch block
    char* arenaScratch;
    size_t arenaScratchStride;
    IslandQueue queue;
    unsigned long long* islandHist;   // [8][8] diagnostics: islands by (position iterations bucket, contact count bucket)
    const int* perm;         // arena order (heavy / light arena split) or null = identity: slot a0 + i of a launch is arena perm[a0 + i]
};
__device__ __forceinline__ int arenaAt(const KParams& p, int i) { return p.perm ? p.perm[p.a0 + i] : p.a0 + i; }

extern __shared__ __align__(16) char g_smem[];

__device__ __forceinline__ PhysWork ctaWork(const KParams& p) {
    Carver cv;
    cv.fast = g_smem;
    cv.fastCap = p.fastCap;
    cv.fastUsed = 0;
    cv.slow = p.slow + (size_t)blockIdx.x * p.slowStride;
    cv.slowUsed = 0;
    return carveWork(p.env.cfg, cv);
}

// Seeded construction of every arena (ps_spawn.cuh)
__global__ void __launch_bounds__(256) k_reset(KParams p) {
    __shared__ int scan[40];
    CtaCtx ctx;
    ctx.scanScratch = scan;
    PhysWork wk = ctaWork(p);
    for (int a = blockIdx.x; a < p.nArenas; a += gridDim.x) {
        ArenaState st = arenaState(p.sp, p.env.cfg, a);
        ResetOut out;
        out.robot = p.sp.robot + (size_t)a * p.env.cfg.d.R;
        out.arena = p.sp.arena + a;
        arenaReset(ctx, p.env, wk, st, out, p.seeds[a], p.robotInit, p.nInformed);
        __syncthreads();
    }
}

// The head of Arena.step (puck shift / shuffle perturbations, ps_spawn.cuh) for arenas [a0, a0 + n): launched before every
// think step when either perturbation is configured; an arena whose stepCount is not due returns at once
__global__ void __launch_bounds__(256) k_perturb(KParams p, int a0, int n, int perArenaScratch) {
    __shared__ int scan[40];
    CtaCtx ctx;
    ctx.scanScratch = scan;
    // arena groups run this concurrently on their own streams, so the working arrays come from the arenas' own scratch
    // blocks (HBM / L2: the perturbations are rare); the fused single-stream mode has CTA slots only
    PhysWork slotWork = ctaWork(p);
    for (int slot = a0 + blockIdx.x; slot < a0 + n; slot += gridDim.x) {
        int a = p.perm ? p.perm[slot] : slot;
        ArenaState st = arenaState(p.sp, p.env.cfg, a);
        PhysWork wk = perArenaScratch ? rebaseWork(p.workOff, p.arenaScratch + (size_t)a * p.arenaScratchStride) : slotWork;
        arenaPerturb(ctx, p.env, wk, st, p.sp.arena + a, p.puckShiftStep, p.puckShuffleStep);
        __syncthreads();
    }
}

// RunOffline.step minus sensing: Robot.move with the current commands + World.step, nTicks times per launch
__global__ void __launch_bounds__(256) k_physics(KParams p, int nTicks) {
    __shared__ int scan[40];
    CtaCtx ctx;
    ctx.scanScratch = scan;
    PhysWork wk = ctaWork(p);
    int R = p.env.cfg.d.R;
    for (int a = blockIdx.x; a < p.nArenas; a += gridDim.x) {
        ArenaState st = arenaState(p.sp, p.env.cfg, a);
        const ps_robot_state* rs = p.sp.robot + (size_t)a * R;
        // the commands live in fx/fy scratch of the force arrays until physLoad consumes them
        for (int t = 0; t < nTicks; ++t) {
            __syncthreads();
            for (int r = threadIdx.x; r < R; r += blockDim.x) {
                wk.fx[r] = rs[r].cmd_forwards;
                wk.fy[r] = rs[r].cmd_torque;
            }
            __syncthreads();
            worldStep(ctx, p.env, wk, st, wk.fx, wk.fy);
            if (threadIdx.x == 0) p.sp.arena[a].update_count += 1;
        }
        if (threadIdx.x < 8) p.stepStats[a * 8 + threadIdx.x] = wk.stats[threadIdx.x];
        __syncthreads();
    }
}

__device__ __forceinline__ PhysWork arenaWork(const KParams& p, int a) { return rebaseWork(p.workOff, p.arenaScratch + (size_t)a * p.arenaScratchStride); }
// The same into one shared-memory copy per CTA ; the caller synchronises. As per-thread values the ~60 pointers spill to local memory, which is HBM traffic.
__device__ __forceinline__ void arenaWorkShared(const KParams& p, int a, PhysWork& wk, ArenaState& st) {
    // PhysWork is ~60 pointers: one per thread (offset from the kernel parameters + the arena's scratch block)
    constexpr int kPtrs = (int)(sizeof(PhysWork) / sizeof(void*));
    static_assert(sizeof(PhysWork) == kPtrs * sizeof(void*), "PhysWork holds pointers only");
    const size_t base = (size_t)(p.arenaScratch + (size_t)a * p.arenaScratchStride);
    constexpr int kFatNow = (int)(offsetof(PhysWork, fatNow) / sizeof(void*));   // not part of the scratch block: null = st.fat
    for (int i = threadIdx.x; i < kPtrs; i += blockDim.x) reinterpret_cast<size_t*>(&wk)[i] = i == kFatNow ? 0 : reinterpret_cast<const size_t*>(&p.workOff)[i] + base;
    if (threadIdx.x == 32 % blockDim.x) {
        const PhysCfg& c = p.env.cfg;
        size_t A = (size_t)a;
        st.body = p.sp.body + A * 6 * c.d.NB;
        st.senseAabb = p.sp.senseAabb + A * 4 * c.d.NB;
        st.fat = p.sp.fat + A * 4 * c.d.NP;
        st.ckey = p.sp.ckey + A * c.MC;
        st.cinfo = p.sp.cinfo + A * c.MC;
        st.cimp = p.sp.cimp + A * c.MC * 4;
        st.nContacts = &p.sp.arena[a].n_contacts;
        st.invDt0 = &p.sp.arena[a].inv_dt0;
        st.errorFlags = &p.sp.arena[a].error_flags;
    }
}

// Shared memory of k_phys_pre beyond its static part: the contact graph and the DFS bookkeeping, sized for the configuration's
// bodies and for up to `capT` touching contacts (an arena with more keeps those arrays in its scratch block for that tick).
struct PreSmem {
    int nb;        // bodies the shared arrays hold (0: none -- the arena is too large)
    int capT;      // touching contacts the shared arrays hold
    __host__ __device__ static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
    __host__ __device__ size_t bytes() const {
        if (nb <= 0) return 0;
        return align16((size_t)(nb + 1) * 4) + align16((size_t)nb * 4) + align16((size_t)nb * 2) + align16((size_t)nb) + align16((size_t)capT * 4) + align16((size_t)capT * 4) +
               align16((size_t)capT);
    }
};
struct PreHook {
    PhysWork* wk;
    char* base;
    int capT;
    __device__ void operator()() const {
        // (physCollide ended with a barrier; counters[0] = touching contacts of this tick)
        if (base && threadIdx.x == 0 && wk->counters[0] <= capT) {
            char* q = base;
            wk->tcPair = (uint32_t*)q;
            q += PreSmem::align16((size_t)capT * 4);
            wk->adj = (uint16_t*)q;
            q += PreSmem::align16((size_t)capT * 4);
            wk->tcFlag = (uint8_t*)q;
        }
        __syncthreads();
    }
};

// stage A: one CTA per arena
__global__ void __launch_bounds__(128) k_phys_pre(KParams p, PreSmem sm) {
    extern __shared__ __align__(16) char dsm[];
    __shared__ int scan[40];
    CtaCtx ctx;
    ctx.scanScratch = scan;
    int a = arenaAt(p, blockIdx.x);
    int R = p.env.cfg.d.R;
    __shared__ PhysWork wk;     // one copy per CTA (see k_phys_post)
    __shared__ ArenaState st;
    arenaWorkShared(p, a, wk, st);
    __syncthreads();
    // the contact graph and the DFS bookkeeping are used by this stage only and are walked with dependent loads: they live in
    // shared memory when the arena is small enough (the usual case)
    PreHook hook = {&wk, nullptr, sm.capT};
    if (sm.nb > 0) {
        char* q = dsm;
        int* adjOff = (int*)q;
        q += PreSmem::align16((size_t)(sm.nb + 1) * 4);
        int* adjFill = (int*)q;
        q += PreSmem::align16((size_t)sm.nb * 4);
        uint16_t* stack = (uint16_t*)q;
        q += PreSmem::align16((size_t)sm.nb * 2);
        uint8_t* bodyFlag = (uint8_t*)q;
        q += PreSmem::align16((size_t)sm.nb);
        hook.base = q;
        if (threadIdx.x == 0) {
            wk.adjOff = adjOff;
            wk.adjFill = adjFill;
            wk.stack = stack;
            wk.bodyFlag = bodyFlag;
        }
    }
    __syncthreads();
    const ps_robot_state* rs = p.sp.robot + (size_t)a * R;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        wk.fx[r] = rs[r].cmd_forwards;
        wk.fy[r] = rs[r].cmd_torque;
    }
    __syncthreads();
    long long pc[4] = {0, 0, 0, 0};
    physPre(ctx, p.env, wk, st, wk.fx, wk.fy, p.queue, a, pc, hook);
#if defined(PS_PHASE_CLOCKS)
    if (threadIdx.x == 0) {   // census [11..14]: cycles of the stage's phases, [15]: arena-ticks
        for (int i = 0; i < 4; ++i) atomicAdd(&p.islandHist[64 + 11 + i], (unsigned long long)pc[i]);
        atomicAdd(&p.islandHist[64 + 15], 1ULL);
    }
#endif
}

// entry `idx` of the warp-bucket islands in hand-out order: buckets kBucketXXL, kBucketXL, kBucketLarge, then bucket 0
__device__ __forceinline__ unsigned warpBucketEntry(const IslandQueue& q, int idx, int n3, int n2, int nLarge) {
    size_t cap = (size_t)q.cap;
    if (idx < n3) return q.entries[kBucketXXL * cap + idx];
    if (idx < n2) return q.entries[kBucketXL * cap + (idx - n3)];
    if (idx < nLarge) return q.entries[kBucketLarge * cap + (idx - n2)];
    return q.entries[idx - nLarge];
}

// stage B1: islands too large for a lane slot (queue bucket 0): one warp per island, staged in shared memory.
// VEL_ONLY: the velocity half only (k_island_vel); the position sweeps follow in k_island_pos on the same stream.
// mode 0: every warp-bucket island; 1: the heavy ones only (buckets kBucketXXL, kBucketXL); 2: the others (kBucketLarge, bucket 0)
template <bool VEL_ONLY>
__device__ __forceinline__ void islandWarpKernel(const KParams& p, int mode) {
    extern __shared__ __align__(16) char smem[];
    WarpSlot* slot = reinterpret_cast<WarpSlot*>(smem) + (threadIdx.x >> 5);
    // the heaviest islands are handed out first (buckets kBucketXXL, kBucketXL, kBucketLarge), then bucket 0
    int n3 = min(p.queue.counts[kBucketXXL], p.queue.cap), n2 = n3 + min(p.queue.counts[kBucketXL], p.queue.cap);
    int nLarge = n2 + min(p.queue.counts[kBucketLarge], p.queue.cap);
    int total = nLarge + min(p.queue.counts[0], p.queue.cap);
    unsigned NB = (unsigned)p.env.cfg.d.NB;
    int lane = threadIdx.x & 31;
    int first = mode == 2 ? n2 : 0, last = mode == 1 ? n2 : total;
    int* head = &p.queue.counts[VEL_ONLY ? kQHeadVel : kQHeadWarp];
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = first + atomicAdd(head, 1);
        idx = __shfl_sync(0xFFFFFFFFu, idx, 0);
        if (idx >= last) break;
        unsigned entry = warpBucketEntry(p.queue, idx, n3, n2, nLarge);
        int a = (int)(entry / NB), k = (int)(entry % NB);
        PhysWork wk = arenaWork(p, a);
        ArenaState st = arenaState(p.sp, p.env.cfg, a);
        int iters = 0;
        long long tw0 = clock64();
        bool ok = VEL_ONLY ? islandVelocityWarp(p.env, wk, st, k, *slot, &iters) : islandSolveWarp(p.env, wk, st, k, *slot, &iters);
        if (!ok && lane == 0) {
            atomicAdd(&p.islandHist[62], 1ULL);   // diagnostic: islands that did not fit a slot
            // too large for a slot: lane 0 runs the island straight from HBM / L2
            IslandCursor cur;
            islandBegin(cur, wk, k);
            while (islandStep(p.env, wk, st, cur, &iters)) {
            }
        }
        if (lane == 0) {
            atomicMax(&wk.stats[0], iters);
            // diagnostic: the slowest oversized island of the run: kilo-cycles << 32 | contacts << 20 | bodies << 10 | sweeps
            unsigned long long kc = (unsigned long long)((clock64() - tw0) >> 10);
            unsigned long long nc = (unsigned long long)(wk.islandStart[k + 1] - wk.islandStart[k]);
            unsigned long long nb = (unsigned long long)(wk.islandBodyStart[k + 1] - wk.islandBodyStart[k]);
            atomicMax(&p.islandHist[63], (kc << 32) | ((nc & 0xFFF) << 20) | ((nb & 0x3FF) << 10) | (unsigned long long)(iters & 0x3FF) | (ok ? 0ULL : (1ULL << 31)));
            // census: warp-slot islands by duration (bins of 2^18 cycles), their total kilo-cycles and count
            atomicAdd(&p.islandHist[64 + 16 + (int)min(15ULL, kc >> 8)], 1ULL);
            atomicAdd(&p.islandHist[64 + 48], kc);
            atomicAdd(&p.islandHist[64 + 49], 1ULL);
            if (ok) {
                atomicAdd(&p.islandHist[64 + 70], (unsigned long long)slot->nRows * (unsigned long long)iters);
                atomicAdd(&p.islandHist[64 + 71], nc * (unsigned long long)iters);
                atomicAdd(&p.islandHist[64 + 72], nc);
                atomicAdd(&p.islandHist[64 + 73], nb);
                atomicAdd(&p.islandHist[64 + 74], (unsigned long long)iters);
            }
        }
        __syncwarp();
    }
}
__global__ void __launch_bounds__(128) k_island_warp(KParams p, int mode) { islandWarpKernel<false>(p, mode); }
__global__ void __launch_bounds__(128) k_island_vel(KParams p, int mode) { islandWarpKernel<true>(p, mode); }

// stage B1b: the position sweeps of the islands k_island_vel prepared: G lanes per island, 32 / G islands per warp (one warp
// per CTA), every group on its own row schedule, finished groups written back and refilled by the whole warp.
template <int G>
__global__ void __launch_bounds__(32) k_island_pos(KParams p, int mode) {
    extern __shared__ __align__(16) char smem[];
    constexpr int NG = 32 / G;
    constexpr unsigned GMASK = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u);
    const unsigned FULL = 0xFFFFFFFFu;
    PosSlot* slots = reinterpret_cast<PosSlot*>(smem);
    int lane = threadIdx.x & 31, grp = lane / G, sub = lane % G;
    PosSlot& S = slots[grp];
    SlotCtx<PosSlot> L;
    L.S = &S;
    int posIters = p.env.cfg.posIters;
    if (posIters <= 0) return;
    int n3 = min(p.queue.counts[kBucketXXL], p.queue.cap), n2 = n3 + min(p.queue.counts[kBucketXL], p.queue.cap);
    int nLarge = n2 + min(p.queue.counts[kBucketLarge], p.queue.cap);
    int total = nLarge + min(p.queue.counts[0], p.queue.cap);
    unsigned NB = (unsigned)p.env.cfg.d.NB;
    int first = mode == 2 ? n2 : 0;
    bool running = false, exhausted = false;
    int myEntry = -1;                  // arena * NB + island of this group's island
    int nRows = 0, r = 0, base = 0, rowEnd = 0, iters = 0;
    float minSep = 0.0f;
    bool changed = false;
    for (;;) {
        // ---- finish + refill, by the whole warp, for every group that has no island
        unsigned need = __ballot_sync(FULL, !running && !exhausted);
        while (need) {
            int g = (__ffs(need) - 1) / G;
            need &= ~(GMASK << (g * G));
            int ent = __shfl_sync(FULL, myEntry, g * G);
            int its = __shfl_sync(FULL, iters, g * G);
            PosSlot& T = slots[g];
            if (ent >= 0) {
                int a = (int)((unsigned)ent / NB), k = (int)((unsigned)ent % NB);
                PhysWork wk = arenaWork(p, a);
                posFinish(p.env, wk, k, T, lane);
                if (lane == 0) atomicMax(&wk.stats[0], its);
            }
            bool mine = grp == g;
            if (mine) myEntry = -1;
            bool staged = false;
            unsigned entry = 0;
            while (!staged) {   // islands that outgrew the slot were solved by k_island_vel's fallback: skip them
                int idx = 0;
                if (lane == 0) idx = first + atomicAdd(&p.queue.counts[kQHeadPos], 1);
                idx = __shfl_sync(FULL, idx, 0);
                if (idx >= total) break;
                entry = warpBucketEntry(p.queue, idx, n3, n2, nLarge);
                int a = (int)(entry / NB), k = (int)(entry % NB);
                PhysWork wk = arenaWork(p, a);
                __syncwarp();
                staged = posStage(p.env, wk, k, T, lane);
            }
            __syncwarp();
            if (mine) {
                if (staged) {
                    myEntry = (int)entry;
                    running = true;
                    iters = 0;
                    nRows = T.nRows;
                    r = 0;
                    base = 0;
                    rowEnd = T.rowStart[1];
                    minSep = 0.0f;
                    changed = false;
                } else {
                    exhausted = true;
                }
            }
        }
        if (__all_sync(FULL, !running)) break;
        // ---- one pass of every running group's current row: a lane pair per constraint, G / 2 constraints per pass
        {
            int q = base + (sub >> 1);
            bool on = running && q < rowEnd;
            int ci = on ? q : 0;
            uint32_t hd = on ? S.pcHead[ci] : 0u;
            int bA = (int)(hd & 255u), bB = (int)((hd >> 8) & 255u), type = (int)((hd >> 16) & 255u), np = (int)(hd >> 24);
            V2 ln = mk(S.pc[ci][0], S.pc[ci][1]), lp = mk(S.pc[ci][2], S.pc[ci][3]);
            float radius = S.pc[ci][8];
            float sep0 = lkCorrectPointPair(p.env, L, on && np > 0, sub & 1, bA, bB, type, ln, lp, mk(S.pc[ci][4], S.pc[ci][5]), radius, &changed);
            minSep = fminf_(minSep, sep0);
            if (__any_sync(FULL, np > 1)) {
                float sep1 = lkCorrectPointPair(p.env, L, on && np > 1, sub & 1, bA, bB, type, ln, lp, mk(S.pc[ci][6], S.pc[ci][7]), radius, &changed);
                minSep = fminf_(minSep, sep1);
            }
        }
        // ---- advance the cursors; at the end of a sweep the group decides like ContactSolver.solvePositionConstraints' caller
        float m = minSep;
        for (int o = 1; o < G; o <<= 1) m = fminf_(m, __shfl_xor_sync(FULL, m, o));
        unsigned chg = __ballot_sync(FULL, changed);
        if (running) {
            base += G / 2;
            if (base >= rowEnd) {
                ++r;
                if (r < nRows) {
                    base = rowEnd;
                    rowEnd = S.rowStart[r + 1];
                } else {
                    ++iters;
                    bool anyChanged = ((chg >> (grp * G)) & GMASK) != 0u;
                    if (m >= -1.5f * kLinearSlop || iters >= posIters) running = false;
                    else if (!anyChanged) {   // fixed point: every later sweep repeats this one
                        iters = posIters;
                        running = false;
                    }
                    r = 0;
                    base = 0;
                    rowEnd = S.rowStart[1];
                    minSep = 0.0f;
                    changed = false;
                }
            }
        }
    }
}

// stage B0: the islands that outgrow a WarpSlot (queue bucket kBucketBig): one warp (= one CTA) per island
__global__ void __launch_bounds__(32) k_island_big(KParams p) {
    extern __shared__ __align__(16) char smem[];
    BigSlot* slot = reinterpret_cast<BigSlot*>(smem);
    int cap = p.queue.cap;
    int total = min(p.queue.counts[kBucketBig], cap);
    unsigned NB = (unsigned)p.env.cfg.d.NB;
    int lane = threadIdx.x & 31;
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
        unsigned entry = p.queue.entries[(size_t)kBucketBig * cap + idx];
        int a = (int)(entry / NB), k = (int)(entry % NB);
        PhysWork wk = arenaWork(p, a);
        ArenaState st = arenaState(p.sp, p.env.cfg, a);
        int iters = 0;
        long long tw0 = clock64();
        bool ok = islandSolveBig(p.env, wk, st, k, *slot, &iters);
        if (!ok && lane == 0) {
            atomicAdd(&p.islandHist[62], 1ULL);   // diagnostic: islands that did not fit any slot
            // lane 0 runs the island straight from HBM / L2
            IslandCursor cur;
            islandBegin(cur, wk, k);
            while (islandStep(p.env, wk, st, cur, &iters)) {
            }
        }
        if (lane == 0) {
            atomicMax(&wk.stats[0], iters);
            unsigned long long kc = (unsigned long long)((clock64() - tw0) >> 10);
            unsigned long long nc = (unsigned long long)(wk.islandStart[k + 1] - wk.islandStart[k]);
            unsigned long long nb = (unsigned long long)(wk.islandBodyStart[k + 1] - wk.islandBodyStart[k]);
            atomicMax(&p.islandHist[63], (kc << 32) | ((nc & 0xFFF) << 20) | ((nb & 0x3FF) << 10) | (unsigned long long)(iters & 0x3FF) | (ok ? 0ULL : (1ULL << 31)));
            atomicAdd(&p.islandHist[64 + 32 + (int)min(15ULL, kc >> 8)], 1ULL);   // census: big islands by duration
            if (ok) {   // ... and their row passes, contact-sweeps, contacts, bodies
                atomicAdd(&p.islandHist[64 + 64], (unsigned long long)slot->nRows * (unsigned long long)iters);
                atomicAdd(&p.islandHist[64 + 65], nc * (unsigned long long)iters);
                atomicAdd(&p.islandHist[64 + 66], nc);
                atomicAdd(&p.islandHist[64 + 67], nb);
                atomicAdd(&p.islandHist[64 + 68], (unsigned long long)iters);
            }
            atomicAdd(&p.islandHist[64 + 50], kc);
            atomicAdd(&p.islandHist[64 + 51], 1ULL);
        }
        __syncwarp();
    }
}

// Heavy / light arena split: the nHeavy arenas with the highest scores (their longest islands lately) go first in `perm`, the
// others follow, both in arena order (deterministic). One CTA; the scores decay so that an arena leaves the heavy group a few
// think steps after its island is gone.
__global__ void __launch_bounds__(1024) k_classify(int* score, int* perm, int nArenas, int nHeavy) {
    __shared__ int scan[40];
    __shared__ int hist[256];
    __shared__ int sh[3];   // threshold bin, ties taken into the heavy group, -
    CtaCtx ctx;
    ctx.scanScratch = scan;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int a = threadIdx.x; a < nArenas; a += blockDim.x) atomicAdd(&hist[min(255, score[a])], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int above = 0, t = 255;
        while (t > 0 && above + hist[t] <= nHeavy) above += hist[t--];
        // bins > t are heavy as a whole (above <= nHeavy); bin t fills the rest in arena order (bin 0: arenas without any scored island)
        sh[0] = t;
        sh[1] = nHeavy - above;
    }
    __syncthreads();
    int t = sh[0], tiesWanted = sh[1];
    int nH = 0, nTie = 0, nL = 0;
    for (int base = 0; base < nArenas; base += blockDim.x) {
        int a = base + threadIdx.x;
        int q = a < nArenas ? min(255, score[a]) : -1;
        bool hv = q > t, tie = q == t, any = q >= 0;
        int totH, totT, totA;
        int pH = ctx.exscanFlag(hv, totH);
        int pT = ctx.exscanFlag(tie, totT);
        int pA = ctx.exscanFlag(any, totA);
        if (any) {
            bool heavy = hv || (tie && nTie + pT < tiesWanted);
            // heavy rank: heavies before it; light rank: everything before it that is not heavy
            int tiesBefore = min(nTie + pT, tiesWanted);
            int heavyBefore = nH + pH + tiesBefore;
            int pos = heavy ? heavyBefore : nHeavy + (base + pA) - heavyBefore;
            perm[pos] = a;
            score[a] = (score[a] * 3) >> 2;
        }
        nH += totH;
        nTie += totT;
        nL += totA;
        __syncthreads();
    }
}

static constexpr int kBatchThreads = 64;   // at most two lane slots (2 x 32 islands) per CTA: a slot is ~41 KB of shared memory (run-time: ps_engine::batchThreads)
// stage B2: all other islands (queue bucket 1): one lane per island, 32 islands per warp, bodies in shared memory
// bucket < 0: one launch for all lane buckets, segBlocks CTAs each, the buckets of the largest islands first
template <class Slots>
__device__ __forceinline__ void islandBatchKernel(const KParams& p, int bucket, int segBlocks) {
    extern __shared__ __align__(16) char smem[];
    LaneCtxT<Slots> L;
    L.S = reinterpret_cast<Slots*>(smem) + (threadIdx.x >> 5);
    L.pmPuck = p.env.hot.pmPuck;
    L.piPuck = p.env.hot.piPuck;
    L.pmRobot = p.env.hot.pmRobot;
    L.piRobot = p.env.hot.piRobot;
    L.lane = threadIdx.x & 31;
    L.nb = L.lo = L.nc = 0;
    int cap = p.queue.cap;
    int blk = blockIdx.x, nBlk = gridDim.x;
    if (bucket < 0) {
        bucket = kLaneBuckets - blk / segBlocks;
        blk = blk % segBlocks;
        nBlk = segBlocks;
    }
    int total = min(p.queue.counts[bucket], cap);
    unsigned NB = (unsigned)p.env.cfg.d.NB;
    int stride = nBlk * blockDim.x;
    int rounds = (total + stride - 1) / stride;
    for (int rnd = 0; rnd < rounds; ++rnd) {
        int idx = rnd * stride + blk * blockDim.x + threadIdx.x;
        bool active = idx < total;
        PhysWork wk;
        ArenaState st;
        int k = 0;
        if (active) {
            unsigned entry = p.queue.entries[(size_t)bucket * cap + idx];
            int a = (int)(entry / NB);
            k = (int)(entry % NB);
            wk = arenaWork(p, a);
            st = arenaState(p.sp, p.env.cfg, a);
            lkStage(p.env, wk, k, L);
        }
        __syncwarp();
        long long tb0 = clock64();
        int iters = lkSolve(p.env, wk, st, L, active);
        if (L.lane == 0) {   // census: warp rounds of the lane batches per bucket, their kilo-cycles, and the lanes' summed sweeps vs the warp's
            int cb = bucket == kBucketLaneX ? 5 : bucket - 1;   // the wide lane slot's rounds go to [62], [63]
            atomicAdd(&p.islandHist[64 + 52 + 2 * cb], (unsigned long long)((clock64() - tb0) >> 10));
            atomicAdd(&p.islandHist[64 + 53 + 2 * cb], 1ULL);
        }
        {
            int sum = active ? iters : 0, mx = sum;
            for (int o = 16; o > 0; o >>= 1) {
                sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
                mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
            }
            if (L.lane == 0) {
                atomicAdd(&p.islandHist[64 + 60], (unsigned long long)sum);
                atomicAdd(&p.islandHist[64 + 61], (unsigned long long)(32 * mx));
            }
        }
        if (active) {
            lkFinish(p.env, wk, L);
            atomicMax(&wk.stats[0], iters);
            int nc = wk.islandStart[k + 1] - wk.islandStart[k];
            if (nc >= 3) {
                int ib = iters <= 1 ? 0 : iters <= 3 ? 1 : iters <= 6 ? 2 : iters <= 12 ? 3 : iters <= 25 ? 4 : iters <= 50 ? 5 : iters < 100 ? 6 : 7;
                int cb = nc <= 3 ? 0 : nc <= 5 ? 1 : nc <= 8 ? 2 : nc <= 12 ? 3 : nc <= 16 ? 4 : nc <= 32 ? 5 : nc <= 64 ? 6 : 7;
                atomicAdd(&p.islandHist[ib * 8 + cb], 1ULL);
            }
        }
        __syncwarp();
    }
}
__global__ void __launch_bounds__(kBatchThreads) k_island_batch(KParams p, int bucket, int segBlocks) { islandBatchKernel<LaneSlotsU>(p, bucket, segBlocks); }
// the wide lane slot (queue bucket kBucketLaneX): one warp per CTA
__global__ void __launch_bounds__(32) k_island_batch_x(KParams p, int segBlocks) { islandBatchKernel<LaneSlotsX>(p, kBucketLaneX, segBlocks); }

// stage C: one CTA per arena. Dynamic shared memory: nbFat bodies' box unions, then npFat proxies' fat boxes (0: not staged)
#ifndef PS_POST_MIN_CTAS
#define PS_POST_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(128, PS_POST_MIN_CTAS) k_phys_post(KParams p, int nbFat, int npFat) {
    extern __shared__ __align__(16) char dsm[];
    __shared__ int scan[40];
    CtaCtx ctx;
    ctx.scanScratch = scan;
    int a = arenaAt(p, blockIdx.x);
    // the ~60 array pointers of the arena live in shared memory, one copy per CTA: as per-thread values they spill to
    // local memory, and local memory goes through L2 to HBM
    __shared__ PhysWork wk;
    __shared__ ArenaState st;
    arenaWorkShared(p, a, wk, st);
    __syncthreads();
    // the pair search reads every body's box union once per moved body, and this step's fat boxes of every proxy: both tables
    // are kept in shared memory when they fit
    if (threadIdx.x == 0) {
        float* sBodyFat = (float*)dsm;
        float* sFatNow = sBodyFat + 4 * nbFat;
        if (nbFat > 0) wk.bodyFat = sBodyFat;
        if (npFat > 0) wk.fatNow = sFatNow;
    }
    __syncthreads();
    physPost(ctx, p.env, wk, st);
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.queue.load) {   // this tick's count of warp-class islands becomes "the previous tick's"
        p.queue.load[0] = p.queue.load[1];
        p.queue.load[1] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x <= kQueueBuckets)   // census: islands per queue bucket, ticks
        atomicAdd(&p.islandHist[64 + threadIdx.x], threadIdx.x < kQueueBuckets ? (unsigned long long)p.queue.counts[threadIdx.x] : 1ULL);
    if (threadIdx.x == 0) p.sp.arena[a].update_count += 1;
    if (threadIdx.x < 8) p.stepStats[a * 8 + threadIdx.x] = wk.stats[threadIdx.x];
}

// OR of every arena's error_flags (+ bit1 for a robot whose last local map overflowed) and the number of flagged arenas
__global__ void k_error_scan(const ps_arena_state* arena, const ps_localmap* lm, int nArenas, int R, int* out) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    int f = 0;
    if (a < nArenas) {
        f = arena[a].error_flags;
        if (lm)
            for (int r = 0; r < R; ++r)
                if (lm[(size_t)a * R + r].overflow) f |= ERR_LM_CAP;
    }
    unsigned any = __ballot_sync(0xFFFFFFFFu, f != 0);
    for (int o = 16; o > 0; o >>= 1) f |= __shfl_xor_sync(0xFFFFFFFFu, f, o);
    if ((threadIdx.x & 31) == 0 && any) {
        atomicOr(&out[0], f);
        atomicAdd(&out[1], __popc(any));
    }
}

// The MathUtils sin table (228 KB) is the same for every handle: one device copy per GPU, shared by the handles of the
// process, so that several engines on one GPU do not multiply its L1 / L2 footprint (it sits on the solver's critical path).
struct SharedLut {
    float* ptr = nullptr;
    int refs = 0;
};
static std::mutex g_lutMutex;
static std::map<int, SharedLut> g_lut;
static float* acquireSharedLut(int device, const float* host) {
    std::lock_guard<std::mutex> lock(g_lutMutex);
    SharedLut& L = g_lut[device];
    if (!L.ptr) {
        if (cudaMalloc(&L.ptr, kLutLength * sizeof(float)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(L.ptr, host, kLutLength * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(L.ptr);
            L.ptr = nullptr;
            return nullptr;
        }
    }
    L.refs++;
    return L.ptr;
}
static void releaseSharedLut(int device) {
    std::lock_guard<std::mutex> lock(g_lutMutex);
    auto it = g_lut.find(device);
    if (it == g_lut.end() || it->second.refs <= 0) return;
    if (--it->second.refs == 0) {
        cudaFree(it->second.ptr);
        it->second.ptr = nullptr;
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct ps_engine {
    HostSetup hs;
    SenseSetup sense;
    int device = 0;
    cudaStream_t stream = nullptr;
    // Arena groups: arenas are independent, so the batch is cut into groups that run their kernel sequences on separate
    // streams; the latency tail of one group's slowest island overlaps with the other groups' work.
    struct Group {
        cudaStream_t s = nullptr, s2 = nullptr;   // s2: the few oversized islands, next to the lane batches
        cudaStream_t sb[kQueueBuckets] = {};      // the lane batches of the other buckets run side by side as well
        cudaEvent_t evb[kQueueBuckets] = {};
        cudaEvent_t evPre = nullptr, evBig = nullptr, done = nullptr;
        int a0 = 0, n = 0;
        IslandQueue q = {};
    };
    std::vector<Group> groups;
    cudaEvent_t evStart = nullptr;
    bool mainDirty = true;     // the main stream has work the group streams have not been ordered after yet
    bool groupsDirty = false;  // the group streams have work the main stream has not been ordered after yet
    KParams kp;
    Scene* dScene = nullptr;
    float* dLut = nullptr;
    ps_robot_state* dRobotInit = nullptr;
    int64_t* dSeeds = nullptr;
    int* dStepStats = nullptr;
    int nSlots = 0, threads = 128;
    bool fused = false;      // PS_FUSED=1: the single-kernel World.step (kept for comparison)
    int workerBlocks = 0, batchBlocks = 0, batchXBlocks = 0, bigBlocks = 0, posBlocks = 0, batchThreads = kBatchThreads;
    int* dPerm = nullptr;      // heavy / light arena split (PS_HEAVY_ARENAS): arena order, the nHeavy heaviest first; null = off
    int* dScore = nullptr;     // [nArenas] heaviness scores (IslandQueue::score)
    int nHeavy = 0;
    bool islandSplit = false;   // PS_ISLAND_SPLIT=1 with PS_ISLAND_G: only the lighter warp-bucket islands go through the split solver
    int islandG = 0;         // PS_ISLAND_G = 2 / 4 / 8: split solver (k_island_vel + k_island_pos with G lanes per island); 0 (default): the one-kernel
                             // warp-per-island solver. The split solver issues 4-5x fewer instructions but every island's sweeps take longer (the groups
                             // of a warp share each other's divergent paths), and the tick waits for the longest island: measured slower (profiles/r02_summary.md)
    bool mergeBatch = true;    // one launch for all lane buckets (PS_MERGE_BATCH=0: one launch and stream per bucket)
    size_t smemBytes = 0, bigSmem = 0;
    PreSmem preSmem = {0, 0};   // k_phys_pre / k_phys_post: CTA width and what they keep in shared memory
    bool preThreadsSet = false;
    int preThreads = 64, postThreads = 128, postNbFat = 0, postNpFat = 0;
    int64_t launches = 0;
    int updateCount = 0;     // all arenas advance in lock step
    SenseDevice sd;
    StatsDev* dStatsDev = nullptr;
    ps_stats* dStats = nullptr;
    float* dCmd = nullptr;
    int* dErr = nullptr;     // [2]: OR of the arenas' error flags, flagged arenas (k_error_scan)
    SnapPtrs snap = {};      // the reference's periodic record + Analyze accumulators (ps_sense_dev.cuh)
    int stepCount = 0;       // Arena.stepCount of the batch (think steps done; all arenas in lock step)
    float analysisThreshold = 1.5f, analysisTarget = 50.0f;
    bool isReset = false;
    // optional per-kernel timing (CUDA events on the engine's stream): classes 0 sense, 1 think, 2 physics, 3 other
    bool profiling = false;
    struct Timed {
        cudaEvent_t a, b;
        int cls;
    };
    std::vector<Timed> timed;
    double kernelMs[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int64_t kernelCount[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

struct ScopedTimer {   // records an event pair around the launches issued during its lifetime
    ps_engine* e;
    ps_engine::Timed t;
    bool on;
    cudaStream_t s;
    ScopedTimer(ps_engine* e_, int cls, cudaStream_t s_ = nullptr) : e(e_), on(e_->profiling), s(s_ ? s_ : e_->stream) {
        if (!on) return;
        t.cls = cls;
        if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) {
            on = false;
            return;
        }
        cudaEventRecord(t.a, s);
    }
    void stop() {
        if (!on) return;
        cudaEventRecord(t.b, s);
        e->timed.push_back(t);
        on = false;
    }
    ~ScopedTimer() { stop(); }
};

extern "C" {
static int forkGroups(ps_handle h);
static int joinGroups(ps_handle h);
}

static size_t envSize(const char* name, size_t dflt) {
    const char* v = getenv(name);
    return v && *v ? (size_t)strtoull(v, nullptr, 10) : dflt;
}

extern "C" {

const char* ps_last_error(void) { return g_lastError.c_str(); }
int ps_abi_version(void) { return PS_ABI_VERSION; }
int ps_struct_size(int which) {
    switch (which) {
        case 0: return (int)sizeof(ps_config);
        case 1: return (int)sizeof(ps_robot_state);
        case 2: return (int)sizeof(ps_arena_state);
        case 3: return (int)sizeof(ps_localmap);
        case 4: return (int)sizeof(ps_stats);
        case 5: return (int)sizeof(ps_state_view);
        case 6: return (int)sizeof(ps_obs_view);
        case 7: return (int)sizeof(ps_snapshot_view);
        case 8: return (int)sizeof(ps_arena_analysis);
        default: return -1;
    }
}

void ps_config_defaults(ps_config* c) {
    std::memset(c, 0, sizeof(*c));
    c->abi_version = PS_ABI_VERSION;
    c->n_arenas = 1;
    c->n_robots = 4;
    c->n_pucks = 10;
    c->n_puck_types = 2;
    c->n_informed_robots = 0;
    c->controller = PS_CTRL_CACHECONS;
    c->rng_mode = PS_RNG_JAVA_SHARED;
    c->step_interval = 3;
    c->vel_iters = 3;
    c->pos_iters = 100;
    c->dt = 1.0f / 14.0f;
    c->enclosure_scale = 1.0f;
    c->sincos_lut = 1;
    c->toi_enabled = 1;
    c->max_contacts = 0;
    c->cluster_distance_threshold = 1.5f;
    c->carry_threshold_lo = 0.1f;
    c->carry_threshold_hi = 0.4f;
    c->vfh_safety_gamma = 2.0f;
    c->vfh_safety_mask = 5.0f;
    c->vfh_tau_lo = 0.7f;
    c->vfh_tau_hi = 0.85f;
    c->cc_k1 = 1.0f;
    c->cc_k2 = 8.0f;
    c->cc_pick_up_time = 20;
    c->cc_push_in_time = 1;
    c->cc_backup_time = 5;
    c->cc_prediction_distance_sqd = 100.0f;
    c->cc_st_dev = 0.5f;
    c->cc_exile_time = 20;
    c->cc_spit_out_time = 10;
    c->cc_cache_selection = PS_SEL_RELATIVE;
    c->cc_k_absolute = 40.0f;
    c->cc_k_relative = 8.0f;
    c->cc_forget_prob = 0.001f;
    c->cc_cache_separation = PS_SEP_ACCEPT_LARGER;
    c->cc_home_separation_distance = 50.0f;
    c->cc_ignore_pucks = 0;
    c->bhd_push_in_time = 1;
    c->bhd_backup_time = 5;
    c->ps_placement_time = 40;
    c->puck_shift_step = -1;
    c->puck_shuffle_step = -1;
}

int ps_destroy(ps_handle h) {
    if (!h) return PS_OK;
    cudaSetDevice(h->device);
    for (auto& g : h->groups) {
        if (g.s) cudaStreamSynchronize(g.s);
        if (g.s2) cudaStreamSynchronize(g.s2);
        for (int b = 0; b < kQueueBuckets; b++)
            if (g.sb[b]) cudaStreamSynchronize(g.sb[b]);
    }
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->kp.sp.body);
    cudaFree(h->kp.sp.senseAabb);
    cudaFree(h->kp.sp.fat);
    cudaFree(h->kp.sp.ckey);
    cudaFree(h->kp.sp.cinfo);
    cudaFree(h->kp.sp.cimp);
    cudaFree(h->kp.sp.robot);
    cudaFree(h->kp.sp.arena);
    cudaFree(h->kp.slow);
    cudaFree(h->kp.arenaScratch);
    cudaFree(h->kp.queue.entries);
    cudaFree(h->kp.queue.counts);
    cudaFree(h->kp.islandHist);
    cudaFree(h->dPerm);
    cudaFree(h->dScore);
    cudaFree(h->dScene);
    if (h->dLut) releaseSharedLut(h->device);
    cudaFree(h->dRobotInit);
    cudaFree(h->dSeeds);
    cudaFree(h->dStepStats);
    cudaFree(h->dStatsDev);
    cudaFree(h->dStats);
    cudaFree(h->dCmd);
    cudaFree(h->dErr);
    cudaFree(h->snap.step);
    cudaFree(h->snap.pose);
    cudaFree(h->snap.puck);
    cudaFree(h->snap.cache);
    cudaFree(h->snap.counts);
    cudaFree(h->snap.an);
    h->sd.release();
    for (auto& g : h->groups) {
        if (g.q.load) cudaFree(g.q.load);
        if (g.evPre) cudaEventDestroy(g.evPre);
        if (g.evBig) cudaEventDestroy(g.evBig);
        if (g.done) cudaEventDestroy(g.done);
        if (g.s2) cudaStreamDestroy(g.s2);
        for (int b = 0; b < kQueueBuckets; b++) {
            if (g.evb[b]) cudaEventDestroy(g.evb[b]);
            if (g.sb[b]) cudaStreamDestroy(g.sb[b]);
        }
        if (g.s) cudaStreamDestroy(g.s);
    }
    if (h->evStart) cudaEventDestroy(h->evStart);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return PS_OK;
}

int ps_create(const ps_config* cfg, int device, ps_handle* out) {
    if (!cfg || !out) return setError(PS_E_INVALID, "null argument");
    *out = nullptr;
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
        return setError(PS_E_NODEVICE, "no CUDA device: puckswarm_b200 has no CPU fallback");
    if (device < 0 || device >= nDev) return setError(PS_E_INVALID, "bad device index");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return setError(PS_E_NODEVICE, "device is not sm_100: the kernels are built for sm_100a only");
    ps_engine* e = new ps_engine();
    std::memset(&e->kp, 0, sizeof(e->kp));
    if (!e->hs.init(*cfg)) {
        std::string m = e->hs.error;
        delete e;
        return setError(PS_E_INVALID, m);
    }
    e->device = device;
#define TRY_OR_FREE(expr)                 \
    do {                                  \
        int rc__ = [&]() -> int {         \
            CUDA_TRY(expr);               \
            return PS_OK;                 \
        }();                              \
        if (rc__ != PS_OK) {              \
            ps_destroy(e);                \
            return rc__;                  \
        }                                 \
    } while (0)
    TRY_OR_FREE(cudaSetDevice(device));
    TRY_OR_FREE(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    int prioLow = 0, prioHigh = 0;
    TRY_OR_FREE(cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh));
    if (getenv("PS_NO_PRIORITY")) prioHigh = prioLow;
    TRY_OR_FREE(cudaEventCreateWithFlags(&e->evStart, cudaEventDisableTiming));
    const PhysCfg& pc = e->hs.phys;
    size_t A = cfg->n_arenas;
    TRY_OR_FREE(cudaMalloc(&e->kp.sp.body, A * 6 * pc.d.NB * sizeof(float)));
    TRY_OR_FREE(cudaMalloc(&e->kp.sp.senseAabb, A * 4 * pc.d.NB * sizeof(float)));
    TRY_OR_FREE(cudaMalloc(&e->kp.sp.fat, A * 4 * pc.d.NP * sizeof(float)));
    TRY_OR_FREE(cudaMalloc(&e->kp.sp.ckey, A * pc.MC * sizeof(uint32_t)));
    TRY_OR_FREE(cudaMalloc(&e->kp.sp.cinfo, A * pc.MC * sizeof(uint32_t)));
    TRY_OR_FREE(cudaMalloc(&e->kp.sp.cimp, A * pc.MC * 4 * sizeof(float)));
    TRY_OR_FREE(cudaMalloc(&e->kp.sp.robot, A * pc.d.R * sizeof(ps_robot_state)));
    TRY_OR_FREE(cudaMalloc(&e->kp.sp.arena, A * sizeof(ps_arena_state)));
    TRY_OR_FREE(cudaMemset(e->kp.sp.body, 0, A * 6 * pc.d.NB * sizeof(float)));
    TRY_OR_FREE(cudaMemset(e->kp.sp.senseAabb, 0, A * 4 * pc.d.NB * sizeof(float)));
    TRY_OR_FREE(cudaMemset(e->kp.sp.fat, 0, A * 4 * pc.d.NP * sizeof(float)));
    TRY_OR_FREE(cudaMemset(e->kp.sp.ckey, 0, A * pc.MC * sizeof(uint32_t)));
    TRY_OR_FREE(cudaMemset(e->kp.sp.cinfo, 0, A * pc.MC * sizeof(uint32_t)));
    TRY_OR_FREE(cudaMemset(e->kp.sp.cimp, 0, A * pc.MC * 4 * sizeof(float)));
    TRY_OR_FREE(cudaMemset(e->kp.sp.robot, 0, A * pc.d.R * sizeof(ps_robot_state)));
    TRY_OR_FREE(cudaMemset(e->kp.sp.arena, 0, A * sizeof(ps_arena_state)));
    TRY_OR_FREE(cudaMalloc(&e->dScene, sizeof(Scene)));
    TRY_OR_FREE(cudaMemcpy(e->dScene, &e->hs.scene, sizeof(Scene), cudaMemcpyHostToDevice));
    e->dLut = acquireSharedLut(device, e->hs.sinLut.data());
    if (!e->dLut) {
        ps_destroy(e);
        return setError(PS_E_NOMEM, "cudaMalloc (sin table)");
    }
    TRY_OR_FREE(cudaMalloc(&e->dRobotInit, 2 * sizeof(ps_robot_state)));
    TRY_OR_FREE(cudaMemcpy(e->dRobotInit, e->hs.robotInit, 2 * sizeof(ps_robot_state), cudaMemcpyHostToDevice));
    TRY_OR_FREE(cudaMalloc(&e->dSeeds, A * sizeof(int64_t)));
    TRY_OR_FREE(cudaMalloc(&e->dStepStats, A * 8 * sizeof(int)));
    TRY_OR_FREE(cudaMemset(e->dStepStats, 0, A * 8 * sizeof(int)));
    TRY_OR_FREE(cudaMalloc(&e->dStatsDev, sizeof(StatsDev)));
    TRY_OR_FREE(cudaMalloc(&e->dStats, sizeof(ps_stats)));
    TRY_OR_FREE(cudaMemset(e->dStats, 0, sizeof(ps_stats)));
    TRY_OR_FREE(cudaMalloc(&e->dCmd, 2 * A * pc.d.R * sizeof(float) + 16));
    TRY_OR_FREE(cudaMalloc(&e->dErr, 2 * sizeof(int)));
    TRY_OR_FREE(cudaMalloc(&e->snap.step, A * sizeof(int)));
    TRY_OR_FREE(cudaMalloc(&e->snap.pose, A * pc.d.R * 3 * sizeof(float)));
    TRY_OR_FREE(cudaMalloc(&e->snap.puck, (A * pc.d.P * 2 + 4) * sizeof(float)));
    TRY_OR_FREE(cudaMalloc(&e->snap.cache, A * pc.d.R * PS_MAX_COLOURS * 2 * sizeof(float)));
    TRY_OR_FREE(cudaMalloc(&e->snap.counts, A * pc.d.R * 8 * sizeof(int)));
    TRY_OR_FREE(cudaMalloc(&e->snap.an, A * sizeof(ps_arena_analysis)));
    e->analysisThreshold = cfg->cluster_distance_threshold;
    // working memory: shared memory up to the budget, the rest per slot in HBM
    e->threads = (int)envSize("PS_THREADS", 128);
    if (e->threads < 32 || e->threads > 256 || (e->threads & 31)) e->threads = 128;
    size_t budget = envSize("PS_SMEM_BUDGET", 40 * 1024);
    if (budget > 200 * 1024) budget = 200 * 1024;
    Carver m = {nullptr, budget, 0, nullptr, 0};
    carveWork(pc, m);
    e->smemBytes = m.fastUsed;
    size_t slowStride = (m.slowUsed + 255) & ~(size_t)255;
    int slotsPerSm = (int)envSize("PS_SLOTS_PER_SM", 8);
    e->nSlots = prop.multiProcessorCount * slotsPerSm;
    if ((size_t)e->nSlots > A) e->nSlots = (int)A;
    TRY_OR_FREE(cudaMalloc(&e->kp.slow, (size_t)e->nSlots * slowStride + 256));
    TRY_OR_FREE(cudaMemset(e->kp.slow, 0, (size_t)e->nSlots * slowStride + 256));
    TRY_OR_FREE(cudaFuncSetAttribute(k_reset, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smemBytes));
    TRY_OR_FREE(cudaFuncSetAttribute(k_perturb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smemBytes));
    TRY_OR_FREE(cudaFuncSetAttribute(k_physics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smemBytes));
    // pipeline: one scratch block per arena (all working arrays in HBM / L2), island queue
    {
        Carver mo = {nullptr, 0, 0, nullptr, 0};
        e->kp.workOff = carveWork(pc, mo);
        e->kp.arenaScratchStride = (mo.slowUsed + 255) & ~(size_t)255;
        e->fused = envSize("PS_FUSED", 0) != 0;
        size_t nBlocks = e->fused ? 1 : A;
        TRY_OR_FREE(cudaMalloc(&e->kp.arenaScratch, nBlocks * e->kp.arenaScratchStride + 256));
        TRY_OR_FREE(cudaMemset(e->kp.arenaScratch, 0, nBlocks * e->kp.arenaScratchStride + 256));
        // arena groups, each with its own streams, events and island queue
        int nGroups = (int)envSize("PS_GROUPS", 1);
        if (e->fused || nGroups < 1) nGroups = 1;
        if ((size_t)nGroups > A) nGroups = (int)A;
        // heavy / light arena split: two groups over a device-side arena order, the nHeavy arenas with the longest islands first
        e->nHeavy = (int)envSize("PS_HEAVY_ARENAS", 0);
        if (e->fused || e->nHeavy < 0 || (size_t)e->nHeavy * 2 > A) e->nHeavy = 0;
        if (e->nHeavy > 0) {
            nGroups = 2;
            std::vector<int> ident(A);
            for (size_t i = 0; i < A; ++i) ident[i] = (int)i;
            TRY_OR_FREE(cudaMalloc(&e->dPerm, A * sizeof(int)));
            TRY_OR_FREE(cudaMemcpy(e->dPerm, ident.data(), A * sizeof(int), cudaMemcpyHostToDevice));
            TRY_OR_FREE(cudaMalloc(&e->dScore, A * sizeof(int)));
            TRY_OR_FREE(cudaMemset(e->dScore, 0, A * sizeof(int)));
            e->kp.perm = e->dPerm;
        }
        size_t perGroup = (A + nGroups - 1) / nGroups;
        int qcap = (int)std::min<size_t>((e->nHeavy > 0 ? A : perGroup) * (size_t)pc.d.NB, (size_t)1 << 27);
        TRY_OR_FREE(cudaMalloc(&e->kp.queue.entries, (size_t)nGroups * kQueueBuckets * qcap * sizeof(unsigned)));
        TRY_OR_FREE(cudaMalloc(&e->kp.queue.counts, (size_t)nGroups * kQueueSlots * sizeof(int)));
        e->kp.queue.cap = qcap;
        e->groups.resize(nGroups);
        for (int g = 0; g < nGroups; g++) {
            ps_engine::Group& G = e->groups[g];
            G.a0 = (int)std::min<size_t>(A, g * perGroup);
            G.n = (int)(std::min<size_t>(A, (g + 1) * perGroup) - G.a0);
            if (e->nHeavy > 0) {   // group 0: the light arenas (slots nHeavy..A of the order), group 1: the heavy ones (slots 0..nHeavy)
                G.a0 = g == 0 ? e->nHeavy : 0;
                G.n = g == 0 ? (int)A - e->nHeavy : e->nHeavy;
            }
            G.q.entries = e->kp.queue.entries + (size_t)g * kQueueBuckets * qcap;
            G.q.counts = e->kp.queue.counts + (size_t)g * kQueueSlots;
            G.q.cap = qcap;
            G.q.score = e->dScore;
            // adaptive lane / warp boundary (unless PS_WARP_MIN_CONTACTS pins it): see IslandQueue::load
            G.q.load = nullptr;
            G.q.nArenasLoad = G.n;
            G.q.adaptX100 = (int)envSize("PS_WARP_ADAPT_X100", 40);   // measured (profiles/r02_ab_adapt.log): 13 wins by 3.3 % at 0.13 such islands per arena (ticks 15-75), 17 by 2 % at 0.65 (ticks 105-165)
            if (!getenv("PS_WARP_MIN_CONTACTS") && G.q.adaptX100 > 0) {
                TRY_OR_FREE(cudaMalloc(&G.q.load, 2 * sizeof(int)));
                int init[2] = {1 << 30, 0};   // no history yet: the default boundary
                TRY_OR_FREE(cudaMemcpy(G.q.load, init, sizeof(init), cudaMemcpyHostToDevice));
            }
            TRY_OR_FREE(cudaStreamCreateWithFlags(&G.s, cudaStreamNonBlocking));
            // the island kernels are small and latency-bound: their CTAs go ahead of the other groups' wide kernels
            TRY_OR_FREE(cudaStreamCreateWithPriority(&G.s2, cudaStreamNonBlocking, prioHigh));
            for (int b = 2; b < kQueueBuckets; b++) {
                TRY_OR_FREE(cudaStreamCreateWithPriority(&G.sb[b], cudaStreamNonBlocking, prioHigh));
                TRY_OR_FREE(cudaEventCreateWithFlags(&G.evb[b], cudaEventDisableTiming));
            }
            TRY_OR_FREE(cudaEventCreateWithFlags(&G.evPre, cudaEventDisableTiming));
            TRY_OR_FREE(cudaEventCreateWithFlags(&G.evBig, cudaEventDisableTiming));
            TRY_OR_FREE(cudaEventCreateWithFlags(&G.done, cudaEventDisableTiming));
        }
        TRY_OR_FREE(cudaMalloc(&e->kp.islandHist, 192 * sizeof(unsigned long long)));   // [64..191]: island census (ps_get_island_census)
        TRY_OR_FREE(cudaMemset(e->kp.islandHist, 0, 192 * sizeof(unsigned long long)));
        e->workerBlocks = prop.multiProcessorCount * (int)envSize("PS_WORKER_BLOCKS_PER_SM", 4);
        TRY_OR_FREE(cudaFuncSetAttribute(k_island_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(WarpSlot))));
        TRY_OR_FREE(cudaFuncSetAttribute(k_island_vel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(WarpSlot))));
        e->islandG = (int)envSize("PS_ISLAND_G", 0);
        if (e->islandG != 0 && e->islandG != 2 && e->islandG != 4 && e->islandG != 8) e->islandG = 0;
        e->islandSplit = envSize("PS_ISLAND_SPLIT", 0) != 0;
        TRY_OR_FREE(cudaFuncSetAttribute(k_island_pos<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(16 * sizeof(PosSlot))));
        TRY_OR_FREE(cudaFuncSetAttribute(k_island_pos<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * sizeof(PosSlot))));
        TRY_OR_FREE(cudaFuncSetAttribute(k_island_pos<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(PosSlot))));
        e->posBlocks = prop.multiProcessorCount * (int)envSize("PS_POS_BLOCKS_PER_SM", 4);
        TRY_OR_FREE(cudaFuncSetAttribute(k_island_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((kBatchThreads / 32) * sizeof(LaneSlotsU))));
        TRY_OR_FREE(cudaFuncSetAttribute(k_island_batch_x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LaneSlotsX)));
        e->batchXBlocks = prop.multiProcessorCount * (int)envSize("PS_BATCHX_BLOCKS_PER_SM", 1);
        // stage A / C shared memory by configuration (arenas beyond these sizes work from their scratch blocks)
        e->preThreads = (int)envSize("PS_PRE_THREADS", 64);   // measured (profiles/r02_ab_prepost.log): stage A 1.18 ms per 4096 arenas with 128 threads, 1.08 with 64 (the stage waits on its serial parts; narrow CTAs put more arenas on an SM); stage C is faster with 128
        e->postThreads = (int)envSize("PS_POST_THREADS", 128);
        if (e->preThreads != 32 && e->preThreads != 128) e->preThreads = 64;
        e->preThreadsSet = getenv("PS_PRE_THREADS") != nullptr;
        if (e->postThreads != 32 && e->postThreads != 64) e->postThreads = 128;
        e->preSmem.nb = pc.d.NB <= 512 ? pc.d.NB : 0;
        e->preSmem.capT = (int)std::min<size_t>((size_t)pc.MT, envSize("PS_PRE_SMEM_T", 512));
        e->postNbFat = pc.d.NB <= 512 ? pc.d.NB : 0;
        e->postNpFat = pc.d.NP <= 1024 ? pc.d.NP : 0;
        TRY_OR_FREE(cudaFuncSetAttribute(k_phys_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->preSmem.bytes()));
        TRY_OR_FREE(cudaFuncSetAttribute(k_phys_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)4 * (e->postNbFat + e->postNpFat) * sizeof(float))));
        e->mergeBatch = envSize("PS_MERGE_BATCH", 1) != 0;
        e->bigBlocks = prop.multiProcessorCount * (int)envSize("PS_BIG_BLOCKS_PER_SM", 4);
        e->batchThreads = envSize("PS_BATCH_THREADS", kBatchThreads) == 32 ? 32 : kBatchThreads;
        // PS_BIG_SMEM_KB: shared memory a big-island CTA asks for beyond its slot: the SM then holds fewer warps beside the tick's longest chains
        e->bigSmem = std::min<size_t>(std::max<size_t>(sizeof(BigSlot), envSize("PS_BIG_SMEM_KB", 0) * 1024), 227 * 1024);
        TRY_OR_FREE(cudaFuncSetAttribute(k_island_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->bigSmem));
        e->batchBlocks = prop.multiProcessorCount * (int)envSize("PS_BATCH_BLOCKS_PER_SM", 2);
    }
    e->kp.env.scene = e->dScene;
    e->kp.env.trig.lut = e->dLut;
    e->kp.env.trig.useLut = cfg->sincos_lut;
    e->kp.env.cfg = pc;
    e->kp.env.hot = e->hs.hot;
    e->kp.slowStride = slowStride;
    e->kp.fastCap = e->smemBytes;
    e->kp.nArenas = cfg->n_arenas;
    e->kp.stepInterval = cfg->step_interval;
    e->kp.seeds = e->dSeeds;
    e->kp.robotInit = e->dRobotInit;
    e->kp.nInformed = cfg->n_informed_robots;
    e->kp.puckShiftStep = cfg->puck_shift_step;
    e->kp.puckShuffleStep = cfg->puck_shuffle_step;
    e->kp.stepStats = e->dStepStats;
    *out = e;
    return PS_OK;
}

int ps_load_calibration(ps_handle h, const float* XrYr, const uint16_t* xy, const uint8_t* flags, int n, int img_w, int img_h) {
    if (!h || !XrYr || !xy || !flags || n <= 0) return setError(PS_E_INVALID, "bad calibration arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    if (!h->sense.build(h->hs.cfg, XrYr, xy, flags, n, img_w, img_h)) return setError(PS_E_INVALID, "calibration rejected: " + h->sense.error);
    int rc = h->sd.upload(h->sense, h->hs, h->stream);
    if (rc != PS_OK) return setError(rc, "calibration upload failed: " + h->sd.error);
    return PS_OK;
}

// OR of the arenas' capacity flags and the number of flagged arenas, on the main stream (the caller has joined the groups)
static int scanErrors(ps_handle h, int* flags, int* nFlagged) {
    int host[2] = {0, 0};
    CUDA_TRY(cudaMemsetAsync(h->dErr, 0, 2 * sizeof(int), h->stream));
    const ps_localmap* lm = h->sd.ready ? h->sd.kp.obs.lm : nullptr;
    k_error_scan<<<(h->kp.nArenas + 127) / 128, 128, 0, h->stream>>>(h->kp.sp.arena, lm, h->kp.nArenas, h->hs.phys.d.R, h->dErr);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(host, h->dErr, sizeof(host), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *flags = host[0];
    *nFlagged = host[1];
    return PS_OK;
}

int ps_get_error_flags(ps_handle h, int32_t* flags, int32_t* n_flagged) {
    if (!h) return setError(PS_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    int f = 0, n = 0;
    int rc = scanErrors(h, &f, &n);
    if (rc != PS_OK) return rc;
    if (flags) *flags = f;
    if (n_flagged) *n_flagged = n;
    return PS_OK;
}

int ps_sync(ps_handle h) {
    if (!h) return setError(PS_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    if (!h->isReset) {
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        return PS_OK;
    }
    // a capacity overflow drops a contact / constraint / pair / perceived puck: the arena has left the reference's trajectory
    int f = 0, n = 0;
    int rc = scanErrors(h, &f, &n);
    if (rc != PS_OK) return rc;
    if (f != 0) {
        char msg[256];
        snprintf(msg, sizeof(msg), "capacity overflow in %d arena(s), flags 0x%x (bit0 contact cache, bit1 perceived pucks, bit2 touching "
                 "constraints, bit3 pair buffer, bit4 wall-contact list): raise ps_config.max_contacts; ps_get_state shows the arenas", n, f);
        return setError(PS_E_CAPACITY, msg);
    }
    return PS_OK;
}

__global__ void k_snap_clear(SnapPtrs sn, int nArenas, int R) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nArenas) return;
    sn.step[a] = -1;
    ps_arena_analysis z;
    z.n_snapshots = 0;
    z.steps_to_completion = -1;
    z.twc_num = z.twc_den = z.completion = 0.0;
    sn.an[a] = z;
}
// forget the periodic records and the Analyze accumulators (a new episode starts)
static int clearRecords(ps_handle h) {
    const PhysCfg& pc = h->hs.phys;
    size_t A = (size_t)h->kp.nArenas;
    CUDA_TRY(cudaMemsetAsync(h->snap.pose, 0, A * pc.d.R * 3 * sizeof(float), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->snap.puck, 0, (A * pc.d.P * 2 + 4) * sizeof(float), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->snap.cache, 0, A * pc.d.R * PS_MAX_COLOURS * 2 * sizeof(float), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->snap.counts, 0, A * pc.d.R * 8 * sizeof(int), h->stream));
    k_snap_clear<<<(unsigned)((A + 127) / 128), 128, 0, h->stream>>>(h->snap, (int)A, pc.d.R);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return PS_OK;
}

int ps_reset(ps_handle h, const int64_t* seeds) {
    if (!h || !seeds) return setError(PS_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    CUDA_TRY(cudaMemcpyAsync(h->dSeeds, seeds, (size_t)h->kp.nArenas * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    k_reset<<<h->nSlots, h->threads, h->smemBytes, h->stream>>>(h->kp);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    { int cr = clearRecords(h); if (cr != PS_OK) return cr; }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->updateCount = 0;
    h->stepCount = 0;
    h->isReset = true;
    return PS_OK;
}

// Arena.step's perturbations for arenas [a0, a0 + n) on stream s (nothing is launched when none is configured)
static int launchPerturb(ps_handle h, int a0, int n, cudaStream_t s) {
    if (h->kp.puckShiftStep == -1 && h->kp.puckShuffleStep == -1) return PS_OK;
    k_perturb<<<std::max(1, std::min(h->nSlots, n)), h->threads, h->smemBytes, s>>>(h->kp, a0, n, h->fused ? 0 : 1);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return PS_OK;
}

// Arena.step of one group on a think tick: perturbations, then on-device sense + think (think == 1) or the step counter alone
static constexpr int kStorageInterval = 100;   // Arena.STORAGE_INTERVAL (Arena.java:39)
static int launchThinkStage(ps_handle h, ps_engine::Group& G, int think, int stepCount) {
    int rc = launchPerturb(h, G.a0, G.n, G.s);
    if (rc != PS_OK) return rc;
    int R = h->hs.phys.d.R;
    // the reference's periodic record is due when Arena.stepCount is a multiple of the storage interval (Arena.java:223)
    bool storage = stepCount % kStorageInterval == 0;
    if (think == 1) {
        {
            ScopedTimer timer(h, 0, G.s);
            rc = h->sd.launchSense(h->kp.env, h->kp.sp, G.s, &h->launches, false, G.a0 * R, G.n * R, h->dPerm);
        }
        if (rc != PS_OK) return setError(rc, h->sd.error);
        if (storage) {
            k_snap_pre<<<(G.n * R + 127) / 128, 128, 0, G.s>>>(h->kp.sp, h->snap, h->dPerm, G.a0, G.n, R, kStorageInterval);
            h->launches++;
        }
        {
            ScopedTimer timer(h, 1, G.s);
            rc = h->sd.launchThink(G.a0, G.n, G.s, &h->launches, h->dPerm);
        }
        if (rc != PS_OK) return setError(rc, h->sd.error);
    } else {
        k_count_step<<<(G.n + 127) / 128, 128, 0, G.s>>>(h->kp.sp.arena, h->dPerm, G.a0, G.n);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    if (storage && (think == 1 || h->sd.ready)) {
        // host controllers (think == 2): the poses are those of the ps_sense the caller has just run for this step
        int P = h->hs.phys.d.P;
        size_t sh = (size_t)2 * (P > 0 ? P : 1) * sizeof(int);
        k_snap_post<<<G.n, 128, sh, G.s>>>(h->kp.env, h->kp.sp, h->snap, h->sd.kp.obs.pose, h->dPerm, G.a0, kStorageInterval, think == 1 ? 1 : 0, h->hs.cfg.controller,
                                           h->hs.cfg.n_puck_types, h->hs.cfg.n_pucks, h->analysisThreshold, h->analysisTarget);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return PS_OK;
}

// Stages A and B of one World.step for one arena group: k_phys_pre on the group's stream, the island kernels beside each
// other on the group's island streams; the group's stream is left waiting for them.
static int enqueuePreIslands(ps_handle h, ps_engine::Group& G) {
    ScopedTimer timer(h, 2, G.s);
    KParams kp = h->kp;
    kp.a0 = G.a0;
    kp.queue = G.q;
    CUDA_TRY(cudaMemsetAsync(G.q.counts, 0, kQueueSlots * sizeof(int), G.s));
    {
        ScopedTimer t4(h, 4, G.s);
        // CTA width by batch size, like k_think: narrow CTAs put more arenas on an SM when the launch fills the GPU (1.08 against 1.18 ms
        // per 4096 arenas), a small launch is one arena's latency (1024 arenas: 0.36 ms with 128 threads, 0.43 with 64)
        int pt = h->preThreadsSet ? h->preThreads : (G.n >= 2368 ? 64 : 128);
        k_phys_pre<<<G.n, pt, h->preSmem.bytes(), G.s>>>(kp, h->preSmem);
    }
    CUDA_TRY(cudaEventRecord(G.evPre, G.s));
    {
        ScopedTimer t5(h, 5, G.s);
        // islands are independent: the oversized ones (one warp each) run beside the lane batches
        CUDA_TRY(cudaStreamWaitEvent(G.s2, G.evPre, 0));
        // the largest islands start first, on a stream of their own
        CUDA_TRY(cudaStreamWaitEvent(G.sb[kBucketBig], G.evPre, 0));
        ScopedTimer t8(h, 8, G.sb[kBucketBig]);   // each island kernel from the end of stage A: which one does the tick wait for?
        k_island_big<<<std::max(1, std::min(h->bigBlocks, G.n)), 32, h->bigSmem, G.sb[kBucketBig]>>>(kp);
        t8.stop();
        CUDA_TRY(cudaEventRecord(G.evb[kBucketBig], G.sb[kBucketBig]));
        if (h->hs.phys.laneXMaxContacts > 0) {
            // the wide lane slots next: their lanes run the longest lane chains (up to 28 contacts x 100 sweeps)
            CUDA_TRY(cudaStreamWaitEvent(G.sb[kBucketLaneX], G.evPre, 0));
            ScopedTimer t11(h, 11, G.sb[kBucketLaneX]);
            int xb = std::max(1, std::min(h->batchXBlocks, (G.n + 31) / 32 * 2));
            k_island_batch_x<<<xb, 32, sizeof(LaneSlotsX), G.sb[kBucketLaneX]>>>(kp, xb);
            t11.stop();
            CUDA_TRY(cudaEventRecord(G.evb[kBucketLaneX], G.sb[kBucketLaneX]));
            h->launches++;
        }
        int wb = std::max(1, std::min(h->workerBlocks, G.n * 3));
        ScopedTimer t9(h, 9, G.s2);
        if (h->islandG == 0) {
            k_island_warp<<<wb, 128, 4 * sizeof(WarpSlot), G.s2>>>(kp, 0);
        } else if (h->islandSplit) {
            // the heavy islands (the tick's longest chains) a warp each; the others G lanes each: fewer instructions beside the chains
            k_island_warp<<<std::max(1, wb / 2), 128, 4 * sizeof(WarpSlot), G.s2>>>(kp, 1);
            cudaStream_t sl = G.sb[kBucketLarge];
            CUDA_TRY(cudaStreamWaitEvent(sl, G.evPre, 0));
            k_island_vel<<<wb, 128, 4 * sizeof(WarpSlot), sl>>>(kp, 2);
            int pb = std::max(1, std::min(h->posBlocks, (G.n * 2 * h->islandG + 31) / 32));
            size_t sm = (size_t)(32 / h->islandG) * sizeof(PosSlot);
            if (h->islandG == 2) k_island_pos<2><<<pb, 32, sm, sl>>>(kp, 2);
            else if (h->islandG == 4) k_island_pos<4><<<pb, 32, sm, sl>>>(kp, 2);
            else k_island_pos<8><<<pb, 32, sm, sl>>>(kp, 2);
            CUDA_TRY(cudaEventRecord(G.evb[kBucketLarge], sl));
            CUDA_TRY(cudaStreamWaitEvent(G.s2, G.evb[kBucketLarge], 0));
            h->launches += 2;
        } else {
            // velocity half one warp per island, then the position sweeps G lanes per island (32 / G islands per warp)
            k_island_vel<<<wb, 128, 4 * sizeof(WarpSlot), G.s2>>>(kp, 0);
            int pb = std::max(1, std::min(h->posBlocks, (G.n * 2 * h->islandG + 31) / 32));
            size_t sm = (size_t)(32 / h->islandG) * sizeof(PosSlot);
            if (h->islandG == 2) k_island_pos<2><<<pb, 32, sm, G.s2>>>(kp, 0);
            else if (h->islandG == 4) k_island_pos<4><<<pb, 32, sm, G.s2>>>(kp, 0);
            else k_island_pos<8><<<pb, 32, sm, G.s2>>>(kp, 0);
            h->launches++;
        }
        t9.stop();
        CUDA_TRY(cudaEventRecord(G.evBig, G.s2));
        int bt = h->batchThreads;
        int bb = std::max(1, std::min(h->batchBlocks * (kBatchThreads / bt), (G.n * 12 + bt - 1) / bt));
        if (h->mergeBatch) {
            ScopedTimer t10(h, 10, G.s);
            k_island_batch<<<bb * kLaneBuckets, bt, (bt / 32) * sizeof(LaneSlotsU), G.s>>>(kp, -1, bb);
        } else {
            for (int bucket = kLaneBuckets; bucket >= 2; --bucket) {
                CUDA_TRY(cudaStreamWaitEvent(G.sb[bucket], G.evPre, 0));
                k_island_batch<<<bb, bt, (bt / 32) * sizeof(LaneSlotsU), G.sb[bucket]>>>(kp, bucket, bb);
                CUDA_TRY(cudaEventRecord(G.evb[bucket], G.sb[bucket]));
            }
            k_island_batch<<<bb, bt, (bt / 32) * sizeof(LaneSlotsU), G.s>>>(kp, 1, bb);
            for (int bucket = 2; bucket <= kLaneBuckets; ++bucket) CUDA_TRY(cudaStreamWaitEvent(G.s, G.evb[bucket], 0));
        }
        CUDA_TRY(cudaStreamWaitEvent(G.s, G.evBig, 0));
        CUDA_TRY(cudaStreamWaitEvent(G.s, G.evb[kBucketBig], 0));
        if (h->hs.phys.laneXMaxContacts > 0) CUDA_TRY(cudaStreamWaitEvent(G.s, G.evb[kBucketLaneX], 0));
    }
    h->launches += 3 + (h->mergeBatch ? 1 : kLaneBuckets);
    CUDA_TRY(cudaGetLastError());
    return PS_OK;
}
// Stage C (k_phys_post)
static int enqueuePost(ps_handle h, ps_engine::Group& G) {
    KParams kp = h->kp;
    kp.a0 = G.a0;
    kp.queue = G.q;
    {
        ScopedTimer timer(h, 2, G.s);
        ScopedTimer t6(h, 6, G.s);
        k_phys_post<<<G.n, h->postThreads, (size_t)4 * (h->postNbFat + h->postNpFat) * sizeof(float), G.s>>>(kp, h->postNbFat, h->postNpFat);
    }
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return PS_OK;
}

// RunOffline.step x nTicks for every arena group: {Arena.step on think ticks | nothing} + World.step.
// (Two ways of steering how the groups overlap were built and measured in the steady state. Chaining the groups'
// CTA-per-arena sections by events, so that one group's wide kernels run beside the others' island tails instead of beside
// their wide kernels: slower at every group count -- a wide kernel beside island kernels runs ~2x slower (the island
// kernels hold the SMs' shared memory), and serialised wide sections then bound the step. Starting the groups out of phase
// after a fork, by one k_phys_pre or by a whole lane-batch kernel: no measurable difference.)
// think = 0: no Arena.step (the caller has set the commands and counted the think step), 1: on-device sense + think,
// 2: perturbations + step counter only (external controller)
static int enqueueTicks(ps_handle h, int nTicks, int think) {
    if (h->fused) {
        ps_engine::Group& G = h->groups[0];
        int uc = h->updateCount, interval = h->hs.cfg.step_interval, t = 0;
        int sc = h->stepCount;
        while (t < nTicks) {
            if (think && uc % interval == 0) {
                int rc = launchThinkStage(h, G, think, sc++);
                if (rc != PS_OK) return rc;
            }
            int run = think ? interval - uc % interval : nTicks;
            if (run > nTicks - t) run = nTicks - t;
            ScopedTimer timer(h, 2, G.s);
            k_physics<<<h->nSlots, h->threads, h->smemBytes, G.s>>>(h->kp, run);
            h->launches++;
            CUDA_TRY(cudaGetLastError());
            t += run;
            uc += run;
        }
        return PS_OK;
    }
    // group-major: a group's whole sequence is enqueued before the next group's. Measured against tick-major enqueueing
    // (all groups' tick t, then tick t + 1): 36.1 vs 39.8 ms per think step in the steady state -- the head start keeps the
    // groups out of phase, so one group's CTA-per-arena kernels run beside the other's island tail
    int interval = h->hs.cfg.step_interval;
    if (h->nHeavy > 0) {
        // heavy / light arena split: the groups' membership is redrawn at every think interval (all groups joined, k_classify on the
        // main stream, fork); in between, the few heavy arenas run their long island chains on their own streams while the light
        // ones go through their ticks without waiting for them
        int t = 0, sc = h->stepCount;
        while (t < nTicks) {
            int uc = h->updateCount + t;
            int run = interval - uc % interval;
            if (run > nTicks - t) run = nTicks - t;
            if (uc % interval == 0) {
                h->groupsDirty = true;
                int rc = joinGroups(h);
                if (rc != PS_OK) return rc;
                k_classify<<<1, 1024, 0, h->stream>>>(h->dScore, h->dPerm, h->kp.nArenas, h->nHeavy);
                h->launches++;
                CUDA_TRY(cudaGetLastError());
                if ((rc = forkGroups(h)) != PS_OK) return rc;
            }
            for (auto& G : h->groups) {
                if (G.n == 0) continue;
                for (int k = 0; k < run; ++k) {
                    int rc;
                    if (think && (uc + k) % interval == 0 && (rc = launchThinkStage(h, G, think, sc)) != PS_OK) return rc;
                    if ((rc = enqueuePreIslands(h, G)) != PS_OK) return rc;
                    if ((rc = enqueuePost(h, G)) != PS_OK) return rc;
                }
            }
            if (uc % interval == 0) ++sc;
            t += run;
        }
        return PS_OK;
    }
    for (auto& G : h->groups) {
        if (G.n == 0) continue;
        int sc = h->stepCount;
        for (int t = 0; t < nTicks; ++t) {
            int uc = h->updateCount + t, rc;
            if (think && uc % interval == 0 && (rc = launchThinkStage(h, G, think, sc++)) != PS_OK) return rc;
            if ((rc = enqueuePreIslands(h, G)) != PS_OK) return rc;
            if ((rc = enqueuePost(h, G)) != PS_OK) return rc;
        }
    }
    return PS_OK;
}

// The group streams run free: consecutive ps_step calls just keep appending to them, so a group whose slowest island
// takes long this tick does not hold the others back. They are ordered after the main stream only when that stream got
// new work (fork), and the main stream waits for them only when somebody needs the result (join).
static int forkGroups(ps_handle h) {
    if (!h->mainDirty) return PS_OK;
    CUDA_TRY(cudaEventRecord(h->evStart, h->stream));
    for (auto& G : h->groups) CUDA_TRY(cudaStreamWaitEvent(G.s, h->evStart, 0));
    h->mainDirty = false;
    return PS_OK;
}
static int joinGroups(ps_handle h) {
    h->mainDirty = true;   // whoever joins is about to use the main stream
    if (!h->groupsDirty) return PS_OK;
    for (auto& G : h->groups) {
        CUDA_TRY(cudaEventRecord(G.done, G.s));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, G.done, 0));
    }
    h->groupsDirty = false;
    return PS_OK;
}

int ps_step(ps_handle h, int n_ticks) {
    if (!h || n_ticks < 0) return setError(PS_E_INVALID, "bad argument");
    if (!h->isReset) return setError(PS_E_INVALID, "ps_reset or ps_set_state first");
    CUDA_TRY(cudaSetDevice(h->device));
    bool onDevice = h->hs.cfg.controller != PS_CTRL_EXTERNAL;
    if (onDevice && !h->sd.ready) return setError(PS_E_NOCALIB, "ps_load_calibration has not been called");
    int rc = forkGroups(h);
    if (rc != PS_OK) return rc;
    rc = enqueueTicks(h, n_ticks, onDevice ? 1 : 2);
    if (rc != PS_OK) return rc;
    {
        int interval = h->hs.cfg.step_interval, uc = h->updateCount;
        for (int t = 0; t < n_ticks; ++t)
            if ((uc + t) % interval == 0) h->stepCount++;
    }
    h->updateCount += n_ticks;
    h->groupsDirty = true;
    return PS_OK;
}

int ps_sense(ps_handle h) {
    if (!h) return setError(PS_E_INVALID, "null handle");
    if (!h->sd.ready) return setError(PS_E_NOCALIB, "ps_load_calibration has not been called");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    int rc = h->sd.launchSense(h->kp.env, h->kp.sp, h->stream, &h->launches, h->sd.haveCamera, 0, (int)h->sd.nRobotsTotal);
    if (rc != PS_OK) return setError(rc, h->sd.error);
    return PS_OK;
}

int ps_step_external(ps_handle h, const float* forwards, const float* torque) {
    if (!h || !forwards || !torque) return setError(PS_E_INVALID, "null argument");
    if (!h->isReset) return setError(PS_E_INVALID, "ps_reset or ps_set_state first");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    size_t n = (size_t)h->kp.nArenas * h->hs.phys.d.R;
    CUDA_TRY(cudaMemcpyAsync(h->dCmd, forwards, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->dCmd + n, torque, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    k_set_commands<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(h->kp.sp.robot, h->dCmd, h->dCmd + n, n);
    { int pr = launchPerturb(h, 0, h->kp.nArenas, h->stream); if (pr != PS_OK) return pr; }
    k_count_step<<<(h->kp.nArenas + 127) / 128, 128, 0, h->stream>>>(h->kp.sp.arena, nullptr, 0, h->kp.nArenas);
    h->launches += 2;
    CUDA_TRY(cudaGetLastError());
    if (h->stepCount % kStorageInterval == 0 && h->sd.ready) {
        // the periodic record of a host-controller run: poses of the ps_sense the caller ran for this step, pucks, Analyze row
        int P = h->hs.phys.d.P;
        size_t sh = (size_t)2 * (P > 0 ? P : 1) * sizeof(int);
        k_snap_post<<<h->kp.nArenas, 128, sh, h->stream>>>(h->kp.env, h->kp.sp, h->snap, h->sd.kp.obs.pose, nullptr, 0, kStorageInterval, 0, h->hs.cfg.controller,
                                                          h->hs.cfg.n_puck_types, h->hs.cfg.n_pucks, h->analysisThreshold, h->analysisTarget);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    h->stepCount++;
    // the caller may reuse its arrays as soon as we return
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    int rc = forkGroups(h);
    if (rc != PS_OK) return rc;
    rc = enqueueTicks(h, h->hs.cfg.step_interval, 0);
    if (rc != PS_OK) return rc;
    h->updateCount += h->hs.cfg.step_interval;
    h->groupsDirty = true;
    return PS_OK;
}

// State copies run group by group on the groups' own streams (the arrays are arena-major, so a group's share is one
// slice of each): a group's upload does not wait for the others to finish their step, and its download starts when it is done.
static int copyState(ps_handle h, const ps_state_view* v, bool toHost) {
    const PhysCfg& pc = h->hs.phys;
    cudaMemcpyKind kind = toHost ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
    bool grouped = !h->fused && !h->groups.empty() && !h->dPerm;   // (with the heavy / light split a group is not a slice of the arrays)
    if (grouped) {
        int rc = forkGroups(h);
        if (rc != PS_OK) return rc;
    } else {
        int jr = joinGroups(h);
        if (jr != PS_OK) return jr;
    }
    size_t nSlices = grouped ? h->groups.size() : 1;
    for (size_t g = 0; g < nSlices; ++g) {
        size_t a0 = grouped ? (size_t)h->groups[g].a0 : 0, n = grouped ? (size_t)h->groups[g].n : (size_t)h->kp.nArenas;
        cudaStream_t s = grouped ? h->groups[g].s : h->stream;
        if (n == 0) continue;
#define CP(field, dptr, perArena)                                                                                              \
    if (v->field) {                                                                                                            \
        size_t off = a0 * (size_t)(perArena), bytes = n * (size_t)(perArena);                                                  \
        if (toHost) CUDA_TRY(cudaMemcpyAsync((char*)v->field + off, (const char*)(dptr) + off, bytes, kind, s));              \
        else CUDA_TRY(cudaMemcpyAsync((char*)(dptr) + off, (const char*)v->field + off, bytes, kind, s));                      \
    }
        CP(body, h->kp.sp.body, 6 * pc.d.NB * sizeof(float))
        CP(sense_aabb, h->kp.sp.senseAabb, 4 * pc.d.NB * sizeof(float))
        CP(fat_aabb, h->kp.sp.fat, 4 * pc.d.NP * sizeof(float))
        CP(contact_key, h->kp.sp.ckey, pc.MC * sizeof(uint32_t))
        CP(contact_info, h->kp.sp.cinfo, pc.MC * sizeof(uint32_t))
        CP(contact_impulse, h->kp.sp.cimp, pc.MC * 4 * sizeof(float))
        CP(robot, h->kp.sp.robot, pc.d.R * sizeof(ps_robot_state))
        CP(arena, h->kp.sp.arena, sizeof(ps_arena_state))
#undef CP
    }
    if (grouped) {
        h->groupsDirty = true;
        for (auto& G : h->groups)
            if (G.n) CUDA_TRY(cudaStreamSynchronize(G.s));
    } else {
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return PS_OK;
}

int ps_get_state(ps_handle h, ps_state_view* v) {
    if (!h || !v) return setError(PS_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    return copyState(h, v, true);
}

int ps_set_state(ps_handle h, const ps_state_view* v) {
    if (!h || !v) return setError(PS_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const PhysCfg& pc = h->hs.phys;
    int nA = h->kp.nArenas;
    if (!h->isReset && !(v->body && v->sense_aabb && v->fat_aabb && v->contact_key && v->contact_info && v->contact_impulse && v->robot && v->arena))
        return setError(PS_E_INVALID, "the first ps_set_state of a handle that was never reset must carry every component");
    // the contact arrays are one group: the cache must stay the set of overlapping fat pairs (see the header)
    if ((v->contact_key || v->contact_info || v->contact_impulse) && !(v->contact_key && v->contact_info && v->contact_impulse && v->arena))
        return setError(PS_E_INVALID, "contact_key, contact_info, contact_impulse and arena (n_contacts) must be set together");
    if (v->arena) {
        // all arenas advance in lock step (the think scheduler is the host's): one update_count for the batch
        int uc = v->arena[0].update_count;
        unsigned maxProxy = (unsigned)(kWallFixtures + pc.d.NP);
        for (int a = 0; a < nA; ++a) {
            const ps_arena_state& as = v->arena[a];
            if (as.update_count != uc || as.step_count != v->arena[0].step_count)
                return setError(PS_E_INVALID, "ps_set_state: update_count / step_count differ between arenas (arenas step in lock step)");
            if (as.n_contacts < 0 || as.n_contacts > pc.MC) return setError(PS_E_INVALID, "ps_set_state: n_contacts outside [0, max_contacts]");
            if (v->contact_key) {
                const uint32_t* key = v->contact_key + (size_t)a * pc.MC;
                for (int i = 0; i < as.n_contacts; ++i) {
                    unsigned pa = key[i] >> 16, pb = key[i] & 0xFFFFu;
                    if (pa >= maxProxy || pb >= maxProxy || pa == pb) return setError(PS_E_INVALID, "ps_set_state: contact_key names a proxy outside the arena");
                }
            }
        }
    }
    int rc = copyState(h, v, false);
    if (rc != PS_OK) return rc;
    if (v->arena) {
        h->updateCount = v->arena[0].update_count;
        h->stepCount = v->arena[0].step_count;
    }
    h->isReset = true;
    return PS_OK;
}

int ps_get_observations(ps_handle h, ps_obs_view* v) {
    if (!h || !v) return setError(PS_E_INVALID, "null argument");
    if (!h->sd.ready) return setError(PS_E_NOCALIB, "ps_load_calibration has not been called");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    int rc = h->sd.readObservations(v, h->kp.nArenas, h->hs.phys.d.R, h->stream);
    if (rc != PS_OK) return setError(rc, h->sd.error);
    return PS_OK;
}

// ps_stats of this handle's arenas into h->dStats (device) and *out (host)
static int computeStatsLocal(ps_handle h, float cluster_threshold, ps_stats* out) {
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    StatsDev init;
    std::memset(&init, 0, sizeof(init));
    init.minStc = ~0ULL;
    init.minSumMax = INT_MAX;
    init.maxSumMax = 0;
    init.minTwcBits = ~0ULL;
    CUDA_TRY(cudaMemcpyAsync(h->dStatsDev, &init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));   // `init` is a stack object
    int P = h->hs.phys.d.P;
    size_t sh = (size_t)2 * (P > 0 ? P : 1) * sizeof(int);
    if (sh > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    k_stats<<<h->kp.nArenas, 128, sh, h->stream>>>(h->kp.env, h->kp.sp, h->hs.cfg.n_puck_types, h->hs.cfg.controller, cluster_threshold, h->dStatsDev, h->snap.an);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    StatsDev sd;
    CUDA_TRY(cudaMemcpyAsync(&sd, h->dStatsDev, sizeof(sd), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    std::memset(out, 0, sizeof(*out));
    out->n_arenas = h->kp.nArenas;
    for (int k = 0; k < PS_MAX_COLOURS; k++)
        for (int i = 0; i < 64; i++) out->cluster_size_hist[k][i] = (int64_t)sd.clusterHist[k][i];
    for (int i = 0; i < 8; i++) out->ctrl_state_hist[i] = (int64_t)sd.ctrlState[i];
    out->n_carrying = (int64_t)sd.nCarrying;
    out->n_caches = (int64_t)sd.nCaches;
    double perPuck = h->hs.cfg.n_pucks > 0 ? 100.0 / (double)h->hs.cfg.n_pucks : 0.0;   // Analyze.java:90,197: / the Arena.nPucks property
    out->sum_completion = (double)sd.sumMax * perPuck;
    out->min_completion = (double)sd.minSumMax * perPuck;
    out->max_completion = (double)sd.maxSumMax * perPuck;
    out->n_touching_points = (int64_t)sd.nTouching;
    out->n_contacts = (int64_t)sd.nContacts;
    out->n_completed = (int64_t)sd.nCompleted;
    out->sum_steps_to_completion = (int64_t)sd.sumStc;
    out->n_snapshots = (int64_t)sd.nSnapshots;
    out->min_steps_to_completion = sd.nCompleted ? (int64_t)sd.minStc : INT64_MAX;
    out->max_steps_to_completion = sd.nCompleted ? (int64_t)sd.maxStcPlus1 - 1 : -1;
    out->sum_twc = sd.sumTwc;
    std::memcpy(&out->min_twc, &sd.minTwcBits, sizeof(double));
    std::memcpy(&out->max_twc, &sd.maxTwcBits, sizeof(double));
    CUDA_TRY(cudaMemcpyAsync(h->dStats, out, sizeof(*out), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PS_OK;
}

int ps_compute_stats(ps_handle h, float cluster_threshold, ps_stats* out) {
    if (!h || !out) return setError(PS_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    return computeStatsLocal(h, cluster_threshold, out);
}

// NCCL is the caller's (the communicator comes from the host process): resolve ncclAllReduce from libnccl.so.2 on first use
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
static nccl_allreduce_fn g_ncclAllReduce = nullptr;
static nccl_errstr_fn g_ncclErrStr = nullptr;
static bool loadNccl() {
    static std::mutex m;
    std::lock_guard<std::mutex> lock(m);
    if (g_ncclAllReduce) return true;
    void* lib = dlopen(nullptr, RTLD_NOW);   // already in the process (torch loads its bundled copy)?
    void* sym = lib ? dlsym(lib, "ncclAllReduce") : nullptr;
    if (!sym) {
        const char* env = getenv("PS_NCCL_LIB");
        lib = dlopen(env && *env ? env : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        sym = lib ? dlsym(lib, "ncclAllReduce") : nullptr;
    }
    if (!sym) return false;
    g_ncclErrStr = (nccl_errstr_fn)dlsym(lib, "ncclGetErrorString");
    g_ncclAllReduce = (nccl_allreduce_fn)sym;
    return true;
}

int ps_get_stats(ps_handle h, float cluster_threshold, ps_stats* out, void* nccl_comm) {
    if (!h || !out) return setError(PS_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = computeStatsLocal(h, cluster_threshold, out);
    if (rc != PS_OK || !nccl_comm) return rc;
    if (!loadNccl()) return setError(PS_E_INVALID, "ps_get_stats: ncclAllReduce not found (libnccl.so.2; PS_NCCL_LIB overrides the path)");
    // one collective per reduction group of ps_stats, in place on the device copy, on the handle's stream
    enum { kInt64 = 4, kFloat64 = 8, kSum = 0, kMax = 2, kMin = 3 };
    struct Seg { size_t off, count; int type, op; };
    const Seg segs[] = {
        {offsetof(ps_stats, n_arenas), (offsetof(ps_stats, sum_completion) - offsetof(ps_stats, n_arenas)) / 8, kInt64, kSum},
        {offsetof(ps_stats, sum_completion), 2, kFloat64, kSum},
        {offsetof(ps_stats, min_steps_to_completion), 1, kInt64, kMin},
        {offsetof(ps_stats, max_steps_to_completion), 1, kInt64, kMax},
        {offsetof(ps_stats, min_completion), 2, kFloat64, kMin},
        {offsetof(ps_stats, max_completion), 2, kFloat64, kMax},
    };
    for (const Seg& sg : segs) {
        char* p = (char*)h->dStats + sg.off;
        int nr = g_ncclAllReduce(p, p, sg.count, sg.type, sg.op, nccl_comm, h->stream);
        if (nr != 0) return setError(PS_E_CUDA, std::string("ncclAllReduce: ") + (g_ncclErrStr ? g_ncclErrStr(nr) : "error"));
    }
    CUDA_TRY(cudaMemcpyAsync(out, h->dStats, sizeof(*out), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PS_OK;
}

int ps_get_snapshot(ps_handle h, ps_snapshot_view* v) {
    if (!h || !v) return setError(PS_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    const PhysCfg& pc = h->hs.phys;
    size_t A = (size_t)h->kp.nArenas;
    if (v->step_count) CUDA_TRY(cudaMemcpyAsync(v->step_count, h->snap.step, A * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (v->robot_pose) CUDA_TRY(cudaMemcpyAsync(v->robot_pose, h->snap.pose, A * pc.d.R * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (v->puck_xy && pc.d.P > 0) CUDA_TRY(cudaMemcpyAsync(v->puck_xy, h->snap.puck, A * pc.d.P * 2 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (v->cache_xy) CUDA_TRY(cudaMemcpyAsync(v->cache_xy, h->snap.cache, A * pc.d.R * PS_MAX_COLOURS * 2 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (v->state_counts) CUDA_TRY(cudaMemcpyAsync(v->state_counts, h->snap.counts, A * pc.d.R * 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PS_OK;
}

int ps_get_analysis(ps_handle h, ps_arena_analysis* out) {
    if (!h || !out) return setError(PS_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    CUDA_TRY(cudaMemcpyAsync(out, h->snap.an, (size_t)h->kp.nArenas * sizeof(ps_arena_analysis), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PS_OK;
}

int ps_analysis_config(ps_handle h, float cluster_threshold, float target_percent) {
    if (!h || !(cluster_threshold >= 0.0f)) return setError(PS_E_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    h->analysisThreshold = cluster_threshold;
    h->analysisTarget = target_percent;
    int rc = clearRecords(h);
    if (rc != PS_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PS_OK;
}

int ps_stats_device_ptr(ps_handle h, void** dptr, int64_t* n_bytes) {
    if (!h || !dptr || !n_bytes) return setError(PS_E_INVALID, "null argument");
    *dptr = h->dStats;
    *n_bytes = (int64_t)sizeof(ps_stats);
    return PS_OK;
}

int ps_get_dims(ps_handle h, int32_t* n_bodies, int32_t* n_proxies, int32_t* max_contacts, int32_t* occ_w, int32_t* occ_h) {
    if (!h) return setError(PS_E_INVALID, "null handle");
    if (n_bodies) *n_bodies = h->hs.phys.d.NB;
    if (n_proxies) *n_proxies = h->hs.phys.d.NP;
    if (max_contacts) *max_contacts = h->hs.phys.MC;
    if (occ_w) *occ_w = h->sense.occW;
    if (occ_h) *occ_h = h->sense.occH;
    return PS_OK;
}

int ps_get_step_stats(ps_handle h, int32_t* out /* [n_arenas][8] */) {
    if (!h || !out) return setError(PS_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    CUDA_TRY(cudaMemcpyAsync(out, h->dStepStats, (size_t)h->kp.nArenas * 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PS_OK;
}

int ps_get_sense_stats(ps_handle h, int32_t* out /* [n_arenas][n_robots][8] */) {
    if (!h || !out) return setError(PS_E_INVALID, "null argument");
    if (!h->sd.ready) return setError(PS_E_NOCALIB, "ps_load_calibration has not been called");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    CUDA_TRY(cudaMemcpyAsync(out, h->sd.kp.senseDbg, h->sd.nRobotsTotal * 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PS_OK;
}

int ps_get_island_hist(ps_handle h, int64_t* out /*[64]*/, int reset) {
    if (!h || !out) return setError(PS_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    CUDA_TRY(cudaMemcpyAsync(out, h->kp.islandHist, 64 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    if (reset) CUDA_TRY(cudaMemsetAsync(h->kp.islandHist, 0, 64 * sizeof(int64_t), h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PS_OK;
}

int ps_get_island_census(ps_handle h, int64_t* out /*[128]*/, int reset) {
    if (!h || !out) return setError(PS_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    CUDA_TRY(cudaMemcpyAsync(out, h->kp.islandHist + 64, 128 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    if (reset) CUDA_TRY(cudaMemsetAsync(h->kp.islandHist + 64, 0, 128 * sizeof(int64_t), h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PS_OK;
}

int ps_profile(ps_handle h, int enable) {
    if (!h) return setError(PS_E_INVALID, "null handle");
    h->profiling = enable != 0;
    return PS_OK;
}

int ps_get_kernel_ms(ps_handle h, double* ms, int64_t* counts) {
    if (!h || !ms || !counts) return setError(PS_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    { int jr = joinGroups(h); if (jr != PS_OK) return jr; }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (auto& t : h->timed) {
        float f = 0;
        if (cudaEventElapsedTime(&f, t.a, t.b) == cudaSuccess) {
            h->kernelMs[t.cls] += f;
            h->kernelCount[t.cls] += 1;
        }
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    h->timed.clear();
    for (int i = 0; i < 12; i++) {
        ms[i] = h->kernelMs[i];
        counts[i] = h->kernelCount[i];
        h->kernelMs[i] = 0;
        h->kernelCount[i] = 0;
    }
    return PS_OK;
}

int ps_join(ps_handle h) {
    if (!h) return setError(PS_E_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    return joinGroups(h);
}
int ps_num_groups(ps_handle h) { return h ? (int)h->groups.size() : 0; }
int64_t ps_kernel_launches(ps_handle h) { return h ? h->launches : 0; }
void* ps_stream(ps_handle h) { return h ? (void*)h->stream : nullptr; }

}  // extern "C"
