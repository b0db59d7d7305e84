// puckswarm_b200 -- device side of sensing / thinking / statistics: table upload, kernels, launch helpers.
// (CUDA only; the per-robot and per-arena programs themselves are in ps_sense.cuh.)
#pragma once
#include <cuda_runtime.h>
#include "ps_host.hpp"
#include "ps_sense.cuh"

namespace ps {

struct SenseKParams {
    PhysEnv env;
    StatePtrs sp;
    SenseTables T;
    SenseCfg sc;
    ObsPtrs obs;
    int nArenas;
    size_t fastCap;
    SenseWork workOff;   // byte offsets of the camera stage's shared-memory arrays
    int r0;            // first robot (k_sense, k_localmap) / first arena (k_think) of the launch
    const int* perm;   // arena order of the launch (heavy / light arena split) or null = identity: slot i of the launch is arena perm[i]
    int* senseDbg;     // [nRobots][8] phase diagnostics of the last sense
    SenseHead* senseHead;   // [nRobots] camera stage -> local-map stage
    BlobRec* blobs;         // [nRobots][kMaxBlobs]
};

extern __shared__ __align__(16) char g_senseSmem[];

// one CTA per (arena, robot): camera + local map
__global__ void __launch_bounds__(256) k_sense(SenseKParams p) {
    __shared__ int scan[40];
    CtaCtx ctx;
    ctx.scanScratch = scan;
    SenseWork wk = rebaseSense(p.workOff, g_senseSmem);
    int R = p.env.cfg.d.R;
    int rg = p.r0 + (int)blockIdx.x;
    int a = rg / R, r = rg - a * R;
    if (p.perm) {
        a = p.perm[a];
        rg = a * R + r;
    }
    ArenaState st = arenaState(p.sp, p.env.cfg, a);
    size_t ri = (size_t)rg;
    int nCells = p.T.occW * p.T.occH;
    uint8_t* cam = p.obs.camera ? p.obs.camera + ri * (size_t)(p.T.imgW * p.T.imgH) : nullptr;
    __shared__ float poseS[4];
    SenseOut so;
    so.head = p.senseHead + ri;
    so.blobs = p.blobs + ri * kMaxBlobs;
    senseRobot(ctx, p.env, p.T, p.sc, wk, st, r, cam, p.obs.occupancy + ri * nCells, poseS, so);
    __syncthreads();
    if (threadIdx.x < 3) p.obs.pose[ri * 3 + threadIdx.x] = poseS[threadIdx.x];
    if (threadIdx.x < 8) p.senseDbg[ri * 8 + threadIdx.x] = wk.phase[threadIdx.x];
}

// LocalMap stage: one warp per robot over the whole batch. The local map is built in shared memory: lane 0 walks the
// blob list like LocalMap.extractPucks / extractClusters do, all lanes share the scan of the occupancy grid for ROBOT
// cells, the warp finishes the cluster filter (clusters in order, the puck pairs of a cluster shared by the lanes), and
// the warp writes the struct out coalesced.
__global__ void __launch_bounds__(128) k_localmap(SenseKParams p, int nRobots) {
    __shared__ ps_localmap lmS[4];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int idx = blockIdx.x * 4 + warp;
    if (idx >= nRobots) return;
    int ri = p.r0 + idx;
    if (p.perm) {
        int R = p.env.cfg.d.R;
        int slot = ri / R;
        ri = p.perm[slot] * R + (ri - slot * R);
    }
    int nCells = p.T.occW * p.T.occH;
    ps_localmap& lm = lmS[warp];
    uint32_t* w = (uint32_t*)&lm;
    for (int i = lane; i < (int)(sizeof(ps_localmap) / 4); i += 32) w[i] = 0u;
    __syncwarp();
    const uint8_t* occ = p.obs.occupancy + (size_t)ri * nCells;
    if (lane == 0) localMapPucks(p.T, p.sc, p.senseHead[ri], p.blobs + (size_t)ri * kMaxBlobs, p.sp.robot + ri, lm);
    __syncwarp();
    if (lm.n_pucks <= 32) localMapClustersWarp(p.sc, lm, lane);
    else if (lane == 0) localMapClusters(p.sc, lm);
    __syncwarp();
    localMapRobotFilter(p.T, occ, lm, lane, 32);
    __syncwarp();
    localMapFinalFilterWarp(p.sc, p.senseHead[ri], lm, lane);
    __syncwarp();
    uint32_t* g = (uint32_t*)(p.obs.lm + ri);
    for (int i = lane; i < (int)(sizeof(ps_localmap) / 4); i += 32) g[i] = w[i];
}

// one CTA per arena: controllers in robot order (shared java.util.Random)
__global__ void __launch_bounds__(128) k_think(SenseKParams p) {
    __shared__ int scan[40];
    __shared__ unsigned primary[kSectors];
    __shared__ unsigned limits[2];
    __shared__ int flag[4];
    __shared__ JRandom rng;
    CtaCtx ctx;
    ctx.scanScratch = scan;
    ThinkWork tw;
    tw.primary = primary;
    tw.limits = limits;
    tw.flag = flag;
    int R = p.env.cfg.d.R;
    int a = p.r0 + (int)blockIdx.x;
    if (p.perm) a = p.perm[a];
    size_t r0 = (size_t)a * R;
    int nCells = p.T.occW * p.T.occH;
    thinkArena(ctx, p.T, p.sc, tw, R, p.obs.occupancy + r0 * nCells, p.obs.lm + r0, p.obs.pose + r0 * 3, p.sp.robot + r0, p.sp.arena + a, &rng);
}

__global__ void k_count_step(ps_arena_state* arena, const int* perm, int a0, int nArenas) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nArenas) arena[perm ? perm[a0 + i] : a0 + i].step_count += 1;
}

__global__ void k_set_commands(ps_robot_state* robot, const float* fwd, const float* tq, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        robot[i].cmd_forwards = fwd[i];
        robot[i].cmd_torque = tq[i];
    }
}

// Per-arena experiment statistics (PositionList.getClusterStats, src/arena/PositionList.java:85-96, over the true puck
// positions of each colour; Analyze.java:187-218 percent completion = 100 * sum_k max cluster size_k / nPucks),
// controller-state occupancy, carry and cache counts, contact counts. One CTA per arena, integer atomics into one ps_stats.
struct StatsDev {
    unsigned long long clusterHist[PS_MAX_COLOURS][64];
    unsigned long long ctrlState[8];
    unsigned long long nCarrying, nCaches, sumMax, nTouching, nContacts;
    unsigned long long nCompleted, sumStc, nSnapshots;
    unsigned long long minStc, maxStcPlus1;          // over the completed arenas; maxStcPlus1 = 0 when none
    int minSumMax, maxSumMax;                        // extrema of sum_k max cluster size_k
    double sumTwc;
    unsigned long long minTwcBits, maxTwcBits;       // time-weighted completion is >= 0: its bit pattern orders like the value
};
__global__ void __launch_bounds__(128) k_stats(PhysEnv env, StatePtrs sp, int nTypes, int controller, float threshold, StatsDev* out, const ps_arena_analysis* an) {
    extern __shared__ int sh[];   // parent[P], size[P]
    const Dims& d = env.cfg.d;
    int a = blockIdx.x;
    ArenaState st = arenaState(sp, env.cfg, a);
    int P = d.P, NB = d.NB;
    int* parent = sh;
    int* size = sh + P;
    int perType = P / nTypes;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        parent[i] = i;
        size[i] = 0;
    }
    __syncthreads();
    CtaCtx ctx;
    ctx.scanScratch = nullptr;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        V2 pi = mk(st.body[i], st.body[NB + i]);
        int k = i / perType;
        for (int j = k * perType; j < i; ++j) {
            V2 pj = mk(st.body[j], st.body[NB + j]);
            bool same = pi.x == pj.x && pi.y == pj.y;
            if (!same && dist(pi, pj) < threshold) ufUnion(ctx, parent, i, j);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += blockDim.x) atomicAdd(&size[ufFind(parent, i)], 1);
    __syncthreads();
    __shared__ int maxOfType[PS_MAX_COLOURS];
    if (threadIdx.x < PS_MAX_COLOURS) maxOfType[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        int s = size[i];
        if (s > 0) {
            int k = i / perType;
            atomicAdd(&out->clusterHist[k][s < 63 ? s : 63], 1ULL);
            atomicMax(&maxOfType[k], s);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int sum = 0;
        for (int k = 0; k < nTypes; ++k) sum += maxOfType[k];
        atomicAdd(&out->sumMax, (unsigned long long)sum);
        atomicMin(&out->minSumMax, sum);
        atomicMax(&out->maxSumMax, sum);
        // analysis.Analyze's per-repetition results (accumulated by k_snap_post)
        const ps_arena_analysis& A = an[a];
        atomicAdd(&out->nSnapshots, (unsigned long long)A.n_snapshots);
        if (A.steps_to_completion >= 0) {
            atomicAdd(&out->nCompleted, 1ULL);
            atomicAdd(&out->sumStc, (unsigned long long)A.steps_to_completion);
            atomicMin(&out->minStc, (unsigned long long)A.steps_to_completion);
            atomicMax(&out->maxStcPlus1, (unsigned long long)A.steps_to_completion + 1ULL);
        }
        double twc = A.twc_den > 0.0 ? A.twc_num / A.twc_den : 0.0;   // Analyze.java:212-216
        atomicAdd(&out->sumTwc, twc);
        atomicMin(&out->minTwcBits, (unsigned long long)__double_as_longlong(twc));
        atomicMax(&out->maxTwcBits, (unsigned long long)__double_as_longlong(twc));
    }
    const ps_robot_state* rs = sp.robot + (size_t)a * d.R;
    for (int r = threadIdx.x; r < d.R; r += blockDim.x) {
        atomicAdd(&out->ctrlState[rs[r].ctrl_state & 7], 1ULL);
        if (rs[r].lm_carrying) atomicAdd(&out->nCarrying, 1ULL);
        if (controller == PS_CTRL_CACHECONS) {
            int c = 0;
            for (int k = 0; k < PS_MAX_COLOURS; ++k) c += rs[r].cache_valid[k] ? 1 : 0;
            if (c) atomicAdd(&out->nCaches, (unsigned long long)c);
        }
    }
    int nc = *st.nContacts;
    int touching = 0;
    for (int i = threadIdx.x; i < nc; i += blockDim.x) touching += (int)(st.cinfo[i] & 3u);
    if (touching) atomicAdd(&out->nTouching, (unsigned long long)touching);
    if (threadIdx.x == 0) atomicAdd(&out->nContacts, (unsigned long long)nc);
}

// ---- the reference's periodic record (Arena.java:223-243, CacheConsController.java:313-333,531-551, BHDController.java:103-123)
// and analysis.Analyze's per-repetition statistics (Analyze.java:176-218), kept per arena on the device.
// At a think step whose Arena.stepCount is a multiple of the storage interval the reference writes robot poses and puck positions
// (after the perturbations, before World.step), every robot's cache points after this step's think, and its recentStateCounts
// (counted up to and including this step, then cleared). k_snap_pre runs before the think kernel (it saves the counters the
// controllers are about to clear), k_snap_post after it.
struct SnapPtrs {
    int* step;            // [A] Arena.stepCount of the record, -1 = none yet
    float* pose;          // [A][R][3]
    float* puck;          // [A][P][2]
    float* cache;         // [A][R][PS_MAX_COLOURS][2], NaN = null
    int* counts;          // [A][R][8]
    ps_arena_analysis* an;   // [A]
};

__global__ void k_snap_pre(StatePtrs sp, SnapPtrs sn, const int* perm, int a0, int nArenas, int R, int interval) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nArenas * R) return;
    int slot = a0 + i / R;
    int a = perm ? perm[slot] : slot;
    if (sp.arena[a].step_count % interval != 0) return;
    size_t ri = (size_t)a * R + (size_t)(i % R);
    const ps_robot_state& rs = sp.robot[ri];
    int* out = sn.counts + ri * 8;
    for (int k = 0; k < 8; ++k) out[k] = rs.state_counts[k];
}

// One CTA per arena. onDevice: the controllers ran on the device (cache points and state counters are theirs); otherwise only
// the arena-level record and the analysis are kept (a host controller owns its own dumps).
__global__ void __launch_bounds__(128) k_snap_post(PhysEnv env, StatePtrs sp, SnapPtrs sn, const float* obsPose, const int* perm, int a0, int interval, int onDevice,
                                                   int controller, int nTypes, int nPucksCfg, float threshold, float targetPercent) {
    extern __shared__ int sh[];   // parent[P], size[P]
    const Dims& d = env.cfg.d;
    int a = perm ? perm[a0 + blockIdx.x] : a0 + (int)blockIdx.x;
    int stepCount = sp.arena[a].step_count - 1;   // the think step that has just run
    if (stepCount < 0 || stepCount % interval != 0) return;
    ArenaState st = arenaState(sp, env.cfg, a);
    int P = d.P, NB = d.NB, R = d.R;
    if (threadIdx.x == 0) sn.step[a] = stepCount;
    for (int i = threadIdx.x; i < R * 3; i += blockDim.x) sn.pose[(size_t)a * R * 3 + i] = obsPose[(size_t)a * R * 3 + i];
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        sn.puck[((size_t)a * P + i) * 2] = st.body[i];
        sn.puck[((size_t)a * P + i) * 2 + 1] = st.body[NB + i];
    }
    if (onDevice) {
        for (int i = threadIdx.x; i < R * PS_MAX_COLOURS; i += blockDim.x) {
            int r = i / PS_MAX_COLOURS, k = i % PS_MAX_COLOURS;
            const ps_robot_state& rs = sp.robot[(size_t)a * R + r];
            bool valid = controller == PS_CTRL_CACHECONS && rs.cache_valid[k];
            float* out = sn.cache + (((size_t)a * R + r) * PS_MAX_COLOURS + k) * 2;
            out[0] = valid ? (float)rs.cache_x[k] : __int_as_float(0x7fc00000);
            out[1] = valid ? (float)rs.cache_y[k] : __int_as_float(0x7fc00000);
        }
        for (int r = threadIdx.x; r < R; r += blockDim.x) sn.counts[((size_t)a * R + r) * 8 + (sp.robot[(size_t)a * R + r].ctrl_state & 7)] += 1;
    }
    // percent completion of this record: PositionList.getClusterStats per colour over the true puck positions
    int* parent = sh;
    int* size = sh + P;
    int perType = nTypes > 0 ? P / nTypes : P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        parent[i] = i;
        size[i] = 0;
    }
    __syncthreads();
    CtaCtx ctx;
    ctx.scanScratch = nullptr;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        V2 pi = mk(st.body[i], st.body[NB + i]);
        int k = i / perType;
        for (int j = k * perType; j < i; ++j) {
            V2 pj = mk(st.body[j], st.body[NB + j]);
            bool same = pi.x == pj.x && pi.y == pj.y;
            if (!same && dist(pi, pj) < threshold) ufUnion(ctx, parent, i, j);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += blockDim.x) atomicAdd(&size[ufFind(parent, i)], 1);
    __syncthreads();
    __shared__ int maxOfType[PS_MAX_COLOURS];
    if (threadIdx.x < PS_MAX_COLOURS) maxOfType[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        if (size[i] > 0) atomicMax(&maxOfType[i / perType], size[i]);
    __syncthreads();
    if (threadIdx.x == 0) {
        int sum = 0;
        for (int k = 0; k < nTypes; ++k) sum += maxOfType[k];
        double pc = nPucksCfg > 0 ? 100.0 * (double)sum / (double)nPucksCfg : 0.0;   // Analyze.java:197
        ps_arena_analysis& an = sn.an[a];
        an.n_snapshots += 1;
        an.completion = pc;
        if (an.steps_to_completion == -1 && pc >= (double)targetPercent) an.steps_to_completion = stepCount;   // :202-204
        an.twc_num += (double)stepCount * pc / 100.0;                                                          // :206
        an.twc_den += (double)stepCount;                                                                       // :212-216
    }
}

struct SenseDevice {
    bool ready = false;
    std::string error;
    SenseKParams kp;
    std::vector<void*> allocs;
    size_t senseSmem = 0;
    int threads = 256;
    bool haveCamera = false;
    size_t nRobotsTotal = 0;

    int fail(int code, const std::string& m) {
        error = m;
        return code;
    }
    template <class Tp>
    int up(const std::vector<Tp>& v, const Tp** dst) {
        void* p = nullptr;
        size_t bytes = v.size() * sizeof(Tp);
        if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) return fail(PS_E_NOMEM, "cudaMalloc (calibration tables)");
        allocs.push_back(p);
        if (bytes && cudaMemcpy(p, v.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) return fail(PS_E_CUDA, "cudaMemcpy (calibration tables)");
        *dst = (const Tp*)p;
        return PS_OK;
    }
    template <class Tp>
    int alloc(Tp** dst, size_t count) {
        void* p = nullptr;
        if (cudaMalloc(&p, count * sizeof(Tp) + 16) != cudaSuccess) return fail(PS_E_NOMEM, "cudaMalloc (observation buffers)");
        if (cudaMemset(p, 0, count * sizeof(Tp) + 16) != cudaSuccess) return fail(PS_E_CUDA, "cudaMemset");
        allocs.push_back(p);
        *dst = (Tp*)p;
        return PS_OK;
    }
    void release() {
        for (void* p : allocs) cudaFree(p);
        allocs.clear();
        ready = false;
    }
    int upload(const SenseSetup& ss, const HostSetup& hs, cudaStream_t stream) {
        release();
        std::memset(&kp, 0, sizeof(kp));
        kp.T = ss.T;
        kp.sc = ss.sc;
        int rc;
#define UP(vec, field) \
    if ((rc = up(ss.vec, &kp.T.field)) != PS_OK) return rc;
        UP(aXr, aXr)
        UP(aYr, aYr)
        UP(aRect, aRect)
        UP(aImg, aImg)
        UP(aBin, aBin)
        UP(aPix, aPix)
        UP(cellRect, cellRect)
        UP(calibIdx, calibIdx)
        UP(cXr, cXr)
        UP(cYr, cYr)
        UP(cFlags, cFlags)
        UP(cellSrc, cellSrc)
        UP(vfhBeta, vfhBeta)
        UP(vfhMag, vfhMag)
        UP(vfhStartK, vfhStartK)
        UP(vfhStopK, vfhStopK)
        UP(vfhSide, vfhSide)
#undef UP
        size_t A = hs.cfg.n_arenas, R = hs.phys.d.R;
        nRobotsTotal = A * R;
        size_t nCells = (size_t)ss.occW * ss.occH, nImg = (size_t)ss.T.imgW * ss.T.imgH;
        if ((rc = alloc(&kp.obs.occupancy, nRobotsTotal * nCells)) != PS_OK) return rc;
        if ((rc = alloc(&kp.obs.lm, nRobotsTotal)) != PS_OK) return rc;
        if ((rc = alloc(&kp.obs.pose, nRobotsTotal * 3)) != PS_OK) return rc;
        if ((rc = alloc(&kp.senseDbg, nRobotsTotal * 8)) != PS_OK) return rc;
        if ((rc = alloc(&kp.senseHead, nRobotsTotal)) != PS_OK) return rc;
        if ((rc = alloc(&kp.blobs, nRobotsTotal * kMaxBlobs)) != PS_OK) return rc;
        // the camera image is an optional observation: kept only when the whole batch fits in 256 MiB
        haveCamera = nRobotsTotal * nImg <= ((size_t)256 << 20);
        kp.obs.camera = nullptr;
        if (haveCamera && (rc = alloc(&kp.obs.camera, nRobotsTotal * nImg)) != PS_OK) return rc;
        Carver m = {nullptr, (size_t)200 * 1024, 0, nullptr, 0};
        kp.workOff = carveSense(kp.T, m);
        if (m.slowUsed) return fail(PS_E_INVALID, "camera image too large for shared memory");
        senseSmem = m.fastUsed;
        kp.fastCap = senseSmem;
        if (cudaFuncSetAttribute(k_sense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)senseSmem) != cudaSuccess)
            return fail(PS_E_CUDA, "cudaFuncSetAttribute(k_sense)");
        const char* tv = getenv("PS_SENSE_THREADS");
        threads = tv && *tv ? atoi(tv) : 256;
        if (threads < 32 || threads > 256 || (threads & 31)) threads = 256;
        kp.nArenas = hs.cfg.n_arenas;
        ready = true;
        return PS_OK;
    }
    void bind(const PhysEnv& env, const StatePtrs& sp) {
        kp.env = env;
        kp.sp = sp;
    }
    int launchSense(const PhysEnv& env, const StatePtrs& sp, cudaStream_t stream, int64_t* launches, bool camera, int firstRobot, int nRobots, const int* perm = nullptr) {
        bind(env, sp);
        SenseKParams p = kp;
        p.perm = perm;
        if (!camera) p.obs.camera = nullptr;
        p.r0 = firstRobot;
        k_sense<<<(unsigned)nRobots, threads, senseSmem, stream>>>(p);
        k_localmap<<<(unsigned)((nRobots + 3) / 4), 128, 0, stream>>>(p, nRobots);
        (*launches) += 2;
        if (cudaGetLastError() != cudaSuccess) return fail(PS_E_CUDA, "k_sense / k_localmap launch failed");
        return PS_OK;
    }
    int launchThink(int firstArena, int nArenas, cudaStream_t stream, int64_t* launches, const int* perm = nullptr) {
        SenseKParams p = kp;
        p.perm = perm;
        p.r0 = firstArena;
        // CTA width by batch size. The stage is the robots' controllers one after the other (shared java.util.Random), with a
        // barrier between every parallel piece and the serial state machine. When the launch fills the GPU, one warp per arena
        // wins (a one-warp barrier costs nothing and four times as many arenas are resident: 2.60 / 1.92 / 1.51 ms per 4096
        // arenas with 128 / 64 / 32 threads); a small launch is one arena's latency, which 128 threads cut (1024 arenas:
        // 0.72 ms with 128 threads, 1.19 ms with 32). PS_THINK_THREADS overrides.
        int thinkThreads = nArenas >= 2368 ? 32 : (nArenas >= 1184 ? 64 : 128);
        {
            const char* v = getenv("PS_THINK_THREADS");
            if (v && *v) thinkThreads = atoi(v) < 64 ? 32 : (atoi(v) < 128 ? 64 : 128);
        }
        k_think<<<nArenas, thinkThreads, 0, stream>>>(p);
        (*launches)++;
        if (cudaGetLastError() != cudaSuccess) return fail(PS_E_CUDA, "k_think launch failed");
        return PS_OK;
    }
    int readObservations(ps_obs_view* v, int nArenas, int R, cudaStream_t stream) {
        size_t n = (size_t)nArenas * R;
        size_t nCells = (size_t)kp.T.occW * kp.T.occH, nImg = (size_t)kp.T.imgW * kp.T.imgH;
        if (v->camera) {
            if (!haveCamera) return fail(PS_E_CAPACITY, "camera images are not kept for a batch this large");
            if (cudaMemcpyAsync(v->camera, kp.obs.camera, n * nImg, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return fail(PS_E_CUDA, "copy camera");
        }
        if (v->occupancy && cudaMemcpyAsync(v->occupancy, kp.obs.occupancy, n * nCells, cudaMemcpyDeviceToHost, stream) != cudaSuccess)
            return fail(PS_E_CUDA, "copy occupancy");
        if (v->localmap && cudaMemcpyAsync(v->localmap, kp.obs.lm, n * sizeof(ps_localmap), cudaMemcpyDeviceToHost, stream) != cudaSuccess)
            return fail(PS_E_CUDA, "copy localmap");
        if (v->pose && cudaMemcpyAsync(v->pose, kp.obs.pose, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, stream) != cudaSuccess)
            return fail(PS_E_CUDA, "copy pose");
        if (cudaStreamSynchronize(stream) != cudaSuccess) return fail(PS_E_CUDA, "sync after observation copy");
        return PS_OK;
    }
};

}  // namespace ps
