// puckswarm_b200 -- sensing and thinking: grid camera, local map, VFH+ and the shipped controllers.
//
// Replaces, per robot and think step (src/arena/Arena.java:197-217):
//   sensors.SimSuite.sense      -> GridCamera.sense (src/sensors/GridCamera.java:48-83) over
//                                  PointSensor.pointSense (src/sensors/PointSensor.java:85-114), then
//   localmap.LocalMap.update    (src/localmap/LocalMap.java:183-437) with sensors.BlobFinder
//                                  (src/sensors/BlobFinder.java:30-101), then
//   controllers.*.computeDesired: CacheConsController (src/controllers/CacheConsController.java:336-705),
//                                  BHDController (src/controllers/BHDController.java:126-191) over
//                                  localmap.VFHPlus (src/localmap/VFHPlus.java:226-480) and MovementUtils /
//                                  MovementCommand (src/localmap/MovementUtils.java:13-58, MovementCommand.java:20-23).
// senseRobot is executed by one CTA per robot (robots only read the world while sensing); thinkArena by one
// CTA per arena because the reference shares one java.util.Random between the robots of an arena, in robot order.
#pragma once
#include <vector>
#include <string>
#include <algorithm>
#include "ps_layout.cuh"

namespace ps {

static constexpr int kMaxBlobs = 128;
static constexpr int kMaxCand = 256;
static constexpr int kMaxColPix = 2048;   // puck-coloured pixels listed per robot (by the camera stage, as it finds them); beyond that the blob passes scan the whole rectangle
static constexpr int kMaxBins = 64;       // robot-frame grid over the pixel footprint
static constexpr int kBinCap = 30;        // candidates per bin; an overfull bin falls back to the full candidate list
static constexpr int kSectors = PS_N_SECTORS;

// Static tables derived from the camera calibration (built on the host by SenseSetup)
struct SenseTables {
    int imgW, imgH;
    int nActive;                 // calibrated pixels that are not gripperBody: the only ones with a PointSensor result
    const float* aXr;            // [nActive] ground-plane point in the robot frame
    const float* aYr;
    const uint16_t* aRect;       // [nActive] index in the bounding rectangle, x-major: (x - x0) * RH + (y - y0)
    const uint16_t* aImg;        // [nActive] index in the camera image, y * imgW + x
    const uint8_t* aBin;         // [nActive] bin of the pixel's ground point in the robot-frame grid below
    const struct PixRec* aPix;   // [nActive] the four tables above in one 16-byte record, SORTED BY BIN (what the camera stage reads)
    uint16_t binStart[65];       // [kMaxBins + 1] range of each bin's pixels in aPix
    float binSize;               // bins of binSize x binSize cm over [bbXmin, bbXmax] x [bbYmin, bbYmax]
    int binsX, binsY;
    int x0, y0, RW, RH;          // bounding rectangle of the active pixels
    const int16_t* calibIdx;     // [imgW * imgH] at x * imgH + y: row in cXr / cYr / cFlags, -1 = uncalibrated
    const float* cXr;
    const float* cYr;
    const uint8_t* cFlags;       // bit0 gripperBody, bit1 gripperHole
    int nHolePixels;
    int occW, occH;
    float maxYr;
    const int16_t* cellSrc;      // [occW * occH] active pixel whose type the cell takes (last writer), -1 = stays NOTHING
    const uint16_t* cellRect;    // [occW * occH] aRect[cellSrc], 0xFFFF = stays NOTHING
    // VFHPlus.initDataStructures (src/localmap/VFHPlus.java:170-217), per occupancy cell
    const float* vfhBeta;
    const float* vfhMag;
    const int16_t* vfhStartK;
    const int16_t* vfhStopK;
    const uint8_t* vfhSide;      // bit0: within the safety radius of the left turning centre, bit1: right
    float maxRange;              // max distance of an active pixel from the camera
    float bbXmin, bbXmax, bbYmin, bbYmax;
    int nPuckTypes, pucksPerType;
};

struct alignas(16) PixRec {
    float xr, yr;
    uint32_t rectBin;            // aRect | aBin << 16
    uint32_t img;                // aImg
};

struct SenseCfg {
    float clusterThreshold, carryLo, carryHi;
    float tauLo, tauHi;
    int controller, rngMode;
    // CacheCons
    float k1, k2, predictionDistSqd, stDev, kAbsolute, kRelative, forgetProb, homeSeparation;
    int pickUpTime, pushInTime, backupTime, exileTime, spitOutTime, cacheSelection, cacheSeparation, ignorePucks;
    int bhdPushInTime, bhdBackupTime;
    int placementTime;   // ProbSeekController.PLACEMENT_TIME
    int nInformed;
};

// Observation buffers in HBM (what ps_get_observations returns)
struct ObsPtrs {
    uint8_t* camera;      // optional [A][R][imgH * imgW]
    uint8_t* occupancy;   // [A][R][occW * occH]
    ps_localmap* lm;      // [A][R]
    float* pose;          // [A][R][3]
};

// What the camera stage hands to the local-map stage, per robot
struct BlobRec {
    uint8_t x0, x1, y0, y1;   // bounding box in image pixels
    uint16_t area;
    uint16_t seed;            // x * imgH + y of BlobFinder's seed pixel, 0xFFFF = the blob never seeds (only row imgH-1)
    uint8_t colour, pad0, pad1, pad2;
};
struct SenseHead {
    int nBlobs, leftSum, rightSum, overflow;
};
struct SenseOut {
    SenseHead* head;
    BlobRec* blobs;           // [kMaxBlobs]
};

// Per-CTA scratch of senseRobot
struct SenseWork {
    uint8_t* img;         // [RW * RH] SensedType per rectangle pixel
    int* parent;          // [RW * RH] union-find forest of the blob labelling
    float* candX;         // [kMaxCand]
    float* candY;
    int* candBody;        // body index
    float* candLx;        // Fixture.m_aabb of the sensed fixture
    float* candLy;
    float* candUx;
    float* candUy;
    float* candC;         // robots: cos / sin of the body transform
    float* candS;
    uint8_t* candType;    // pucks: SensedType of the colour (body / pucksPerType)
    float* candRx;        // candidate origin in the sensing robot's frame (binning only, never decides a pixel)
    float* candRy;
    uint8_t* binCount;    // [kMaxBins]
    uint8_t* binClass;    // [kMaxBins] 0 = general; 1 = no candidate can reach the bin and it lies inside the enclosure's central
                          //   band with a margin: every pixel is NOTHING; 2 = inside the band, candidates present: skip inFreeSpace
    uint8_t* binList;     // [kMaxBins][kBinCap] candidate indices in body-list order
    int* blobArea;        // [kMaxBlobs]
    int* blobX0;
    int* blobX1;
    int* blobY0;
    int* blobY1;
    int* blobSeed;
    int* blobRoot;
    int* blobOrder;
    uint16_t* colList;    // [kMaxColPix] the pixels that show a puck colour: rectangle index (camera stage), then (x - x0) << 8 | (y - y0)
    int* counters;        // [8]
    int* phase;           // [8] diagnostics: cycles / 16 of the phases of senseRobot (thread 0)
};
inline SenseWork carveSense(const SenseTables& t, Carver& cv) {   // host only
    SenseWork w;
    int n = t.RW * t.RH;
    w.counters = (int*)cv.take(8 * sizeof(int));
    w.phase = (int*)cv.take(8 * sizeof(int));
    w.img = (uint8_t*)cv.take((size_t)n);
    w.parent = (int*)cv.take((size_t)n * sizeof(int));
    // the candidate / bin arrays (camera stage) and the blob / colour-pixel arrays (blob labelling, afterwards) share
    // one block: the two stages never overlap
    auto up16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    size_t szA = 11 * up16(kMaxCand * sizeof(float)) + up16(kMaxCand) + 2 * up16(kMaxBins) + up16(kMaxBins * kBinCap);
    size_t szB = 8 * up16(kMaxBlobs * sizeof(int));
    char* blk = (char*)cv.take(szA > szB ? szA : szB);
    {
        char* q = blk;
        auto sub = [&](size_t b) { char* r = q; q += up16(b); return (void*)r; };
        w.candX = (float*)sub(kMaxCand * sizeof(float));
        w.candY = (float*)sub(kMaxCand * sizeof(float));
        w.candBody = (int*)sub(kMaxCand * sizeof(int));
        w.candLx = (float*)sub(kMaxCand * sizeof(float));
        w.candLy = (float*)sub(kMaxCand * sizeof(float));
        w.candUx = (float*)sub(kMaxCand * sizeof(float));
        w.candUy = (float*)sub(kMaxCand * sizeof(float));
        w.candC = (float*)sub(kMaxCand * sizeof(float));
        w.candS = (float*)sub(kMaxCand * sizeof(float));
        w.candType = (uint8_t*)sub(kMaxCand);
        w.candRx = (float*)sub(kMaxCand * sizeof(float));
        w.candRy = (float*)sub(kMaxCand * sizeof(float));
        w.binCount = (uint8_t*)sub(kMaxBins);
        w.binClass = (uint8_t*)sub(kMaxBins);
        w.binList = (uint8_t*)sub(kMaxBins * kBinCap);
    }
    {
        char* q = blk;
        auto sub = [&](size_t b) { char* r = q; q += up16(b); return (void*)r; };
        w.blobArea = (int*)sub(kMaxBlobs * sizeof(int));
        w.blobX0 = (int*)sub(kMaxBlobs * sizeof(int));
        w.blobX1 = (int*)sub(kMaxBlobs * sizeof(int));
        w.blobY0 = (int*)sub(kMaxBlobs * sizeof(int));
        w.blobY1 = (int*)sub(kMaxBlobs * sizeof(int));
        w.blobSeed = (int*)sub(kMaxBlobs * sizeof(int));
        w.blobRoot = (int*)sub(kMaxBlobs * sizeof(int));
        w.blobOrder = (int*)sub(kMaxBlobs * sizeof(int));
    }
    // written by the camera stage, read by the blob labelling: outside the shared block
    w.colList = (uint16_t*)cv.take(kMaxColPix * sizeof(uint16_t));
    return w;
}

// SenseWork whose pointers are byte offsets (carved from a null base on the host) -> pointers into this CTA's shared memory
PS_HD SenseWork rebaseSense(const SenseWork& o, char* base) {
    SenseWork w;
    size_t b = (size_t)base;
#define RB(f) w.f = (decltype(w.f))((size_t)o.f + b);
    RB(img) RB(parent) RB(candX) RB(candY) RB(candBody) RB(candLx) RB(candLy) RB(candUx) RB(candUy) RB(candC) RB(candS) RB(candType) RB(candRx) RB(candRy)
    RB(binCount) RB(binClass) RB(binList) RB(blobArea) RB(blobX0) RB(blobX1) RB(blobY0) RB(blobY1) RB(blobSeed) RB(blobRoot) RB(blobOrder) RB(colList)
    RB(counters) RB(phase)
#undef RB
    return w;
}

PS_HD int ufFind(const int* parent, int i) {
    int p = parent[i];
    while (p != i) {
        i = p;
        p = parent[i];
    }
    return i;
}
// find with path halving; parents only ever decrease, so atomicMin keeps concurrent updates consistent
template <class Ctx>
PS_HD int ufFindHalve(Ctx& ctx, int* parent, int i) {
    int p = parent[i];
    while (p != i) {
        int gp = parent[p];
        if (gp != p) ctx.atomicMinI(&parent[i], gp);
        i = p;
        p = gp;
    }
    return i;
}
template <class Ctx>
PS_HD void ufUnion(Ctx& ctx, int* parent, int a, int b) {
    for (;;) {
        a = ufFindHalve(ctx, parent, a);
        b = ufFindHalve(ctx, parent, b);
        if (a == b) return;
        if (a < b) {
            int t = a;
            a = b;
            b = t;
        }
        // hang the larger root under the smaller one; retry if somebody re-parented it first
        int old = ctx.atomicMinRet(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

// LocalMap.isReachable (src/localmap/LocalMap.java:548-558), double arithmetic
PS_HD bool lmReachable(float vx, float vy) {
    double circleYr = 12.5, circleXr = 0, radius = circleYr + 0;
    double dXr = vx - circleXr;
    double a = fabs(circleYr - vy), b = fabs(circleYr + vy);
    double dYr = a < b ? a : b;   // Math.min
    return sqrt(dXr * dXr + dYr * dYr) > radius;
}

// LocalMap.extractOccupancy (src/localmap/LocalMap.java:210-229): every cell takes the type of its last writer in (x, y)
// scan order (a static table); getFreeerSide tallies on the way
template <class Ctx>
PS_HD void senseOccupancy(Ctx& ctx, const SenseTables& T, const SenseWork& wk, uint8_t* occOut) {
    int nCells = T.occW * T.occH;
    int leftSum = 0, rightSum = 0;
    PS_FOR(c, nCells) {
        int src = T.cellRect[c];
        uint8_t t = src == 0xFFFF ? (uint8_t)PS_T_NOTHING : wk.img[src];
        occOut[c] = t;
        if (t == PS_T_ROBOT || t == PS_T_WALL) {
            // LocalMap.getFreeerSide tallies (src/localmap/LocalMap.java:565-587)
            if (c < (T.occW / 2) * T.occH) leftSum++;
            else rightSum++;
        }
    }
    if (leftSum) ctx.atomicAddI(&wk.counters[1], leftSum);
    if (rightSum) ctx.atomicAddI(&wk.counters[2], rightSum);
}

// BlobFinder (src/sensors/BlobFinder.java:30-101): 8-connected components of equal puck colour over the coloured pixels the
// camera stage listed; leaves the blob table in wk.blob* and the blob count in wk.counters[0]. `ctx` may be a whole CTA or
// a single warp (the passes are short: a few hundred pixels).
template <class Ctx>
PS_HD void senseBlobs(Ctx& ctx, const SenseTables& T, const SenseWork& wk, int nRect) {
    // ---- BlobFinder: 8-connected components of equal puck colour.
    // (1) vertical runs: the rectangle is stored column-major, so a run of equal colour inside a column is a contiguous
    //     index range. Every coloured pixel (the camera stage listed them) walks up to its run's first pixel (runs are a
    //     few pixels long) and labels itself with that index.
    int nCol = wk.counters[6];
    bool listed = nCol <= kMaxColPix;        // otherwise fall back to scanning the rectangle
    int nScan = listed ? nCol : nRect;
    PS_FOR(n, nScan) {
        int i = listed ? (int)wk.colList[n] : n;
        uint8_t t = wk.img[i];
        if (t >= PS_MAX_COLOURS) continue;
        int cx = i / T.RH;
        int colBase = cx * T.RH;
        int start = i;
        while (start > colBase && wk.img[start - 1] == t) --start;
        wk.parent[i] = start;
        if (listed) wk.colList[n] = (uint16_t)((cx << 8) | (i - colBase));
    }
    ctx.sync();
    // decode entry n of the scan into (i, x, y); returns false for pixels without a puck colour
#define PS_COLPIX(n, i, x, y)                                   \
    int i, x, y;                                                \
    if (listed) {                                               \
        int xy_ = wk.colList[n];                                \
        x = xy_ >> 8;                                           \
        y = xy_ & 255;                                          \
        i = x * T.RH + y;                                       \
    } else {                                                    \
        i = n;                                                  \
        x = i / T.RH;                                           \
        y = i - x * T.RH;                                       \
        if (wk.img[i] >= PS_MAX_COLOURS) continue;              \
    }
    // (2) merge runs of neighbouring columns (8-connectivity: rows y-1, y, y+1 of column x-1). A pair of runs is united
    //     once, by the first pixel at which the two runs face each other.
    PS_FOR(n, nScan) {
        PS_COLPIX(n, i, x, y)
        uint8_t t = wk.img[i];
        if (x == 0) continue;
        bool upSame = y > 0 && wk.img[i - 1] == t;
        int j0 = i - T.RH;
        for (int dy = -1; dy <= 1; ++dy) {
            int yy = y + dy;
            if (yy < 0 || yy >= T.RH) continue;
            if (wk.img[j0 + dy] != t) continue;
            // the pixel above already joined these two runs if it saw the neighbour run one row up
            if (upSame && yy - 1 >= 0 && wk.img[j0 + dy - 1] == t) continue;
            ufUnion(ctx, wk.parent, wk.parent[i], wk.parent[j0 + dy]);
        }
    }
    ctx.sync();
    // roots -> blob slots (any order: the local-map stage sorts the blobs by colour and seed)
    PS_FOR(n, nScan) {
        PS_COLPIX(n, i, x, y)
        if (wk.parent[i] == i) {
            int pos = ctx.atomicAddI(&wk.counters[0], 1);
            if (pos < kMaxBlobs) {
                wk.blobRoot[pos] = i;
                wk.blobArea[pos] = 0;
                wk.blobX0[pos] = 1 << 30;
                wk.blobX1[pos] = -(1 << 30);
                wk.blobY0[pos] = 1 << 30;
                wk.blobY1[pos] = -(1 << 30);
                wk.blobSeed[pos] = 1 << 30;
            } else {
                wk.counters[5] = 1;
            }
        }
    }
    ctx.sync();
    int nBlobs = wk.counters[0];
    if (nBlobs > kMaxBlobs) nBlobs = kMaxBlobs;
    // flatten every pixel to its root, then to "-(slot + 1)" via the root
    PS_FOR(n, nScan) {
        PS_COLPIX(n, i, x, y)
        if (wk.parent[i] != i) wk.parent[i] = ufFind(wk.parent, wk.parent[i]);
    }
    ctx.sync();
    PS_FOR(s, nBlobs) wk.parent[wk.blobRoot[s]] = -(s + 1);
    ctx.sync();
    // blob statistics, one update per vertical run (not per pixel): the run's first pixel walks down to its end
    PS_FOR(n, nScan) {
        PS_COLPIX(n, i, x, y)
        uint8_t t = wk.img[i];
        if (y > 0 && wk.img[i - 1] == t) continue;   // not the first pixel of its run
        int yEnd = y;
        while (yEnd + 1 < T.RH && wk.img[i + (yEnd + 1 - y)] == t) ++yEnd;
        int p = wk.parent[i];
        if (p >= 0) p = wk.parent[p];
        int s = -p - 1;
        if (s < 0 || s >= kMaxBlobs) continue;   // blob table overflow
        int px = T.x0 + x, py = T.y0 + y, pyEnd = T.y0 + yEnd;
        ctx.atomicAddI(&wk.blobArea[s], yEnd - y + 1);
        ctx.atomicMinI(&wk.blobX0[s], px);
        ctx.atomicMaxI(&wk.blobX1[s], px);
        ctx.atomicMinI(&wk.blobY0[s], py);
        ctx.atomicMaxI(&wk.blobY1[s], pyEnd);
        // BlobFinder seeds only at y < yMax (= imgH - 1); the seed fixes the blob's position in the list. Within a run the
        // smallest seed index is that of its first pixel
        if (py < T.imgH - 1) ctx.atomicMinI(&wk.blobSeed[s], (px * T.imgH + py));
    }
#undef PS_COLPIX
}

// ---------------------------------------------------------------------------------------------------------------
// senseRobot: camera image -> occupancy grid -> blobs -> perceived pucks -> clusters -> filtered clusters
template <class Ctx>
PS_HD void senseRobot(Ctx& ctx, const PhysEnv& e, const SenseTables& T, const SenseCfg& sc, const SenseWork& wk, const ArenaState& st,
                      int robot, uint8_t* camOut, uint8_t* occOut, float* poseOut, const SenseOut& out) {
    const Dims& d = e.cfg.d;
    const Scene& scene = *e.scene;
    int NB = d.NB;
    int self = d.P + robot;
    // SimAPS / Body.getTransform of the sensing robot
    Xf xf = makeXf(e.trig, st.body[0 * NB + self], st.body[1 * NB + self], st.body[2 * NB + self], scene.robot.lcx, scene.robot.lcy);
    if (ctx.tid() == 0) {
        for (int i = 0; i < 8; i++) wk.counters[i] = 0;
        poseOut[0] = xf.px;
        poseOut[1] = xf.py;
        poseOut[2] = st.body[2 * NB + self];
    }
    int nRect = T.RW * T.RH;
    {
        // (the array is 16-byte aligned and padded to a multiple of 4)
        uint32_t* img4 = (uint32_t*)wk.img;
        PS_FOR(i, (nRect + 3) / 4) img4[i] = 0x01010101u * (uint32_t)PS_T_HIDDEN;
    }
    if (camOut) {
        PS_FOR(i, T.imgW * T.imgH) camOut[i] = PS_T_HIDDEN;
    }
    ctx.sync();
    long long tc0 = phaseClock();
    // ---- candidate bodies in world body-list order (newest first): robots R-1..0, pucks P-1..0
    // conservative cull: a body can only be sensed by a pixel within 22 cm (+ a margin) of its origin
    int nCand = 0;
    {
        float reach = T.maxRange + 22.5f;
        float reachSq = reach * reach;
        int nRound = roundUpThreads(ctx, NB);
        for (int base = 0; base < nRound; base += ctx.nthreads()) {
            int li = base + ctx.tid();          // position in the body list
            bool take = false;
            int b = -1;
            float bx = 0, by = 0, bc = 1.0f, bs = 0.0f, rlx = 0, rly = 0;
            if (li < NB) {
                b = NB - 1 - li;
                if (b != self) {
                    if (b >= d.P) {
                        Xf o = makeXf(e.trig, st.body[0 * NB + b], st.body[1 * NB + b], st.body[2 * NB + b], scene.robot.lcx, scene.robot.lcy);
                        bx = o.px;
                        by = o.py;
                        bc = o.c;
                        bs = o.s;
                    } else {
                        bx = st.body[0 * NB + b];
                        by = st.body[1 * NB + b];
                    }
                    float dx = bx - xf.px, dy = by - xf.py;
                    if (dx * dx + dy * dy <= reachSq) {
                        // robot-frame bounding box of the pixel footprint
                        float lx = dx * xf.c + dy * xf.s, ly = -dx * xf.s + dy * xf.c;
                        rlx = lx;
                        rly = ly;
                        take = lx >= T.bbXmin - 22.5f && lx <= T.bbXmax + 22.5f && ly >= T.bbYmin - 22.5f && ly <= T.bbYmax + 22.5f;
                    }
                }
            }
            int total;
            int pos = nCand + ctx.exscanFlag(take, total);
            if (take) {
                if (pos < kMaxCand) {
                    wk.candX[pos] = bx;
                    wk.candY[pos] = by;
                    wk.candBody[pos] = b;
                    wk.candType[pos] = (uint8_t)(b < d.P ? b / T.pucksPerType : PS_T_ROBOT);
                    Aabb bb = loadAabb(st.senseAabb, NB, b);
                    wk.candLx[pos] = bb.lx;
                    wk.candLy[pos] = bb.ly;
                    wk.candUx[pos] = bb.ux;
                    wk.candUy[pos] = bb.uy;
                    wk.candC[pos] = bc;
                    wk.candS[pos] = bs;
                    wk.candRx[pos] = rlx;
                    wk.candRy[pos] = rly;
                } else {
                    wk.counters[5] = 1;
                }
            }
            nCand += total;
        }
        if (nCand > kMaxCand) nCand = kMaxCand;
    }
    ctx.sync();
    // ---- bin the candidates over the pixel footprint (robot frame). A bin lists, in body-list order, every candidate
    // whose sensed fixture could contain a point of the bin: pucks reach 2.1 cm from their origin, robot hulls < 13 cm.
    // The lists only prune the search; every listed candidate still goes through the reference's exact tests.
    int nBins = T.binsX * T.binsY;
#if defined(__CUDA_ARCH__)
    // one warp per bin: 32 candidates per ballot, appended in candidate (= body-list) order
    for (int bin = (int)(threadIdx.x >> 5); bin < nBins; bin += (int)(blockDim.x >> 5)) {
        int lane = (int)(threadIdx.x & 31u);
        int bxi = bin / T.binsY, byi = bin - bxi * T.binsY;
        float x0 = T.bbXmin + bxi * T.binSize, y0 = T.bbYmin + byi * T.binSize;
        float x1 = x0 + T.binSize, y1 = y0 + T.binSize;
        int n = 0;
        for (int base = 0; base < nCand; base += 32) {
            int c = base + lane;
            bool in = false;
            if (c < nCand) {
                float rad = wk.candBody[c] < d.P ? 2.25f : 13.0f;
                float cx = wk.candRx[c], cy = wk.candRy[c];
                in = !(cx + rad < x0 || cx - rad > x1 || cy + rad < y0 || cy - rad > y1);
            }
            unsigned m = __ballot_sync(0xFFFFFFFFu, in);
            int pos = n + __popc(m & ((1u << lane) - 1u));
            if (in && pos < kBinCap) wk.binList[bin * kBinCap + pos] = (uint8_t)c;
            n += __popc(m);
        }
        if (lane == 0) wk.binCount[bin] = (uint8_t)(n > kBinCap ? 255 : n);
    }
#else
    PS_FOR(bin, nBins) {
        int bxi = bin / T.binsY, byi = bin - bxi * T.binsY;
        float x0 = T.bbXmin + bxi * T.binSize, y0 = T.bbYmin + byi * T.binSize;
        float x1 = x0 + T.binSize, y1 = y0 + T.binSize;
        int n = 0;
        for (int c = 0; c < nCand; ++c) {
            float rad = wk.candBody[c] < d.P ? 2.25f : 13.0f;
            float cx = wk.candRx[c], cy = wk.candRy[c];
            if (cx + rad < x0 || cx - rad > x1 || cy + rad < y0 || cy - rad > y1) continue;
            if (n < kBinCap) wk.binList[bin * kBinCap + n] = (uint8_t)c;
            ++n;
        }
        wk.binCount[bin] = (uint8_t)(n > kBinCap ? 255 : n);
    }
#endif
    ctx.sync();
    // ---- classify the bins. RoundedRectangleEnclosure.inFreeSpace(g, 0) is true for every g with |g.x| < W/2 - R and
    // |g.y| < H/2 (its first clause). A bin is a rectangle in the robot frame; if its four corners, mapped to the world, lie
    // inside that band by a margin far above the rounding of the per-pixel transform (1e-2 cm vs ~1e-5), so does every pixel
    // of the bin (the band is convex), and the per-pixel test can only say "free". With no candidate body listed either, and
    // no sensable wall fixture within reach, PointSensor.pointSense returns NOTHING for all of the bin's pixels.
    {
        const float margin = 0.01f;
        float bandX = scene.W / 2 - scene.R - margin, bandY = scene.H / 2 - margin;
        PS_FOR(bin, nBins) {
            int bxi = bin / T.binsY, byi = bin - bxi * T.binsY;
            float x0 = T.bbXmin + bxi * T.binSize - margin, y0 = T.bbYmin + byi * T.binSize - margin;
            float x1 = x0 + T.binSize + 2 * margin, y1 = y0 + T.binSize + 2 * margin;
            bool inside = true;
            for (int k = 0; k < 4; ++k) {
                V2 g = mulX(xf, mk((k & 1) ? x1 : x0, (k & 2) ? y1 : y0));
                inside = inside && fabsf(g.x) < bandX && fabsf(g.y) < bandY;
            }
            // (the wall body's own fixtures are sensed only within 22 cm of the world origin: leave such bins to the general path)
            bool wallNear = false;
            if (scene.wallSensable)
                for (int k = 0; k < 4; ++k) {
                    V2 g = mulX(xf, mk((k & 1) ? x1 : x0, (k & 2) ? y1 : y0));
                    wallNear = wallNear || distSq(mk(0.0f, 0.0f), g) <= 40.0f * 40.0f;
                }
            int cls = 0;
            if (inside && !wallNear) cls = wk.binCount[bin] == 0 ? 1 : 2;
            wk.binClass[bin] = (uint8_t)cls;
        }
    }
    ctx.sync();
    long long tc1 = phaseClock();
    // ---- PointSensor.pointSense for every active pixel
    const Shape& hull = scene.shapes[kShapeRobot0];
    float innerR = scene.shapes[kShapePuck0 + 1].radius;
    float innerRSq = innerR * innerR;
    FreeSpace fs = makeFreeSpace(scene, 0.0f);
    int visited = 0;   // candidate bodies examined by this thread's pixels (diagnostic: K of SURVEY 8(d))
    auto sensePixel = [&](const PixRec& px) {
        int bin = (int)(px.rectBin >> 16);
        int cls = wk.binClass[bin];
        if (cls == 1) {   // nothing can be seen anywhere in this bin
            wk.img[px.rectBin & 0xFFFFu] = (uint8_t)PS_T_NOTHING;
            if (camOut) camOut[px.img] = (uint8_t)PS_T_NOTHING;
            return;
        }
        V2 g = mulX(xf, mk(px.xr, px.yr));
        int type = PS_T_NOTHING;
        if (cls == 0 && !inFreeSpaceFast(fs, g)) {
            type = PS_T_WALL;
        } else {
            bool hit = false;
            int nList = wk.binCount[bin];
            bool full = nList == 255;   // overfull bin: walk the whole candidate list
            if (full) nList = nCand;
            for (int li = 0; li < nList && !hit; ++li) {
                int c = full ? li : (int)wk.binList[bin * kBinCap + li];
                ++visited;
                V2 bp = mk(wk.candX[c], wk.candY[c]);
                float d2 = distSq(bp, g);
                int b = wk.candBody[c];
                if (b < d.P) {
                    // inner disc, centre = body origin: (g - c).(g - c) <= r * r, same magnitude as d2. The reference tests the
                    // 22 cm body radius and the fixture's AABB first; all three are pure predicates, so the cheap, rarely true
                    // one goes first here (it implies d2 <= 484: the disc radius is 2.1), and the box (four loads) is only
                    // looked at for the few pixels inside the disc
                    if (d2 <= innerRSq && g.x <= wk.candUx[c] && g.x >= wk.candLx[c] && g.y <= wk.candUy[c] && g.y >= wk.candLy[c]) {
                        type = wk.candType[c];
                        hit = true;
                    }
                } else {
                    if (d2 > 484.0f) continue;
                    if (!(g.x <= wk.candUx[c] && g.x >= wk.candLx[c] && g.y <= wk.candUy[c] && g.y >= wk.candLy[c])) continue;
                    Xf o;
                    o.px = bp.x;
                    o.py = bp.y;
                    o.c = wk.candC[c];
                    o.s = wk.candS[c];
                    if (shapeTestPoint(hull, o, g)) {
                        type = PS_T_ROBOT;
                        hit = true;
                    }
                }
            }
            if (!hit && scene.wallSensable) {
                // the wall body (origin (0,0)) comes last in the body list; its fixtures are two-vertex polygons
                if (distSq(mk(0.0f, 0.0f), g) <= 484.0f) {
                    for (int f = kWallFixtures - 1; f >= 0 && !hit; --f) {
                        const Shape& s = scene.shapes[f];
                        Aabb bb = shapeAabb(s, scene.wallXf);
                        if (g.x <= bb.ux && g.x >= bb.lx && g.y <= bb.uy && g.y >= bb.ly && shapeTestPoint(s, scene.wallXf, g)) {
                            type = PS_T_WALL;
                            hit = true;
                        }
                    }
                }
            }
        }
        wk.img[px.rectBin & 0xFFFFu] = (uint8_t)type;
        if (camOut) camOut[px.img] = (uint8_t)type;
        if (type < PS_MAX_COLOURS) {   // the blob labelling only ever looks at these
            int pos = ctx.atomicAddI(&wk.counters[6], 1);
            if (pos < kMaxColPix) wk.colList[pos] = (uint16_t)(px.rectBin & 0xFFFFu);
        }
    };
    {
        // two pixels per trip: both table records are in flight before the first is used
        int nT = ctx.nthreads();
        for (int k0 = ctx.tid(); k0 < T.nActive; k0 += 2 * nT) {
            int k1 = k0 + nT;
            bool two = k1 < T.nActive;
            PixRec pa = T.aPix[k0];
            PixRec pb = T.aPix[two ? k1 : k0];
            sensePixel(pa);
            if (two) sensePixel(pb);
        }
    }
    if (visited) ctx.atomicAddI(&wk.counters[7], visited);
    ctx.sync();
    long long tc2 = phaseClock();
    // ---- LocalMap.extractOccupancy and BlobFinder both only read the image: no barrier between them.
    // (Measured: labelling the blobs on ONE warp with warp-level barriers while the other warps gather the occupancy grid
    // more than doubles the stage, 6.8 -> 15.7 ms per 65 536 robots -- the passes are dependent shared-memory chains, and a
    // single warp has nothing to hide their latency with while the CTA keeps its 56 KB of shared memory.)
    senseOccupancy(ctx, T, wk, occOut);
    senseBlobs(ctx, T, wk, nRect);
    ctx.sync();
    long long tq0 = tc2;
    int nBlobs = wk.counters[0];
    if (nBlobs > kMaxBlobs) nBlobs = kMaxBlobs;
    long long tc3 = phaseClock();
    // ---- blob table for the local-map stage (one record per blob, in root order; sorted by (colour, seed) there)
    PS_FOR(b, nBlobs) {
        BlobRec r;
        r.x0 = (uint8_t)wk.blobX0[b];
        r.x1 = (uint8_t)wk.blobX1[b];
        r.y0 = (uint8_t)wk.blobY0[b];
        r.y1 = (uint8_t)wk.blobY1[b];
        r.area = (uint16_t)wk.blobArea[b];
        r.seed = wk.blobSeed[b] == (1 << 30) ? (uint16_t)0xFFFF : (uint16_t)wk.blobSeed[b];
        r.colour = wk.img[wk.blobRoot[b]];
        r.pad0 = r.pad1 = r.pad2 = 0;
        out.blobs[b] = r;
    }
    ctx.sync();
    if (ctx.tid() == 0) {
        out.head->nBlobs = nBlobs;
        out.head->leftSum = wk.counters[1];
        out.head->rightSum = wk.counters[2];
        out.head->overflow = wk.counters[5];
        long long tc4 = phaseClock();
        wk.phase[0] = (int)((tc1 - tc0) >> 4);   // candidates
        wk.phase[1] = (int)((tc2 - tc1) >> 4);   // camera
        wk.phase[2] = (int)((tc3 - tc2) >> 4);   // occupancy + blob labelling
        wk.phase[3] = (int)((tc4 - tc3) >> 4);   // blob table
        wk.phase[3] = (int)((tq0 - tc2) >> 4);   // occupancy
        wk.phase[4] = wk.counters[7];            // candidate bodies examined, summed over the pixels
        wk.phase[5] = wk.counters[6];            // puck-coloured pixels
        wk.phase[6] = nCand;
        wk.phase[7] = nBlobs;
    }
    ctx.sync();
}

// ---------------------------------------------------------------------------------------------------------------
// LocalMap.update after the blobs are known: extractPucks (:236-317), extractClusters (:326-370), carriedCluster,
// filterClusters (:375-437). Short, branchy and sequential -- executed by ONE thread per robot, all robots of the
// batch in parallel (the working arrays are thread-local).
// Part A (sequential): perceived pucks, carry flag, clusters. Part B (data-parallel over occupancy cells, `first` /
// `stride` select this caller's share): clusters near ROBOT cells. Part C (sequential): the remaining cluster filters.
PS_HD void localMapPucks(const SenseTables& T, const SenseCfg& sc, const SenseHead& head, const BlobRec* blobs, ps_robot_state* rstate,
                         ps_localmap& lm) {
    int nBlobs = head.nBlobs;
    if (nBlobs > kMaxBlobs) nBlobs = kMaxBlobs;
    // blob list order: by colour, then by seed (scan order); blobs that never seed are not found at all
    uint8_t order[kMaxBlobs];
    int16_t calib[kMaxBlobs];     // calibration row of the blob centre, -1 = uncalibrated
    uint8_t alive[kMaxBlobs];
    int nb = 0;
    for (int s = 0; s < nBlobs; ++s)
        if (blobs[s].seed != 0xFFFF) order[nb++] = (uint8_t)s;
    for (int i = 1; i < nb; ++i) {
        int v = order[i];
        int kv = blobs[v].colour, sv = blobs[v].seed;
        int j = i - 1;
        while (j >= 0) {
            int u = order[j];
            int ku = blobs[u].colour, su = blobs[u].seed;
            if (ku < kv || (ku == kv && su < sv)) break;
            order[j + 1] = (uint8_t)u;
            --j;
        }
        order[j + 1] = (uint8_t)v;
    }
    bool lastCarrying = rstate->lm_carrying != 0;
    bool carrying = false;
    int carriedType = PS_T_NOTHING;
    float carriedX = 0, carriedY = 0;
    int nP = 0;
    int overflow = head.overflow;
    int bi = 0;
    for (int k = 0; k < PS_MAX_COLOURS; ++k) {
        int lo = bi;
        while (bi < nb && blobs[order[bi]].colour == k) ++bi;
        int hi = bi;
        if (lo == hi) continue;
        int holeBlob = -1;
        for (int q = lo; q < hi; ++q) {
            int s = order[q];
            int cx = ((int)blobs[s].x0 + (int)blobs[s].x1) / 2, cy = ((int)blobs[s].y0 + (int)blobs[s].y1) / 2;
            int ci = T.calibIdx[cx * T.imgH + cy];
            calib[s] = (int16_t)ci;
            alive[s] = ci >= 0 ? 1 : 0;   // blobs centred outside the calibrated area are dropped
            if (ci < 0) continue;
            if (T.cFlags[ci] & 2)
                if (holeBlob < 0 || blobs[s].area > blobs[holeBlob].area) holeBlob = s;
        }
        for (int q = lo; q < hi; ++q) {
            int s = order[q];
            if (!alive[s] || s == holeBlob) continue;
            if (T.cFlags[calib[s]] & 2) alive[s] = 0;
        }
        if (holeBlob >= 0) {
            int hc = calib[holeBlob];
            float hx = T.cXr[hc], hy = T.cYr[hc];
            const float INNER_WIDTH = 2 / 3.0f * 6.3f;
            for (int q = lo; q < hi; ++q) {
                int s = order[q];
                if (!alive[s] || s == holeBlob) continue;
                float dx = T.cXr[calib[s]] - hx, dy = T.cYr[calib[s]] - hy;
                if (sqrt((double)(dx * dx + dy * dy)) < (double)INNER_WIDTH) alive[s] = 0;
            }
            float holeFraction = blobs[holeBlob].area / (float)T.nHolePixels;
            if ((lastCarrying && holeFraction > sc.carryLo) || (!carrying && holeFraction > sc.carryHi)) {
                carrying = true;
                carriedType = k;
                carriedX = hx;
                carriedY = hy;
            }
        }
        for (int q = lo; q < hi; ++q) {
            int s = order[q];
            if (!alive[s]) continue;
            if (nP >= PS_MAX_LM_PUCKS) {
                overflow = 1;
                break;
            }
            lm.puck_x[nP] = T.cXr[calib[s]];
            lm.puck_y[nP] = T.cYr[calib[s]];
            lm.puck_type[nP] = (uint8_t)k;
            ++nP;
        }
    }
    lm.n_pucks = nP;
    lm.carrying = carrying ? 1 : 0;
    lm.carried_type = carriedType;
    lm.carried_x = carrying ? carriedX : 0.0f;
    lm.carried_y = carrying ? carriedY : 0.0f;
    rstate->lm_carrying = carrying ? 1 : 0;
    lm.overflow = overflow;
}

// LocalMap.extractClusters per colour: graph over distinct positions, edge if distance < threshold, connected components
// in first-vertex order, members in insertion order; then carriedCluster. Sequential version (one thread).
PS_HD void localMapClusters(const SenseCfg& sc, ps_localmap& lm) {
    int nP = lm.n_pucks;
    bool carrying = lm.carrying != 0;
    float carriedX = lm.carried_x, carriedY = lm.carried_y;
    // rep[p] = first puck of the same colour at the bitwise-same position (the graph vertex p stands for)
    uint8_t rep[PS_MAX_LM_PUCKS];
    for (int p = 0; p < nP; ++p) {
        rep[p] = (uint8_t)p;
        for (int q = 0; q < p; ++q)
            if (lm.puck_type[q] == lm.puck_type[p] && lm.puck_x[q] == lm.puck_x[p] && lm.puck_y[q] == lm.puck_y[p]) {
                rep[p] = (uint8_t)q;
                break;
            }
        lm.puck_cluster[p] = 255;
        lm.puck_degree[p] = 0;
    }
    for (int p = 0; p < nP; ++p) {
        if (rep[p] != p) continue;
        int deg = 0;
        for (int q = 0; q < nP; ++q) {
            if (q == p || rep[q] != q || lm.puck_type[q] != lm.puck_type[p]) continue;
            if (dist(mk(lm.puck_x[p], lm.puck_y[p]), mk(lm.puck_x[q], lm.puck_y[q])) < sc.clusterThreshold) ++deg;
        }
        lm.puck_degree[p] = (uint8_t)(deg > 250 ? 250 : deg);
    }
    int nC = 0;
    for (int p = 0; p < nP; ++p) {
        if (rep[p] != p || lm.puck_cluster[p] != 255) continue;
        int cid = nC++;
        lm.puck_cluster[p] = (uint8_t)cid;
        // grow the component until no vertex is added (discovery order is irrelevant: members are listed in insertion order)
        bool grew = true;
        while (grew) {
            grew = false;
            for (int a = 0; a < nP; ++a) {
                if (lm.puck_cluster[a] != cid || rep[a] != a) continue;
                for (int q = 0; q < nP; ++q) {
                    if (lm.puck_cluster[q] != 255 || rep[q] != q || lm.puck_type[q] != lm.puck_type[a]) continue;
                    if (dist(mk(lm.puck_x[a], lm.puck_y[a]), mk(lm.puck_x[q], lm.puck_y[q])) < sc.clusterThreshold) {
                        lm.puck_cluster[q] = (uint8_t)cid;
                        grew = true;
                    }
                }
            }
        }
        float sx = 0.0f, sy = 0.0f;
        int size = 0;
        for (int a = 0; a < nP; ++a)
            if (lm.puck_cluster[a] == cid && rep[a] == a) {
                sx += lm.puck_x[a];
                sy += lm.puck_y[a];
                ++size;
            }
        float inv = 1.0f / size;
        lm.cluster_x[cid] = sx * inv;
        lm.cluster_y[cid] = sy * inv;
        lm.cluster_type[cid] = lm.puck_type[p];
        lm.cluster_size[cid] = (uint8_t)size;
        lm.cluster_kept[cid] = 1;
    }
    // duplicates share the cluster and degree of their first occurrence
    for (int p = 0; p < nP; ++p)
        if (rep[p] != p) {
            lm.puck_cluster[p] = lm.puck_cluster[rep[p]];
            lm.puck_degree[p] = lm.puck_degree[rep[p]];
        }
    lm.n_clusters = nC;
    // carriedCluster: the last cluster (any colour) holding a vertex equal to carriedV
    int carriedCluster = -1;
    if (carrying)
        for (int p = 0; p < nP; ++p)
            if (lm.puck_x[p] == carriedX && lm.puck_y[p] == carriedY && (int)lm.puck_cluster[p] > carriedCluster) carriedCluster = lm.puck_cluster[p];
    lm.carried_cluster = carriedCluster;
}

PS_HD void localMapPucksAndClusters(const SenseTables& T, const SenseCfg& sc, const SenseHead& head, const BlobRec* blobs,
                                    ps_robot_state* rstate, ps_localmap& lm) {
    localMapPucks(T, sc, head, blobs, rstate, lm);
    localMapClusters(sc, lm);
}

#if defined(__CUDACC__)
// The same by one warp for up to 32 perceived pucks: lane p owns puck p; adjacency as one 32-bit mask per lane, components
// by OR-reduction over the member lanes, cluster ids in first-vertex order, centroids summed by lane 0 in insertion order
// (same float order as the sequential version). Every lane of the warp must call it; lm is in shared memory.
__device__ __forceinline__ void localMapClustersWarp(const SenseCfg& sc, ps_localmap& lm, int lane) {
    const unsigned FULL = 0xFFFFFFFFu;
    int nP = lm.n_pucks;
    bool valid = lane < nP;
    float x = valid ? lm.puck_x[lane] : 0.0f, y = valid ? lm.puck_y[lane] : 0.0f;
    int type = valid ? (int)lm.puck_type[lane] : -1;
    int rep = lane;
    if (valid)
        for (int q = 0; q < lane; ++q)
            if ((int)lm.puck_type[q] == type && lm.puck_x[q] == x && lm.puck_y[q] == y) {
                rep = q;
                break;
            }
    bool isRep = valid && rep == lane;
    unsigned repMask = __ballot_sync(FULL, isRep);
    unsigned adj = 0;
    if (isRep)
        for (int q = 0; q < nP; ++q) {
            if (q == lane || !((repMask >> q) & 1u) || (int)lm.puck_type[q] != type) continue;
            if (dist(mk(x, y), mk(lm.puck_x[q], lm.puck_y[q])) < sc.clusterThreshold) adj |= 1u << q;
        }
    int deg = __popc(adj);
    int cluster = 255;
    int nC = 0;
    unsigned remaining = repMask;
    while (remaining) {
        int seed = __ffs((int)remaining) - 1;
        unsigned comp = 1u << seed;
        for (;;) {
            unsigned grown = comp | __reduce_or_sync(FULL, ((comp >> lane) & 1u) ? adj : 0u);
            if (grown == comp) break;
            comp = grown;
        }
        if ((comp >> lane) & 1u) cluster = nC;
        if (lane == 0) {
            float sx = 0.0f, sy = 0.0f;
            int size = 0;
            for (unsigned m = comp; m; m &= m - 1u) {
                int a = __ffs((int)m) - 1;
                sx += lm.puck_x[a];
                sy += lm.puck_y[a];
                ++size;
            }
            float inv = 1.0f / size;
            lm.cluster_x[nC] = sx * inv;
            lm.cluster_y[nC] = sy * inv;
            lm.cluster_type[nC] = lm.puck_type[seed];
            lm.cluster_size[nC] = (uint8_t)size;
            lm.cluster_kept[nC] = 1;
        }
        remaining &= ~comp;
        ++nC;
    }
    // duplicates share the cluster and degree of their first occurrence
    int repCluster = __shfl_sync(FULL, cluster, rep & 31);
    int repDeg = __shfl_sync(FULL, deg, rep & 31);
    if (valid) {
        lm.puck_cluster[lane] = (uint8_t)(isRep ? cluster : repCluster);
        lm.puck_degree[lane] = (uint8_t)(isRep ? deg : repDeg);
    }
    // carriedCluster: the last cluster (any colour) holding a vertex equal to carriedV
    int mine = -1;
    if (lm.carrying != 0 && valid && x == lm.carried_x && y == lm.carried_y) mine = isRep ? cluster : repCluster;
    for (int o = 16; o > 0; o >>= 1) mine = max(mine, __shfl_xor_sync(FULL, mine, o));
    if (lane == 0) {
        lm.n_clusters = nC;
        lm.carried_cluster = mine;
    }
}
#endif

// LocalMap.filterClusters, part 1: clusters with a puck closer than 10 cm to a ROBOT cell. Handles the 4-cell words
// first, first + stride, ... of the occupancy grid (the writes are idempotent, so callers may share the work freely).
PS_HD void localMapRobotFilter(const SenseTables& T, const uint8_t* occ, ps_localmap& lm, int first, int stride) {
    int nP = lm.n_pucks;
    if (nP == 0) return;
    int nCells = T.occW * T.occH;
    bool aligned = (((size_t)occ) & 3) == 0;
    int nWords = (nCells + 3) / 4;
    for (int w = first; w < nWords; w += stride) {
        int c0 = w * 4;
        if (aligned && c0 + 4 <= nCells) {
            // most words hold no ROBOT cell at all
            uint32_t wv = *(const uint32_t*)(occ + c0);
            uint32_t xr = wv ^ (0x01010101u * (uint32_t)PS_T_ROBOT);
            if (((xr - 0x01010101u) & ~xr & 0x80808080u) == 0) continue;
        }
        for (int c = c0; c < c0 + 4 && c < nCells; ++c) {
            if (occ[c] != PS_T_ROBOT) continue;
            int i = c / T.occH, j = c - i * T.occH;
            V2 robotV = mk(1.0f * j, T.maxYr - 1.0f * i);
            for (int p = 0; p < nP; ++p)
                if (distSq(robotV, mk(lm.puck_x[p], lm.puck_y[p])) < 100.0f) lm.cluster_kept[lm.puck_cluster[p]] = 0;
        }
    }
}

// LocalMap.filterClusters, part 2: sequential pass over the surviving clusters
PS_HD void localMapFinalFilter(const SenseCfg& sc, const SenseHead& head, ps_localmap& lm) {
    int nP = lm.n_pucks, nC = lm.n_clusters;
    bool carrying = lm.carrying != 0;
    for (int c = 0; c < nC; ++c) {
        if (!lm.cluster_kept[c]) continue;
        bool remove = false;
        if (carrying && c == lm.carried_cluster) remove = true;
        else if (!lmReachable(lm.cluster_x[c], lm.cluster_y[c])) remove = true;
        else if (carrying && lm.carried_type != lm.cluster_type[c]) remove = true;
        else if (carrying) {
            for (int o = 0; o < nC && !remove; ++o) {
                if (o == c || !lm.cluster_kept[o] || lm.cluster_type[o] == lm.cluster_type[c]) continue;
                for (int p = 0; p < nP && !remove; ++p) {
                    if (lm.puck_cluster[p] != c) continue;
                    for (int q = 0; q < nP; ++q)
                        if (lm.puck_cluster[q] == o &&
                            dist(mk(lm.puck_x[p], lm.puck_y[p]), mk(lm.puck_x[q], lm.puck_y[q])) < sc.clusterThreshold) {
                            remove = true;
                            break;
                        }
                }
            }
        }
        if (remove) lm.cluster_kept[c] = 0;
    }
    lm.left_sum = head.leftSum;
    lm.right_sum = head.rightSum;
}

#if defined(__CUDACC__)
// localMapFinalFilter by one warp. The clusters are still visited one after the other (a cluster removed earlier no longer
// counts as a neighbour of the later ones), but "some puck of c lies within the threshold of some puck of a kept cluster
// of another colour" is an existence test over puck pairs, so the lanes share the pairs and vote.
__device__ __forceinline__ void localMapFinalFilterWarp(const SenseCfg& sc, const SenseHead& head, ps_localmap& lm, int lane) {
    const unsigned FULL = 0xFFFFFFFFu;
    int nP = lm.n_pucks, nC = lm.n_clusters;
    bool carrying = lm.carrying != 0;
    for (int c = 0; c < nC; ++c) {
        if (!lm.cluster_kept[c]) continue;   // shared memory, same for every lane
        bool remove = false;
        if (carrying && c == lm.carried_cluster) remove = true;
        else if (!lmReachable(lm.cluster_x[c], lm.cluster_y[c])) remove = true;
        else if (carrying && lm.carried_type != lm.cluster_type[c]) remove = true;
        else if (carrying) {
            bool found = false;
            int typeC = lm.cluster_type[c];
            for (int idx = lane; idx < nP * nP && !found; idx += 32) {
                int pi = idx / nP, q = idx - pi * nP;
                if (lm.puck_cluster[pi] != c) continue;
                int o = lm.puck_cluster[q];
                if (o == c || !lm.cluster_kept[o] || lm.cluster_type[o] == typeC) continue;
                if (dist(mk(lm.puck_x[pi], lm.puck_y[pi]), mk(lm.puck_x[q], lm.puck_y[q])) < sc.clusterThreshold) found = true;
            }
            remove = __any_sync(FULL, found);
        }
        __syncwarp();
        if (remove && lane == 0) lm.cluster_kept[c] = 0;
        __syncwarp();
    }
    if (lane == 0) {
        lm.left_sum = head.leftSum;
        lm.right_sum = head.rightSum;
    }
}
#endif

PS_HD void localMapFromBlobs(const SenseTables& T, const SenseCfg& sc, const SenseHead& head, const BlobRec* blobs, const uint8_t* occ,
                             ps_robot_state* rstate, ps_localmap& lm) {
    localMapPucksAndClusters(T, sc, head, blobs, rstate, lm);
    localMapRobotFilter(T, occ, lm, 0, 1);
    localMapFinalFilter(sc, head, lm);
}

// ---------------------------------------------------------------------------------------------------------------
// Thinking. One CTA per arena; the robots of an arena are visited in creation order (Arena.java:197).

PS_HD double constrainAngle(double angle) {   // AngleUtils.constrainAngle
    const double PI = 3.14159265358979323846;
    while (angle > PI) angle -= 2 * PI;
    while (angle <= -PI) angle += 2 * PI;
    return angle;
}
PS_HD double angularDifference(double a, double b) {   // AngleUtils.getAngularDifference
    const double PI = 3.14159265358979323846;
    a = constrainAngle(a);
    b = constrainAngle(b);
    double error = fabs(a - b);
    if (error > PI) {
        error -= PI * 2;
        error = fabs(error);
    }
    return error;
}
PS_HD float satf(float x, float a, float b) { return x <= a ? a : (x >= b ? b : x); }
// MovementCommand (src/localmap/MovementCommand.java:20-23) over MovementUtils.getSpeedScale / getTorqueScale
PS_HD void movementCommand(float forwardSpeed, float turnAngle, float& forwards, float& torque) {
    const float PI_OVER_4f = (float)(3.14159265358979323846 / 4.0);
    const float MAX_FORWARDS = 5 * 150000, MAX_TORQUE = 5 * 475000;
    float speedScale = satf((PI_OVER_4f - fabsf(turnAngle)) / PI_OVER_4f, 0.25f, 1);
    float torqueScale = satf((float)(2.0f * turnAngle / (0.5 * 3.14159265358979323846)), -1, 1);
    forwards = MAX_FORWARDS * forwardSpeed * speedScale;
    torque = MAX_TORQUE * torqueScale;
}

static constexpr float kVfhStart = (float)(3.14159265358979323846 / 2);
static constexpr float kVfhStop = (float)(-3.14159265358979323846 / 2);
static constexpr float kVfhAlpha = (kVfhStart - kVfhStop) / (kSectors - 1);
PS_HD float vfhIndexToAngle(int k) { return kVfhStart - k * kVfhAlpha; }
PS_HD int javaIntCast(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
PS_HD int vfhAngleToIndex(float angle) { return javaIntCast((kVfhStart - angle) / kVfhAlpha); }

struct ThinkWork {
    unsigned* primary;    // [kSectors] float bits (magnitudes are >= 0, so unsigned order == float order)
    unsigned* limits;     // [2]: left limit (positive float bits, min), right limit (negative float bits: min bits == max value)
    int* flag;            // [4]
};
PS_HD ThinkWork carveThink(Carver& cv) {
    ThinkWork w;
    w.primary = (unsigned*)cv.take(kSectors * sizeof(unsigned));
    w.limits = (unsigned*)cv.take(2 * sizeof(unsigned));
    w.flag = (int*)cv.take(4 * sizeof(int));
    return w;
}
PS_HD unsigned fbits(float f) {
    union {
        float f;
        unsigned u;
    } c;
    c.f = f;
    return c.u;
}
PS_HD float bitsf(unsigned u) {
    union {
        float f;
        unsigned u;
    } c;
    c.u = u;
    return c.f;
}

// VFHPlus.computeTurnAngle. occ may have one puck colour masked out (LocalMap.postFilterOccupancy is applied by the
// caller on the grid itself). Returns false for Java's null. Collective: every thread of the CTA calls it.
template <class Ctx>
PS_HD bool vfhTurnAngle(Ctx& ctx, const SenseTables& T, const SenseCfg& sc, const ThinkWork& tw, const uint8_t* occ, ps_robot_state* rs,
                        float targetAngle, bool ignorePucks, float& result) {
    int nCells = T.occW * T.occH;
    PS_FOR(k, kSectors) tw.primary[k] = 0u;
    if (ctx.tid() == 0) {
        tw.limits[0] = fbits(kVfhStart);
        tw.limits[1] = fbits(kVfhStop);
    }
    ctx.sync();
    PS_FOR(c, nCells) {
        uint8_t t = occ[c];
        if (t == PS_T_NOTHING || (ignorePucks && t < PS_MAX_COLOURS)) continue;
        unsigned mag = fbits(T.vfhMag[c]);
        int k0 = T.vfhStartK[c], k1 = T.vfhStopK[c];
        if (k0 < 0) k0 = 0;
        if (k1 > kSectors - 1) k1 = kSectors - 1;
        for (int k = k0; k <= k1; ++k) ctx.atomicMaxU(&tw.primary[k], mag);
        float b = T.vfhBeta[c];
        uint8_t side = T.vfhSide[c];
        // rightLimit = max over b < 0 (bits of negative floats grow with magnitude: min bits == max value)
        if (b < 0 && (side & 2)) ctx.atomicMinU(&tw.limits[1], fbits(b));
        if (b > 0 && (side & 1)) ctx.atomicMinU(&tw.limits[0], fbits(b));
    }
    ctx.sync();
    bool ok = false;
    if (ctx.tid() == 0) {
        float leftLimit = bitsf(tw.limits[0]), rightLimit = bitsf(tw.limits[1]);
        // binary histogram with hysteresis, then the masked histogram
        unsigned char masked[kSectors];
        for (int k = 0; k < kSectors; ++k) {
            float pv = bitsf(tw.primary[k]);
            unsigned bit = (rs->vfh_binary[k >> 5] >> (k & 31)) & 1u;
            if (pv > sc.tauHi) bit = 1;
            else if (pv < sc.tauLo) bit = 0;
            if (bit) rs->vfh_binary[k >> 5] |= 1u << (k & 31);
            else rs->vfh_binary[k >> 5] &= ~(1u << (k & 31));
            unsigned char m = 0;
            if (bit == 0) {
                float sectorAngle = vfhIndexToAngle(k);
                if (sectorAngle > leftLimit || sectorAngle < rightLimit) m = 1;
            } else
                m = 1;
            masked[k] = m;
        }
        // candidate directions (goalDirected == true for both shipped controllers) and the cheapest one
        int targetSector = vfhAngleToIndex(targetAngle);
        int bestSector = -1;
        float bestCost = 0.0f;
        bool opening = masked[0] == 0;
        int left = 0;
        for (int k = 1; k <= kSectors; ++k) {
            int right = -1;
            if (k < kSectors) {
                if (opening) {
                    if (masked[k - 1] == 0 && masked[k] == 1) {
                        right = k - 1;
                        opening = false;
                    }
                } else if (masked[k - 1] == 1 && masked[k] == 0) {
                    left = k;
                    opening = true;
                }
            } else if (opening && masked[kSectors - 1] == 0) {
                right = kSectors - 1;
            }
            if (right < 0) continue;
            int cand[3], nc = 0;
            int width = right - left + 1;
            if (width < 8) cand[nc++] = (left + right) / 2;
            else {
                cand[nc++] = left + 4;
                cand[nc++] = right - 4;
                if (targetSector >= left && targetSector <= right) cand[nc++] = targetSector;
            }
            for (int i = 0; i < nc; ++i) {
                double sectorAngle = vfhIndexToAngle(cand[i]);
                float cost = (float)(5.0f * angularDifference(sectorAngle, targetAngle) + 2.0f * angularDifference(sectorAngle, 0) +
                                     3.0f * angularDifference(sectorAngle, rs->vfh_last_angle));
                if (bestSector < 0 || cost < bestCost) {
                    bestSector = cand[i];
                    bestCost = cost;
                }
            }
        }
        if (bestSector >= 0) {
            rs->vfh_last_sector = bestSector;
            rs->vfh_last_angle = vfhIndexToAngle(bestSector);
            tw.flag[0] = 1;
        } else {
            tw.flag[0] = 0;
        }
    }
    ctx.sync();
    ok = tw.flag[0] != 0;
    result = rs->vfh_last_angle;
    ctx.sync();
    return ok;
}

// MovementUtils.applyVFH (src/localmap/MovementUtils.java:46-58)
template <class Ctx>
PS_HD void applyVFH(Ctx& ctx, const SenseTables& T, const SenseCfg& sc, const ThinkWork& tw, const uint8_t* occ, const ps_localmap* lm,
                    ps_robot_state* rs, float desiredTurn, bool ignorePucks, float& forwards, float& torque) {
    float result;
    bool ok = vfhTurnAngle(ctx, T, sc, tw, occ, rs, desiredTurn, ignorePucks, result);
    const double PI = 3.14159265358979323846;
    if (!ok) {
        int freeer = lm->left_sum <= lm->right_sum ? 1 : -1;
        movementCommand(0.5f, (float)(0.5f * PI * freeer), forwards, torque);
    } else {
        movementCommand(1.0f, (float)(0.5f * PI * result), forwards, torque);
    }
}

struct PoseD {
    double x, y, theta;
};
PS_HD double poseAngleFromTo(const PoseD& a, double bx, double by) {   // Pose.getAngleFromTo
    return constrainAngle(atan2(by - a.y, bx - a.x) - a.theta);
}
// CacheConsController.getCachePointInRobotCoords (:695-705)
PS_HD V2 cachePointInRobot(const PoseD& pose, const ps_robot_state* rs, int k) {
    double dx = rs->cache_x[k] - pose.x, dy = rs->cache_y[k] - pose.y;
    double dd = sqrt(dx * dx + dy * dy);
    double alpha = poseAngleFromTo(pose, rs->cache_x[k], rs->cache_y[k]);
    return mk((float)(dd * (float)cos(alpha)), (float)(dd * (float)sin(alpha)));
}
PS_HD float depositProb(float f, float K2) {
    float p = (f / (K2 + f));
    return p * p;
}
PS_HD float pickupProb(float f, float K1) {
    float p = (K1 / (K1 + f));
    return p * p;
}

// CacheConsController.computeDesired for one robot; sequential part by thread 0, VFH by the whole CTA.
template <class Ctx>
PS_HD void thinkCacheCons(Ctx& ctx, const SenseTables& T, const SenseCfg& sc, const ThinkWork& tw, uint8_t* occ, ps_localmap* lm,
                          ps_robot_state* rs, JRandom* rng, const float* poseIn, int stepCount, bool preset) {
    // phase 1 (thread 0): caches, filtering, state transition; publishes what the movement phase needs in tw.flag
    //   flag[1] = movement kind: 0 VFH(angle flagf), 1 direct command
    if (ctx.tid() == 0) {
        PoseD pose;
        pose.x = poseIn[0];
        pose.y = poseIn[1];
        pose.theta = poseIn[2];
        bool carrying = lm->carrying != 0;
        int carriedType = lm->carried_type;
        int nP = lm->n_pucks, nC = lm->n_clusters;
        float thr = sc.clusterThreshold;
        // checkForget (:556-566)
        if (sc.cacheSelection == PS_SEL_RELATIVE_FORGET) {
            if (!preset && sc.forgetProb != -1 && jrNextFloat(*rng) < sc.forgetProb)
                for (int k = 0; k < PS_MAX_COLOURS; k++) {
                    rs->cache_valid[k] = 0;
                    rs->cache_x[k] = rs->cache_y[k] = 0.0;
                    rs->remembered[k] = 0;
                }
        }
        // remembered cache sizes (:346-362): the last raw cluster of colour k holding the cache point
        for (int k = 0; k < PS_MAX_COLOURS; k++) {
            if (!rs->cache_valid[k]) continue;
            V2 cp = cachePointInRobot(pose, rs, k);
            int cacheCluster = -1;
            for (int p = 0; p < nP; ++p)
                if (lm->puck_type[p] == k && dist(mk(lm->puck_x[p], lm->puck_y[p]), cp) < thr && (int)lm->puck_cluster[p] > cacheCluster)
                    cacheCluster = lm->puck_cluster[p];
            if (cacheCluster >= 0) {
                float sz = (float)lm->cluster_size[cacheCluster];
                rs->remembered[k] = rs->remembered[k] > sz ? rs->remembered[k] : sz;   // Math.max
            }
        }
        // new cache points (:364-386): the first largest raw cluster of every colour is the candidate
        for (int k = 0; k < PS_MAX_COLOURS; k++) {
            int cand = -1;
            for (int c = 0; c < nC; ++c)
                if (lm->cluster_type[c] == k && (cand < 0 || lm->cluster_size[c] > lm->cluster_size[cand])) cand = c;
            if (cand < 0) continue;
            float csize = (float)lm->cluster_size[cand];
            // checkCacheSelection (:569-590)
            bool sel = false;
            if (sc.cacheSelection == PS_SEL_PROB_ABSOLUTE) sel = jrNextFloat(*rng) < depositProb(csize, sc.kAbsolute);
            else if (sc.cacheSelection == PS_SEL_PROB_RELATIVE) sel = jrNextFloat(*rng) < depositProb(csize - rs->remembered[k], sc.kRelative);
            else sel = csize > rs->remembered[k];
            if (!sel) continue;
            // checkCacheSeparation (:598-647)
            V2 cc = mk(lm->cluster_x[cand], lm->cluster_y[cand]);
            unsigned nearby = 0;
            for (int o = 0; o < PS_MAX_COLOURS; o++)
                if (o != k && rs->cache_valid[o]) {
                    V2 other = cachePointInRobot(pose, rs, o);
                    if (dist(cc, other) < sc.homeSeparation) nearby |= 1u << o;
                }
            bool accept = true;
            if (sc.cacheSeparation == PS_SEP_IGNORE) accept = true;
            else if (sc.cacheSeparation == PS_SEP_NULLIFY_OLD) {
                for (int o = 0; o < PS_MAX_COLOURS; o++)
                    if (nearby & (1u << o)) {
                        rs->cache_valid[o] = 0;
                        rs->cache_x[o] = rs->cache_y[o] = 0.0;
                        rs->remembered[o] = 0;
                    }
            } else if (sc.cacheSeparation == PS_SEP_ACCEPT_ISOLATED) accept = nearby == 0;
            else {
                bool newOneIsLargest = true;
                for (int o = 0; o < PS_MAX_COLOURS; o++)
                    if ((nearby & (1u << o)) && csize <= rs->remembered[o]) newOneIsLargest = false;
                if (newOneIsLargest)
                    for (int o = 0; o < PS_MAX_COLOURS; o++)
                        if (nearby & (1u << o)) {
                            rs->cache_valid[o] = 0;
                        rs->cache_x[o] = rs->cache_y[o] = 0.0;
                            rs->remembered[o] = 0;
                        }
                accept = newOneIsLargest;
            }
            if (accept) {
                // initializeCache (:649-656): Pose.getTranslated(robotPose, centroid)
                double ddist = sqrt((double)(cc.x * cc.x + cc.y * cc.y));
                double angle = atan2((double)cc.y, (double)cc.x) + pose.theta;
                rs->cache_x[k] = pose.x + ddist * cos(angle);
                rs->cache_y[k] = pose.y + ddist * sin(angle);
                rs->cache_valid[k] = 1;
                rs->remembered[k] = csize;
            }
        }
        // localMap.postFilterClusters(k) for colours without a cache (:388-391)
        for (int c = 0; c < nC; ++c)
            if (!rs->cache_valid[lm->cluster_type[c]]) lm->cluster_kept[c] = 0;

        int state = rs->ctrl_state;
        int startCount = rs->start_count;
#define TRANSITION(s)        \
    do {                     \
        state = (s);         \
        startCount = stepCount; \
    } while (0)
        if (carrying && !rs->cache_valid[carriedType]) {
            TRANSITION(PS_CC_SPIT_OUT);
        } else {
            switch (state) {
                case PS_CC_PU_SCAN:
                    if (carrying) TRANSITION(PS_CC_HOMING);
                    else {
                        // ClusterTargetSelector.selectTarget(localMap, false, true) over ExtremaSelector.selectCluster
                        int hasTarget = 0;
                        float tx = 0, ty = 0;
                        int ttype = 0;
                        int extreme = -1;
                        for (int c = 0; c < nC; ++c) {
                            if (!lm->cluster_kept[c]) continue;
                            if (extreme < 0) {
                                extreme = c;
                                // (the loop in the reference re-visits the first element; that is a no-op)
                                continue;
                            }
                            if (lm->cluster_size[c] < lm->cluster_size[extreme]) extreme = c;
                            else if (lm->cluster_size[c] == lm->cluster_size[extreme] && fabsf(lm->cluster_y[c]) < fabsf(lm->cluster_y[extreme])) extreme = c;
                        }
                        if (extreme >= 0) {
                            float prob = sc.k1 < 0 ? 1.0f : pickupProb((float)lm->cluster_size[extreme], sc.k1);
                            if (jrNextFloat(*rng) < prob) {
                                // selectPickupPuckCluster: extreme-Yr member of lower degree, if reachable
                                float largestYr = -INFINITY, smallestYr = INFINITY;
                                int li = -1, si = -1;
                                for (int p = 0; p < nP; ++p) {
                                    if (lm->puck_cluster[p] != extreme) continue;
                                    // skip repeated vertices (the subgraph holds each position once)
                                    bool dup = false;
                                    for (int q = 0; q < p; ++q)
                                        if (lm->puck_cluster[q] == extreme && lm->puck_x[q] == lm->puck_x[p] && lm->puck_y[q] == lm->puck_y[p]) dup = true;
                                    if (dup) continue;
                                    if (lm->puck_y[p] > largestYr) {
                                        largestYr = lm->puck_y[p];
                                        li = p;
                                    }
                                    if (lm->puck_y[p] < smallestYr) {
                                        smallestYr = lm->puck_y[p];
                                        si = p;
                                    }
                                }
                                int pick = lm->puck_degree[li] <= lm->puck_degree[si] ? li : si;
                                if (lmReachable(lm->puck_x[pick], lm->puck_y[pick])) {
                                    hasTarget = 1;
                                    tx = lm->puck_x[pick];
                                    ty = lm->puck_y[pick];
                                    ttype = lm->cluster_type[extreme];
                                }
                            }
                        }
                        if (hasTarget) {
                            if (!rs->cache_valid[ttype]) hasTarget = 0;
                            else {
                                // cachePointInCluster(target): target is a one-puck cluster
                                V2 cp = cachePointInRobot(pose, rs, ttype);
                                if (dist(mk(tx, ty), cp) < thr) hasTarget = 0;
                                else TRANSITION(PS_CC_PU_TARGET);
                            }
                        }
                        rs->has_target = hasTarget;
                        rs->target_x = hasTarget ? tx : 0.0f;
                        rs->target_y = hasTarget ? ty : 0.0f;
                        rs->target_type = hasTarget ? ttype : 0;
                    }
                    break;
                case PS_CC_PU_TARGET:
                    if (carrying) TRANSITION(PS_CC_HOMING);
                    else {
                        // ProbSeekController.matchTarget with carrying == false: closest puck of the kept clusters
                        int have = 0;
                        float bx = 0, by = 0, smallest = INFINITY;
                        int btype = -1;
                        V2 last = mk(rs->target_x, rs->target_y);
                        for (int c = 0; c < nC; ++c) {
                            if (!lm->cluster_kept[c]) continue;
                            for (int p = 0; p < nP; ++p) {
                                if (lm->puck_cluster[p] != c) continue;
                                float ds = distSq(last, mk(lm->puck_x[p], lm->puck_y[p]));
                                if (ds < sc.predictionDistSqd && ds < smallest) {
                                    smallest = ds;
                                    bx = lm->puck_x[p];
                                    by = lm->puck_y[p];
                                    btype = lm->cluster_type[c];
                                    have = 1;
                                }
                            }
                        }
                        if (have && !lmReachable(bx, by)) have = 0;
                        rs->has_target = have;
                        rs->target_x = have ? bx : 0.0f;
                        rs->target_y = have ? by : 0.0f;
                        rs->target_type = have ? btype : 0;
                        if (!have) TRANSITION(PS_CC_PU_SCAN);
                        else if (!rs->cache_valid[btype]) TRANSITION(PS_CC_PU_SCAN);
                        else if (dist(mk(bx, by), cachePointInRobot(pose, rs, btype)) < thr) TRANSITION(PS_CC_PU_SCAN);
                        else if (stepCount - startCount > sc.pickUpTime) TRANSITION(PS_CC_PU_SCAN);
                    }
                    break;
                case PS_CC_HOMING:
                    if (!carrying) TRANSITION(PS_CC_PU_SCAN);
                    else if (!rs->cache_valid[carriedType]) TRANSITION(PS_CC_PU_SCAN);
                    else {
                        // carriedPuckInCacheCluster (:674-693)
                        bool in = false;
                        if (lm->carried_cluster >= 0) {
                            V2 cp = cachePointInRobot(pose, rs, carriedType);
                            for (int p = 0; p < nP && !in; ++p)
                                if (lm->puck_cluster[p] == lm->carried_cluster && dist(mk(lm->puck_x[p], lm->puck_y[p]), cp) < thr) in = true;
                        }
                        if (in) TRANSITION(PS_CC_DE_PUSH);
                    }
                    break;
                case PS_CC_DE_PUSH:
                    if (stepCount - startCount >= sc.pushInTime) TRANSITION(PS_CC_DE_BACKUP);
                    break;
                case PS_CC_DE_BACKUP:
                    if (stepCount - startCount >= sc.backupTime) {
                        if (rs->last_carried_type < 0 || !rs->cache_valid[rs->last_carried_type]) TRANSITION(PS_CC_PU_SCAN);
                        else TRANSITION(PS_CC_EXILE);
                    }
                    break;
                case PS_CC_EXILE:
                    if (stepCount - startCount >= sc.exileTime) TRANSITION(PS_CC_PU_SCAN);
                    else if (rs->last_carried_type < 0 || !rs->cache_valid[rs->last_carried_type]) TRANSITION(PS_CC_PU_SCAN);
                    break;
                case PS_CC_SPIT_OUT:
                    if (stepCount - startCount >= sc.spitOutTime) TRANSITION(PS_CC_PU_SCAN);
                    break;
            }
        }
#undef TRANSITION
        rs->ctrl_state = state;
        rs->start_count = startCount;
        // movement (:480-526)
        int useVfh = 0;
        float angle = 0.0f, fwd = 0.0f, tq = 0.0f;
        int filterType = -1;
        if (state == PS_CC_PU_SCAN) {
            angle = sc.stDev * (float)jrNextGaussian(*rng);
            useVfh = 1;
        } else if (state == PS_CC_PU_TARGET) {
            movementCommand(1.0f, (float)atan2((double)rs->target_y, (double)rs->target_x), fwd, tq);
        } else if (state == PS_CC_HOMING) {
            angle = (float)poseAngleFromTo(pose, rs->cache_x[carriedType], rs->cache_y[carriedType]);
            filterType = carriedType;
            useVfh = 1;
        } else if (state == PS_CC_DE_PUSH) {
            movementCommand(1.0f, 0.0f, fwd, tq);
        } else if (state == PS_CC_DE_BACKUP) {
            movementCommand(-1.0f, 0.0f, fwd, tq);
        } else if (state == PS_CC_EXILE) {
            int lc = rs->last_carried_type;
            angle = (float)poseAngleFromTo(pose, rs->cache_x[lc], rs->cache_y[lc]);
            angle = (float)constrainAngle(angle + 3.14159265358979323846);
            useVfh = 1;
        } else if (state == PS_CC_SPIT_OUT) {
            float randomTurn = sc.stDev * (float)jrNextGaussian(*rng);
            movementCommand(-1.0f, randomTurn, fwd, tq);
        }
        if (carrying) rs->last_carried_type = carriedType;
        rs->cmd_forwards = fwd;
        rs->cmd_torque = tq;
        tw.flag[1] = useVfh;
        tw.flag[2] = filterType;
        tw.limits[0] = fbits(angle);   // handed to the movement phase below (vfhTurnAngle resets limits afterwards)
    }
    ctx.sync();
    int useVfh = tw.flag[1], filterType = tw.flag[2];
    float angle = bitsf(tw.limits[0]);
    ctx.sync();
    if (useVfh) {
        if (filterType >= 0) {
            // LocalMap.postFilterOccupancy(carriedType)
            PS_FOR(c, T.occW * T.occH) if (occ[c] == filterType) occ[c] = PS_T_NOTHING;
            ctx.sync();
        }
        float fwd, tq;
        bool ignore = preset ? false : sc.ignorePucks != 0;
        applyVFH(ctx, T, sc, tw, occ, lm, rs, angle, ignore, fwd, tq);
        if (ctx.tid() == 0) {
            rs->cmd_forwards = fwd;
            rs->cmd_torque = tq;
        }
    }
    if (ctx.tid() == 0) {
        // updateStateCounts (:313-333), file output omitted
        rs->state_counts[rs->ctrl_state & 7]++;
        if (stepCount % 100 == 0)
            for (int i = 0; i < 8; i++) rs->state_counts[i] = 0;
    }
    ctx.sync();
}

// BHDController.computeDesired (src/controllers/BHDController.java:126-191)
template <class Ctx>
PS_HD void thinkBHD(Ctx& ctx, const SenseTables& T, const SenseCfg& sc, const ThinkWork& tw, uint8_t* occ, ps_localmap* lm, ps_robot_state* rs,
                    JRandom* rng, int stepCount) {
    int state = rs->ctrl_state;
    ctx.sync();
    bool turn = false;
    if (state == PS_BHD_FORWARD) {
        bool push = lm->carried_cluster >= 0 && lm->cluster_size[lm->carried_cluster] > 1;
        if (!push) {
            float result;
            bool ok = vfhTurnAngle(ctx, T, sc, tw, occ, rs, 0.0f, true, result);
            if (!ok || fabs((double)result) > 3.14159265358979323846 / 4) turn = true;
        }
        if (ctx.tid() == 0) {
            if (push) {
                rs->ctrl_state = PS_BHD_PUSH;
                rs->start_count = stepCount;
            }
        }
    }
    ctx.sync();
    if (ctx.tid() == 0) {
        const int TURNTIME_MIN = 10 / 2, ONE_EIGHTY_TIME = 10;
        bool initTurn = false;
        if (state == PS_BHD_FORWARD) {
            if (turn) {
                rs->ctrl_state = PS_BHD_TURN;
                rs->start_count = stepCount;
                initTurn = true;
            }
        } else if (state == PS_BHD_PUSH) {
            if (stepCount - rs->start_count >= sc.bhdPushInTime) {
                rs->ctrl_state = PS_BHD_BACKUP;
                rs->start_count = stepCount;
            }
        } else if (state == PS_BHD_BACKUP) {
            if (stepCount - rs->start_count >= sc.bhdBackupTime) {
                rs->ctrl_state = PS_BHD_TURN;
                rs->start_count = stepCount;
                initTurn = true;
            }
        } else if (state == PS_BHD_TURN) {
            if (stepCount - rs->start_count >= rs->bhd_turn_time) {
                rs->ctrl_state = PS_BHD_FORWARD;
                rs->start_count = stepCount;
            }
        }
        if (initTurn) {
            rs->bhd_turn_dir = lm->left_sum <= lm->right_sum ? 1 : -1;
            rs->bhd_turn_time = TURNTIME_MIN + jrNextInt(*rng, ONE_EIGHTY_TIME - TURNTIME_MIN + 1);
        }
        float fwd = 0, tq = 0;
        int s = rs->ctrl_state;
        if (s == PS_BHD_FORWARD || s == PS_BHD_PUSH) movementCommand(1.0f, 0.0f, fwd, tq);
        else if (s == PS_BHD_BACKUP) movementCommand(-1.0f, 0.0f, fwd, tq);
        else movementCommand(0.0f, (float)(3.14159265358979323846 / 2 * rs->bhd_turn_dir), fwd, tq);
        rs->cmd_forwards = fwd;
        rs->cmd_torque = tq;
        rs->state_counts[s & 7]++;
        if (stepCount % 100 == 0)
            for (int i = 0; i < 8; i++) rs->state_counts[i] = 0;
    }
    ctx.sync();
}

// ProbSeekController.computeDesired (src/controllers/ProbSeekController.java:176-279): the SI-2012 benchmark controller.
// Same selector (ExtremaSelector over ClusterTargetSelector), matchTarget and VFH+ as CacheCons; no caches.
// The ProbSeekController.* property keys travel in the cc_* fields of ps_config (same names, same defaults) plus
// ps_placement_time; ALWAYS_TARGET_ISOLATED is false (the reference reads it from the K2 key as a boolean, :107).
template <class Ctx>
PS_HD void thinkProbSeek(Ctx& ctx, const SenseTables& T, const SenseCfg& sc, const ThinkWork& tw, uint8_t* occ, ps_localmap* lm, ps_robot_state* rs,
                         JRandom* rng, int stepCount) {
    ctx.sync();
    if (ctx.tid() == 0) {
        int nP = lm->n_pucks, nC = lm->n_clusters;
        bool carrying = lm->carrying != 0;
        int state = rs->ctrl_state;
        int startCount = rs->start_count;
#define TRANSITION(s)           \
    do {                        \
        state = (s);            \
        startCount = stepCount; \
    } while (0)
        // ClusterTargetSelector.selectTarget(carrying, !carrying): ExtremaSelector.selectCluster, then (pick-up only) the
        // extreme-Yr member of lower degree, if reachable (ClusterTargetSelector.java:48-101, ExtremaSelector.java:36-59)
        auto selectTarget = [&](bool preferLargest, bool isolatePuck) {
            int hasTarget = 0;
            float tx = 0, ty = 0;
            int ttype = 0;
            int extreme = -1;
            for (int c = 0; c < nC; ++c) {
                if (!lm->cluster_kept[c]) continue;
                if (extreme < 0) {
                    extreme = c;
                    continue;
                }
                if (preferLargest && lm->cluster_size[c] > lm->cluster_size[extreme]) extreme = c;
                else if (!preferLargest && lm->cluster_size[c] < lm->cluster_size[extreme]) extreme = c;
                else if (lm->cluster_size[c] == lm->cluster_size[extreme] && fabsf(lm->cluster_y[c]) < fabsf(lm->cluster_y[extreme])) extreme = c;
            }
            if (extreme >= 0) {
                float prob;
                if (preferLargest) prob = sc.k2 < 0 ? 1.0f : depositProb((float)lm->cluster_size[extreme], sc.k2);
                else prob = sc.k1 < 0 ? 1.0f : pickupProb((float)lm->cluster_size[extreme], sc.k1);
                if (jrNextFloat(*rng) < prob) {
                    if (!isolatePuck) {
                        hasTarget = 1;
                        tx = lm->cluster_x[extreme];
                        ty = lm->cluster_y[extreme];
                        ttype = lm->cluster_type[extreme];
                    } else {
                        float largestYr = -INFINITY, smallestYr = INFINITY;
                        int li = -1, si = -1;
                        for (int p = 0; p < nP; ++p) {
                            if (lm->puck_cluster[p] != extreme) continue;
                            bool dup = false;   // the subgraph holds each position once
                            for (int q = 0; q < p; ++q)
                                if (lm->puck_cluster[q] == extreme && lm->puck_x[q] == lm->puck_x[p] && lm->puck_y[q] == lm->puck_y[p]) dup = true;
                            if (dup) continue;
                            if (lm->puck_y[p] > largestYr) {
                                largestYr = lm->puck_y[p];
                                li = p;
                            }
                            if (lm->puck_y[p] < smallestYr) {
                                smallestYr = lm->puck_y[p];
                                si = p;
                            }
                        }
                        int pick = lm->puck_degree[li] <= lm->puck_degree[si] ? li : si;
                        if (lmReachable(lm->puck_x[pick], lm->puck_y[pick])) {
                            hasTarget = 1;
                            tx = lm->puck_x[pick];
                            ty = lm->puck_y[pick];
                            ttype = lm->cluster_type[extreme];
                        }
                    }
                }
            }
            rs->has_target = hasTarget;
            rs->target_x = hasTarget ? tx : 0.0f;
            rs->target_y = hasTarget ? ty : 0.0f;
            rs->target_type = hasTarget ? ttype : 0;
            return hasTarget != 0;
        };
        // ProbSeekController.matchTarget (:286-301): the closest cluster centroid (carrying; never the carried cluster)
        // or the closest puck of the kept clusters (not carrying) to the last target, within the prediction distance
        auto matchTarget = [&]() {
            int have = 0;
            float bx = 0, by = 0, smallest = INFINITY;
            int btype = -1;
            V2 last = mk(rs->target_x, rs->target_y);
            for (int c = 0; c < nC; ++c) {
                if (!lm->cluster_kept[c]) continue;
                if (carrying) {
                    if (c == lm->carried_cluster) continue;
                    float ds = distSq(last, mk(lm->cluster_x[c], lm->cluster_y[c]));
                    if (ds < sc.predictionDistSqd && ds < smallest) {
                        smallest = ds;
                        bx = lm->cluster_x[c];
                        by = lm->cluster_y[c];
                        btype = lm->cluster_type[c];
                        have = 1;
                    }
                } else {
                    for (int p = 0; p < nP; ++p) {
                        if (lm->puck_cluster[p] != c) continue;
                        float ds = distSq(last, mk(lm->puck_x[p], lm->puck_y[p]));
                        if (ds < sc.predictionDistSqd && ds < smallest) {
                            smallest = ds;
                            bx = lm->puck_x[p];
                            by = lm->puck_y[p];
                            btype = lm->cluster_type[c];
                            have = 1;
                        }
                    }
                }
            }
            if (have && !lmReachable(bx, by)) have = 0;
            rs->has_target = have;
            rs->target_x = have ? bx : 0.0f;
            rs->target_y = have ? by : 0.0f;
            rs->target_type = have ? btype : 0;
            return have != 0;
        };
        switch (state) {
            case PS_PS_PU_SCAN:
                if (carrying) TRANSITION(PS_PS_DE_SCAN);
                else if (selectTarget(false, true)) TRANSITION(PS_PS_PU_TARGET);
                break;
            case PS_PS_PU_TARGET:
                if (carrying) TRANSITION(PS_PS_DE_SCAN);
                else if (!matchTarget()) TRANSITION(PS_PS_PU_SCAN);
                else if (stepCount - startCount > sc.pickUpTime) TRANSITION(PS_PS_PU_SCAN);
                break;
            case PS_PS_DE_SCAN:
                if (!carrying) TRANSITION(PS_PS_PU_SCAN);
                else if (selectTarget(true, false)) TRANSITION(PS_PS_DE_TARGET);
                break;
            case PS_PS_DE_TARGET:
                if (!carrying) TRANSITION(PS_PS_PU_SCAN);
                else if (lm->carried_cluster >= 0 && lm->cluster_size[lm->carried_cluster] > 1) TRANSITION(PS_PS_DE_PUSH);
                else if (!matchTarget()) TRANSITION(PS_PS_DE_SCAN);
                else if (stepCount - startCount > sc.placementTime) TRANSITION(PS_PS_DE_SCAN);
                break;
            case PS_PS_DE_PUSH:
                if (stepCount - startCount >= sc.pushInTime) TRANSITION(PS_PS_DE_BACKUP);
                break;
            case PS_PS_DE_BACKUP:
                if (stepCount - startCount >= sc.backupTime) {
                    TRANSITION(PS_PS_DE_TURN);
                    // initiateRandomTurn (:303-308)
                    const int TURNTIME_MIN = 10 / 2, ONE_EIGHTY_TIME = 10;
                    rs->bhd_turn_dir = lm->left_sum <= lm->right_sum ? 1 : -1;
                    rs->bhd_turn_time = TURNTIME_MIN + jrNextInt(*rng, ONE_EIGHTY_TIME - TURNTIME_MIN + 1);
                }
                break;
            case PS_PS_DE_TURN:
                if (stepCount - startCount >= rs->bhd_turn_time) TRANSITION(PS_PS_PU_SCAN);
                break;
        }
#undef TRANSITION
        rs->ctrl_state = state;
        rs->start_count = startCount;
        int useVfh = 0;
        float angle = 0.0f, fwd = 0.0f, tq = 0.0f;
        if (state == PS_PS_PU_SCAN || state == PS_PS_DE_SCAN) {
            angle = sc.stDev * (float)jrNextGaussian(*rng);   // wander
            useVfh = 1;
        } else if (state == PS_PS_PU_TARGET || state == PS_PS_DE_TARGET) {
            movementCommand(1.0f, (float)atan2((double)rs->target_y, (double)rs->target_x), fwd, tq);
        } else if (state == PS_PS_DE_PUSH) {
            movementCommand(1.0f, 0.0f, fwd, tq);
        } else if (state == PS_PS_DE_BACKUP) {
            movementCommand(-1.0f, 0.0f, fwd, tq);
        } else {
            movementCommand(0.0f, (float)(3.14159265358979323846 / 2 * rs->bhd_turn_dir), fwd, tq);
        }
        rs->cmd_forwards = fwd;
        rs->cmd_torque = tq;
        tw.flag[1] = useVfh;
        tw.limits[0] = fbits(angle);
    }
    ctx.sync();
    int useVfh = tw.flag[1];
    float angle = bitsf(tw.limits[0]);
    ctx.sync();
    if (useVfh) {
        float fwd, tq;
        applyVFH(ctx, T, sc, tw, occ, lm, rs, angle, false, fwd, tq);
        if (ctx.tid() == 0) {
            rs->cmd_forwards = fwd;
            rs->cmd_torque = tq;
        }
    }
    if (ctx.tid() == 0) {
        rs->state_counts[rs->ctrl_state & 7]++;
        if (stepCount % 100 == 0)
            for (int i = 0; i < 8; i++) rs->state_counts[i] = 0;
    }
    ctx.sync();
}

// Arena.step's think loop for one arena (sense already done): robots in creation order, shared or per-robot RNG
template <class Ctx>
PS_HD void thinkArena(Ctx& ctx, const SenseTables& T, const SenseCfg& sc, const ThinkWork& tw, int R, uint8_t* occ /*[R][cells]*/,
                      ps_localmap* lm /*[R]*/, const float* pose /*[R][3]*/, ps_robot_state* rs /*[R]*/, ps_arena_state* as, JRandom* rngShared) {
    int nCells = T.occW * T.occH;
    if (ctx.tid() == 0) {
        rngShared->seed = as->rng_seed;
        rngShared->haveNextNextGaussian = as->rng_have_gaussian;
        rngShared->nextNextGaussian = as->rng_next_gaussian;
    }
    ctx.sync();
    int stepCount = as->step_count;
    for (int r = 0; r < R; ++r) {
        JRandom* rng = rngShared;
        if (sc.rngMode == PS_RNG_PER_ROBOT) {
            // per-robot streams live in the robot state; use the shared slot as the working copy
            if (ctx.tid() == 0) {
                rngShared->seed = rs[r].rng_seed;
                rngShared->haveNextNextGaussian = rs[r].rng_have_gaussian;
                rngShared->nextNextGaussian = rs[r].rng_next_gaussian;
            }
            ctx.sync();
        }
        if (sc.controller == PS_CTRL_BHD) thinkBHD(ctx, T, sc, tw, occ + (size_t)r * nCells, lm + r, rs + r, rng, stepCount);
        else if (sc.controller == PS_CTRL_PROBSEEK) thinkProbSeek(ctx, T, sc, tw, occ + (size_t)r * nCells, lm + r, rs + r, rng, stepCount);
        else thinkCacheCons(ctx, T, sc, tw, occ + (size_t)r * nCells, lm + r, rs + r, rng, pose + 3 * r, stepCount, r < sc.nInformed);
        if (sc.rngMode == PS_RNG_PER_ROBOT) {
            if (ctx.tid() == 0) {
                rs[r].rng_seed = rngShared->seed;
                rs[r].rng_have_gaussian = rngShared->haveNextNextGaussian;
                rs[r].rng_next_gaussian = rngShared->nextNextGaussian;
            }
            ctx.sync();
        }
    }
    if (ctx.tid() == 0) {
        if (sc.rngMode != PS_RNG_PER_ROBOT) {
            as->rng_seed = rngShared->seed;
            as->rng_have_gaussian = rngShared->haveNextNextGaussian;
            as->rng_next_gaussian = rngShared->nextNextGaussian;
        }
        as->step_count = stepCount + 1;
    }
    ctx.sync();
}

// ------------------------------------------------------------------ host-side table construction
struct SenseSetup {
    bool ready = false;
    std::string error;
    int occW = 0, occH = 0;
    SenseTables T;      // pointers into the vectors below (host) -- the engine re-points a copy at device memory
    SenseCfg sc;
    std::vector<float> aXr, aYr, cXr, cYr, vfhBeta, vfhMag;
    std::vector<uint16_t> aRect, aImg;
    std::vector<int16_t> calibIdx, cellSrc, vfhStartK, vfhStopK;
    std::vector<uint8_t> cFlags, vfhSide, aBin;
    std::vector<PixRec> aPix;
    std::vector<uint16_t> cellRect;

    bool build(const ps_config& cfg, const float* XrYr, const uint16_t* xy, const uint8_t* flags, int n, int w, int h) {
        ready = false;
        if (w <= 0 || h <= 0 || w * h > 65535 || n > 32767) {
            error = "unsupported image size";
            return false;
        }
        std::memset(&T, 0, sizeof(T));
        T.imgW = w;
        T.imgH = h;
        calibIdx.assign((size_t)w * h, -1);
        cXr.assign(n, 0);
        cYr.assign(n, 0);
        cFlags.assign(n, 0);
        for (int k = 0; k < n; k++) {
            int x = xy[2 * k], y = xy[2 * k + 1];
            if (x >= w || y >= h) {
                error = "calibration pixel outside the image";
                return false;
            }
            calibIdx[(size_t)x * h + y] = (int16_t)k;   // a later row for the same pixel replaces the earlier one
            cXr[k] = XrYr[2 * k];
            cYr[k] = XrYr[2 * k + 1];
            cFlags[k] = flags[k];
        }
        // LocalMap constructor (src/localmap/LocalMap.java:90-121): extent of the calibrated ground-plane points
        float minXr = INFINITY, maxXr = -INFINITY, minYr = INFINITY, maxYr = -INFINITY;
        int nHole = 0;
        int x0 = w, x1 = -1, y0 = h, y1 = -1;
        for (int x = 0; x < w; x++)
            for (int y = 0; y < h; y++) {
                int k = calibIdx[(size_t)x * h + y];
                if (k < 0) continue;
                minXr = std::min(minXr, cXr[k]);
                maxXr = std::max(maxXr, cXr[k]);
                minYr = std::min(minYr, cYr[k]);
                maxYr = std::max(maxYr, cYr[k]);
                if (cFlags[k] & 2) nHole++;
                if (!(cFlags[k] & 1)) {
                    x0 = std::min(x0, x);
                    x1 = std::max(x1, x);
                    y0 = std::min(y0, y);
                    y1 = std::max(y1, y);
                }
            }
        if (x1 < 0) {
            error = "no active pixels";
            return false;
        }
        T.nHolePixels = nHole;
        T.maxYr = maxYr;
        occW = (int)(0.5 + (maxYr - minYr) / 1.0f) + 1;
        occH = (int)(0.5 + (maxXr) / 1.0f) + 1;
        T.occW = occW;
        T.occH = occH;
        T.x0 = x0;
        T.y0 = y0;
        T.RW = x1 - x0 + 1;
        T.RH = y1 - y0 + 1;
        // active pixels in GridCamera's scan order does not matter (each pixel is independent); keep x-major order
        aXr.clear();
        aYr.clear();
        aRect.clear();
        aImg.clear();
        std::vector<int> activeOfPixel((size_t)w * h, -1);
        float maxRangeSq = 0;
        T.bbXmin = T.bbYmin = INFINITY;
        T.bbXmax = T.bbYmax = -INFINITY;
        for (int x = 0; x < w; x++)
            for (int y = 0; y < h; y++) {
                int k = calibIdx[(size_t)x * h + y];
                if (k < 0 || (cFlags[k] & 1)) continue;
                activeOfPixel[(size_t)x * h + y] = (int)aXr.size();
                aXr.push_back(cXr[k]);
                aYr.push_back(cYr[k]);
                aRect.push_back((uint16_t)((x - x0) * T.RH + (y - y0)));
                aImg.push_back((uint16_t)(y * w + x));
                maxRangeSq = std::max(maxRangeSq, cXr[k] * cXr[k] + cYr[k] * cYr[k]);
                T.bbXmin = std::min(T.bbXmin, cXr[k]);
                T.bbXmax = std::max(T.bbXmax, cXr[k]);
                T.bbYmin = std::min(T.bbYmin, cYr[k]);
                T.bbYmax = std::max(T.bbYmax, cYr[k]);
            }
        T.nActive = (int)aXr.size();
        T.maxRange = std::sqrt(maxRangeSq);
        // robot-frame grid over the footprint of the active pixels (candidate binning of the camera stage)
        T.binSize = 8.0f;
        for (;;) {
            T.binsX = (int)std::floor((T.bbXmax - T.bbXmin) / T.binSize) + 1;
            T.binsY = (int)std::floor((T.bbYmax - T.bbYmin) / T.binSize) + 1;
            if (T.binsX * T.binsY <= kMaxBins) break;
            T.binSize *= 1.25f;
        }
        aBin.resize(aXr.size());
        for (size_t k = 0; k < aXr.size(); k++) {
            int bxi = std::min(T.binsX - 1, std::max(0, (int)std::floor((aXr[k] - T.bbXmin) / T.binSize)));
            int byi = std::min(T.binsY - 1, std::max(0, (int)std::floor((aYr[k] - T.bbYmin) / T.binSize)));
            aBin[k] = (uint8_t)(bxi * T.binsY + byi);
        }
        // LocalMap.extractOccupancy (:210-229): writers are the calibrated, non-hole pixels whose type is not HIDDEN,
        // i.e. statically the non-gripperBody ones; the last writer in (x outer, y inner) order wins.
        cellSrc.assign((size_t)occW * occH, -1);
        for (int x = 0; x < w; x++)
            for (int y = 0; y < h; y++) {
                int k = calibIdx[(size_t)x * h + y];
                if (k < 0 || (cFlags[k] & 2) || (cFlags[k] & 1)) continue;
                int gi = (int)(0.5 + (maxYr - cYr[k]) / 1.0f);
                int gj = (int)(0.5 + (cXr[k]) / 1.0f);
                if (gi < 0 || gi >= occW || gj < 0 || gj >= occH) continue;
                cellSrc[(size_t)gi * occH + gj] = (int16_t)activeOfPixel[(size_t)x * h + y];
            }
        // VFHPlus.initDataStructures (src/localmap/VFHPlus.java:170-217) and the per-cell constants of
        // computePrimary / computeMasked (:313-345, :358-403)
        size_t nc = (size_t)occW * occH;
        vfhBeta.assign(nc, 0);
        vfhMag.assign(nc, 0);
        vfhStartK.assign(nc, 0);
        vfhStopK.assign(nc, 0);
        vfhSide.assign(nc, 0);
        const float MAX_SENSED_DISTANCE_SQD = 10000.0f / 3;
        const float A = 1.0f, B = 1.0f / MAX_SENSED_DISTANCE_SQD;
        float lcx = -6.5f, lcy = 12.5f;
        float safetyRadiusSqd = (float)std::pow((double)(lcy + cfg.vfh_safety_mask), 2.0);
        for (int i = 0; i < occW; i++)
            for (int j = 0; j < occH; j++) {
                float vx = 1.0f * j, vy = maxYr - 1.0f * i;
                float dSqd = vx * vx + vy * vy;
                size_t c = (size_t)i * occH + j;
                vfhMag[c] = std::max(0.0f, A - B * dSqd);
                float beta = (float)std::atan2((double)vy, (double)vx);
                float gamma = (float)(std::asin((double)cfg.vfh_safety_gamma / std::sqrt((double)dSqd)));
                vfhBeta[c] = beta;
                int sk = vfhAngleToIndex(beta + gamma), ek = vfhAngleToIndex(beta - gamma) + 1;
                vfhStartK[c] = (int16_t)std::max(-30000, std::min(30000, sk));
                vfhStopK[c] = (int16_t)std::max(-30000, std::min(30000, ek));
                uint8_t side = 0;
                V2 v = mk(vx, vy);
                if (distSq(v, mk(lcx, lcy)) < safetyRadiusSqd) side |= 1;
                if (distSq(v, mk(lcx, -lcy)) < safetyRadiusSqd) side |= 2;
                vfhSide[c] = side;
            }
        T.nPuckTypes = cfg.n_puck_types;
        T.pucksPerType = std::max(1, cfg.n_pucks / cfg.n_puck_types);
        T.aXr = aXr.data();
        T.aYr = aYr.data();
        T.aRect = aRect.data();
        T.aImg = aImg.data();
        T.aBin = aBin.data();
        aPix.resize(aXr.size());
        for (size_t k = 0; k < aXr.size(); k++) {
            aPix[k].xr = aXr[k];
            aPix[k].yr = aYr[k];
            aPix[k].rectBin = (uint32_t)aRect[k] | ((uint32_t)aBin[k] << 16);
            aPix[k].img = aImg[k];
        }
        // the camera stage walks the pixels bin by bin (a warp then sees one bin class and one candidate list)
        std::stable_sort(aPix.begin(), aPix.end(), [](const PixRec& a, const PixRec& b) { return (a.rectBin >> 16) < (b.rectBin >> 16); });
        {
            int nB = T.binsX * T.binsY;
            size_t k = 0;
            for (int b = 0; b <= kMaxBins; b++) {
                while (b < nB && k < aPix.size() && (int)(aPix[k].rectBin >> 16) < b) k++;
                T.binStart[b] = (uint16_t)(b < nB ? k : aPix.size());
            }
        }
        T.aPix = aPix.data();
        cellRect.resize(cellSrc.size());
        for (size_t c = 0; c < cellSrc.size(); c++) cellRect[c] = cellSrc[c] < 0 ? (uint16_t)0xFFFF : aRect[cellSrc[c]];
        T.cellRect = cellRect.data();
        T.calibIdx = calibIdx.data();
        T.cXr = cXr.data();
        T.cYr = cYr.data();
        T.cFlags = cFlags.data();
        T.cellSrc = cellSrc.data();
        T.vfhBeta = vfhBeta.data();
        T.vfhMag = vfhMag.data();
        T.vfhStartK = vfhStartK.data();
        T.vfhStopK = vfhStopK.data();
        T.vfhSide = vfhSide.data();
        sc.clusterThreshold = cfg.cluster_distance_threshold;
        sc.carryLo = cfg.carry_threshold_lo;
        sc.carryHi = cfg.carry_threshold_hi;
        sc.tauLo = cfg.vfh_tau_lo * A;
        sc.tauHi = cfg.vfh_tau_hi * A;
        sc.controller = cfg.controller;
        sc.rngMode = cfg.rng_mode;
        sc.placementTime = cfg.ps_placement_time;
        sc.k1 = cfg.cc_k1;
        sc.k2 = cfg.cc_k2;
        sc.predictionDistSqd = cfg.cc_prediction_distance_sqd;
        sc.stDev = cfg.cc_st_dev;
        sc.kAbsolute = cfg.cc_k_absolute;
        sc.kRelative = cfg.cc_k_relative;
        sc.forgetProb = cfg.cc_forget_prob;
        sc.homeSeparation = cfg.cc_home_separation_distance;
        sc.pickUpTime = cfg.cc_pick_up_time;
        sc.pushInTime = cfg.cc_push_in_time;
        sc.backupTime = cfg.cc_backup_time;
        sc.exileTime = cfg.cc_exile_time;
        sc.spitOutTime = cfg.cc_spit_out_time;
        sc.cacheSelection = cfg.cc_cache_selection;
        sc.cacheSeparation = cfg.cc_cache_separation;
        sc.ignorePucks = cfg.cc_ignore_pucks;
        sc.bhdPushInTime = cfg.bhd_push_in_time;
        sc.bhdBackupTime = cfg.bhd_backup_time;
        sc.nInformed = cfg.n_informed_robots;
        ready = true;
        return true;
    }
};

#if defined(PS_HOST_EMU)
// ------------------------------------------------------------------ host emulation entry points (tests only)
struct EmuObs {
    std::vector<uint8_t> camera, occ;
    std::vector<ps_localmap> lm;
    std::vector<float> pose;
};
inline void emuSenseArena(HostCtx& ctx, const PhysEnv& e, const SenseSetup& ss, std::vector<char>& buf, const ArenaState& st, ps_robot_state* rs,
                          EmuObs& o, bool camera) {
    int R = e.cfg.d.R;
    int nCells = ss.occW * ss.occH;
    Carver m = {nullptr, 0, 0, nullptr, 0};
    carveSense(ss.T, m);
    carveThink(m);
    if (buf.size() < m.slowUsed + 64) buf.assign(m.slowUsed + 64, 0);
    Carver cv = {nullptr, 0, 0, buf.data(), 0};
    SenseWork sw = carveSense(ss.T, cv);
    o.occ.assign((size_t)R * nCells, 0);
    o.lm.assign(R, ps_localmap());
    o.pose.assign((size_t)R * 3, 0);
    if (camera) o.camera.assign((size_t)R * ss.T.imgW * ss.T.imgH, 0);
    std::vector<BlobRec> blobs(kMaxBlobs);
    for (int r = 0; r < R; r++) {
        std::memset(&o.lm[r], 0, sizeof(ps_localmap));
        SenseHead head;
        SenseOut so = {&head, blobs.data()};
        // camera stage (one CTA per robot on the GPU), then the local-map stage (one thread per robot)
        senseRobot(ctx, e, ss.T, ss.sc, sw, st, r, camera ? o.camera.data() + (size_t)r * ss.T.imgW * ss.T.imgH : nullptr,
                   o.occ.data() + (size_t)r * nCells, o.pose.data() + 3 * r, so);
        localMapFromBlobs(ss.T, ss.sc, head, blobs.data(), o.occ.data() + (size_t)r * nCells, rs + r, o.lm[r]);
    }
}
template <class Ctx>
void senseAndThinkArena(Ctx& ctx, const PhysEnv& e, const SenseSetup& ss, std::vector<char>& buf, const ArenaState& st, ps_robot_state* rs,
                        ps_arena_state* as, void*) {
    EmuObs o;
    emuSenseArena(ctx, e, ss, buf, st, rs, o, false);
    Carver cv = {nullptr, 0, 0, buf.data(), 0};
    carveSense(ss.T, cv);
    ThinkWork tw = carveThink(cv);
    JRandom rng;
    thinkArena(ctx, ss.T, ss.sc, tw, e.cfg.d.R, o.occ.data(), o.lm.data(), o.pose.data(), rs, as, &rng);
}
template <class Ctx>
void senseArenaToObs(Ctx& ctx, const PhysEnv& e, const SenseSetup& ss, std::vector<char>& buf, const ArenaState& st, ps_robot_state* rs, int a,
                     ps_obs_view* v) {
    EmuObs o;
    emuSenseArena(ctx, e, ss, buf, st, rs, o, v->camera != nullptr);
    int R = e.cfg.d.R;
    size_t nCells = (size_t)ss.occW * ss.occH, nImg = (size_t)ss.T.imgW * ss.T.imgH;
    if (v->camera) std::memcpy(v->camera + (size_t)a * R * nImg, o.camera.data(), R * nImg);
    if (v->occupancy) std::memcpy(v->occupancy + (size_t)a * R * nCells, o.occ.data(), R * nCells);
    if (v->localmap) std::memcpy(v->localmap + (size_t)a * R, o.lm.data(), R * sizeof(ps_localmap));
    if (v->pose) std::memcpy(v->pose + (size_t)a * R * 3, o.pose.data(), R * 3 * sizeof(float));
}
#endif

}  // namespace ps
