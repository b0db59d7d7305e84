// puckswarm_b200 -- memory layout of the engine: per-arena state in HBM and per-CTA working memory.
//
// Persistent state (HBM, arena-major, SoA inside an arena; identical to ps_state_view so get/set are plain
// copies):  body [A][6][NB] f32 | senseAabb [A][4][NB] f32 | fat [A][4][NP] f32 | ckey, cinfo [A][MC] u32 |
//           cimp [A][MC][4] f32 | robot [A][R] ps_robot_state | arena [A] ps_arena_state
// Working memory of one CTA (PhysWork) is carved from a "fast" buffer (shared memory) in priority order; what
// does not fit goes to a per-CTA-slot "slow" buffer in HBM (L2-resident: slots are reused arena after arena).
#pragma once
#include "ps_spawn.cuh"

namespace ps {

struct StatePtrs {      // base pointers of the batched state arrays
    float* body;
    float* senseAabb;
    float* fat;
    uint32_t* ckey;
    uint32_t* cinfo;
    float* cimp;
    ps_robot_state* robot;
    ps_arena_state* arena;
};

PS_HD ArenaState arenaState(const StatePtrs& s, const PhysCfg& c, int a) {
    ArenaState st;
    size_t A = (size_t)a;
    st.body = s.body + A * 6 * c.d.NB;
    st.senseAabb = s.senseAabb + A * 4 * c.d.NB;
    st.fat = s.fat + A * 4 * c.d.NP;
    st.ckey = s.ckey + A * c.MC;
    st.cinfo = s.cinfo + A * c.MC;
    st.cimp = s.cimp + A * c.MC * 4;
    st.nContacts = &s.arena[a].n_contacts;
    st.invDt0 = &s.arena[a].inv_dt0;
    st.errorFlags = &s.arena[a].error_flags;
    return st;
}

struct Carver {
    char* fast;
    size_t fastCap, fastUsed;
    char* slow;
    size_t slowUsed;
    // With null base pointers (host-side measuring pass) the returned addresses are meaningless and never used.
    PS_HD void* take(size_t bytes) {
        bytes = (bytes + 15) & ~(size_t)15;
        if (fastUsed + bytes <= fastCap) {
            void* p = (void*)(fast + fastUsed);
            fastUsed += bytes;
            return p;
        }
        void* p = (void*)(slow + slowUsed);
        slowUsed += bytes;
        return p;
    }
};

// Carve PhysWork. Call with null buffers to measure (fastUsed / slowUsed), then with real ones.
PS_HD PhysWork carveWork(const PhysCfg& c, Carver& cv) {
    PhysWork w;
    int NB = c.d.NB, R = c.d.R, NP = c.d.NP;
    w.fatNow = nullptr;
    w.counters = (int*)cv.take(8 * sizeof(int));
    w.stats = (int*)cv.take(8 * sizeof(int));
    float* bodyArr = (float*)cv.take((size_t)9 * NB * sizeof(float));
    w.cx = bodyArr;
    w.cy = bodyArr + NB;
    w.a = bodyArr + 2 * NB;
    w.vx = bodyArr + 3 * NB;
    w.vy = bodyArr + 4 * NB;
    w.w = bodyArr + 5 * NB;
    w.c0x = bodyArr + 6 * NB;
    w.c0y = bodyArr + 7 * NB;
    w.a0 = bodyArr + 8 * NB;
    float* rob = (float*)cv.take((size_t)11 * R * sizeof(float));
    w.rpx = rob;
    w.rpy = rob + R;
    w.rc = rob + 2 * R;
    w.rs = rob + 3 * R;
    w.r0px = rob + 4 * R;
    w.r0py = rob + 5 * R;
    w.r0c = rob + 6 * R;
    w.r0s = rob + 7 * R;
    w.fx = rob + 8 * R;
    w.fy = rob + 9 * R;
    w.tq = rob + 10 * R;
    w.adjOff = (int*)cv.take((size_t)(NB + 1) * sizeof(int));
    w.adjFill = (int*)cv.take((size_t)NB * sizeof(int));
    w.islandStart = (int*)cv.take((size_t)(NB + 1) * sizeof(int));
    w.islandBodyStart = (int*)cv.take((size_t)(NB + 1) * sizeof(int));
    w.ibody = (uint16_t*)cv.take((size_t)NB * sizeof(uint16_t));
    w.stack = (uint16_t*)cv.take((size_t)NB * sizeof(uint16_t));
    w.bodyFlag = (uint8_t*)cv.take((size_t)NB);
    w.adj = (uint16_t*)cv.take((size_t)2 * c.MT * sizeof(uint16_t));
    w.order = (uint16_t*)cv.take((size_t)c.MT * sizeof(uint16_t));
    w.tcFlag = (uint8_t*)cv.take((size_t)c.MT);
    w.tcPair = (uint32_t*)cv.take((size_t)c.MT * sizeof(uint32_t));
    w.bodyFat = (float*)cv.take((size_t)4 * NB * sizeof(float));
    w.moved = (uint8_t*)cv.take((size_t)NP);
    w.movedList = (int*)cv.take((size_t)NP * sizeof(int));
    w.movedBody = (uint16_t*)cv.take((size_t)NB * sizeof(uint16_t));
    w.bodyPairs = (uint32_t*)cv.take((size_t)4 * NB * sizeof(uint32_t));
    w.pairs = (uint32_t*)cv.take((size_t)c.MPAIR * sizeof(uint32_t));
    w.fatOld = (float*)cv.take((size_t)4 * NP * sizeof(float));
    w.wallList = (int*)cv.take((size_t)c.MW * sizeof(int));
    w.wallBody = (int*)cv.take((size_t)c.MW * sizeof(int));
    w.wallT = (float*)cv.take((size_t)c.MW * sizeof(float));
    w.wallState = (uint8_t*)cv.take((size_t)c.MW);
    w.wallNeed = (uint8_t*)cv.take((size_t)c.MW);
    w.toiT = (float*)cv.take((size_t)NB * sizeof(float));
    w.toiContact = (int*)cv.take((size_t)NB * sizeof(int));
    w.toiCur = (int*)cv.take((size_t)NB * sizeof(int));
    w.toiCount = (int*)cv.take((size_t)NB * sizeof(int));
    w.toiFound = (uint8_t*)cv.take((size_t)NB);
    w.toiIter = (uint8_t*)cv.take((size_t)NB);
    w.toiActive = (uint8_t*)cv.take((size_t)NB);
    w.tc = (TC*)cv.take((size_t)c.MT * sizeof(TC));
    return w;
}

// PhysWork whose pointers are byte offsets (carved from a null base) -> real pointers for one arena's scratch block.
// The offsets live in kernel parameters, so a kernel only materialises the few pointers it uses (one add each).
PS_HD PhysWork rebaseWork(const PhysWork& o, char* base) {
    PhysWork w;
    size_t b = (size_t)base;
#define RB(f) w.f = (decltype(w.f))((size_t)o.f + b);
    RB(cx) RB(cy) RB(a) RB(vx) RB(vy) RB(w) RB(c0x) RB(c0y) RB(a0)
    RB(rpx) RB(rpy) RB(rc) RB(rs) RB(r0px) RB(r0py) RB(r0c) RB(r0s) RB(fx) RB(fy) RB(tq)
    RB(adjOff) RB(adjFill) RB(adj) RB(order) RB(islandStart) RB(islandBodyStart) RB(ibody) RB(stack) RB(bodyFlag) RB(tcFlag) RB(tcPair)
    RB(moved) RB(bodyFat) RB(movedList) RB(movedBody) RB(bodyPairs) RB(fatOld) RB(pairs) RB(wallList) RB(wallBody) RB(wallT) RB(wallState) RB(wallNeed)
    RB(toiT) RB(toiContact) RB(toiCur) RB(toiCount) RB(toiFound) RB(toiIter) RB(toiActive) RB(tc) RB(counters) RB(stats)
#undef RB
    w.fatNow = nullptr;
    return w;
}
// the same, written member by member into an existing (shared-memory) PhysWork
PS_HD void rebaseWorkInto(PhysWork& w, const PhysWork& o, char* base) {
    size_t b = (size_t)base;
#define RB(f) w.f = (decltype(w.f))((size_t)o.f + b);
    RB(cx) RB(cy) RB(a) RB(vx) RB(vy) RB(w) RB(c0x) RB(c0y) RB(a0)
    RB(rpx) RB(rpy) RB(rc) RB(rs) RB(r0px) RB(r0py) RB(r0c) RB(r0s) RB(fx) RB(fy) RB(tq)
    RB(adjOff) RB(adjFill) RB(adj) RB(order) RB(islandStart) RB(islandBodyStart) RB(ibody) RB(stack) RB(bodyFlag) RB(tcFlag) RB(tcPair)
    RB(moved) RB(bodyFat) RB(movedList) RB(movedBody) RB(bodyPairs) RB(fatOld) RB(pairs) RB(wallList) RB(wallBody) RB(wallT) RB(wallState) RB(wallNeed)
    RB(toiT) RB(toiContact) RB(toiCur) RB(toiCount) RB(toiFound) RB(toiIter) RB(toiActive) RB(tc) RB(counters) RB(stats)
#undef RB
    w.fatNow = nullptr;
}

}  // namespace ps
