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
    PS_HD void* take(size_t bytes) {
        bytes = (bytes + 15) & ~(size_t)15;
        if (fastUsed + bytes <= fastCap) {
            void* p = fast ? (void*)(fast + fastUsed) : nullptr;
            fastUsed += bytes;
            return p;
        }
        void* p = slow ? (void*)(slow + slowUsed) : nullptr;
        slowUsed += bytes;
        return p;
    }
};

// Carve PhysWork. Call with null buffers to measure (fastUsed / slowUsed), then with real ones.
PS_HD PhysWork carveWork(const PhysCfg& c, Carver& cv) {
    PhysWork w;
    int NB = c.d.NB, R = c.d.R, NP = c.d.NP;
    w.counters = (int*)cv.take(8 * sizeof(int));
    w.stats = (int*)cv.take(8 * sizeof(int));
    float* bodyArr = (float*)cv.take((size_t)9 * NB * sizeof(float));
    w.cx = bodyArr;
    w.cy = bodyArr ? bodyArr + NB : nullptr;
    w.a = bodyArr ? bodyArr + 2 * NB : nullptr;
    w.vx = bodyArr ? bodyArr + 3 * NB : nullptr;
    w.vy = bodyArr ? bodyArr + 4 * NB : nullptr;
    w.w = bodyArr ? bodyArr + 5 * NB : nullptr;
    w.c0x = bodyArr ? bodyArr + 6 * NB : nullptr;
    w.c0y = bodyArr ? bodyArr + 7 * NB : nullptr;
    w.a0 = bodyArr ? bodyArr + 8 * NB : nullptr;
    float* rob = (float*)cv.take((size_t)11 * R * sizeof(float));
    w.rpx = rob;
    w.rpy = rob ? rob + R : nullptr;
    w.rc = rob ? rob + 2 * R : nullptr;
    w.rs = rob ? rob + 3 * R : nullptr;
    w.r0px = rob ? rob + 4 * R : nullptr;
    w.r0py = rob ? rob + 5 * R : nullptr;
    w.r0c = rob ? rob + 6 * R : nullptr;
    w.r0s = rob ? rob + 7 * R : nullptr;
    w.fx = rob ? rob + 8 * R : nullptr;
    w.fy = rob ? rob + 9 * R : nullptr;
    w.tq = rob ? rob + 10 * R : nullptr;
    w.adjOff = (int*)cv.take((size_t)(NB + 1) * sizeof(int));
    w.adjFill = (int*)cv.take((size_t)NB * sizeof(int));
    w.islandStart = (int*)cv.take((size_t)(NB + 1) * sizeof(int));
    w.stack = (uint16_t*)cv.take((size_t)NB * sizeof(uint16_t));
    w.bodyFlag = (uint8_t*)cv.take((size_t)NB);
    w.adj = (uint16_t*)cv.take((size_t)2 * c.MT * sizeof(uint16_t));
    w.order = (uint16_t*)cv.take((size_t)c.MT * sizeof(uint16_t));
    w.tcFlag = (uint8_t*)cv.take((size_t)c.MT);
    w.bodyFat = (float*)cv.take((size_t)4 * NB * sizeof(float));
    w.moved = (uint8_t*)cv.take((size_t)NP);
    w.movedList = (int*)cv.take((size_t)NP * sizeof(int));
    w.pairs = (uint32_t*)cv.take((size_t)c.MPAIR * sizeof(uint32_t));
    w.fatOld = (float*)cv.take((size_t)4 * NP * sizeof(float));
    w.wallList = (int*)cv.take((size_t)c.MW * sizeof(int));
    w.tc = (TC*)cv.take((size_t)c.MT * sizeof(TC));
    return w;
}

}  // namespace ps
