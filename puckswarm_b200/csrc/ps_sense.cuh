#pragma once
#include "ps_layout.cuh"
#include <vector>
namespace ps {
struct SenseSetup {
    bool ready = false;
    int occW = 0, occH = 0;
    bool build(const ps_config&, const float*, const uint16_t*, const uint8_t*, int, int, int) { ready = true; return true; }
};
template <class Ctx>
void senseAndThinkArena(Ctx&, const PhysEnv&, const SenseSetup&, std::vector<char>&, const ArenaState&, ps_robot_state*, ps_arena_state*, void*) {}
template <class Ctx>
void senseArenaToObs(Ctx&, const PhysEnv&, const SenseSetup&, std::vector<char>&, const ArenaState&, ps_robot_state*, int, ps_obs_view*) {}
}
