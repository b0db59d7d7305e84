cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PS_ENGINE_LIB=$GRAFT_REPO_ROOT/puckswarm_b200/csrc/libpuckswarm_b200_diag.so python tools/gpu_island_census.py 4096 300 30 2>&1 | grep "k_phys_pre" | tee gpurun_out/r2_pre_phases.log
PS_ENGINE_LIB=$GRAFT_REPO_ROOT/puckswarm_b200/csrc/libpuckswarm_b200_diag.so python tools/gpu_phase_stats.py 4096 2>&1 | grep -E "^sense|findnew|syncfix|toi|cycles per|robot-steps" | tee -a gpurun_out/r2_pre_phases.log
