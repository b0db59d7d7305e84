cd $GRAFT_REPO_ROOT
bash tools/gpu_bench_short.sh
PS_ENGINE_LIB=$GRAFT_REPO_ROOT/puckswarm_b200/csrc/libpuckswarm_b200_diag.so python tools/gpu_phase_stats.py 4096 2>&1 | grep -E "^sense|findnew|syncfix|toi_cycles|cycles per"
