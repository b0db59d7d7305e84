# Round 2: camera-stage changes (bin classes, run-based blob statistics, coloured-pixel list from the camera pass): parity, phase split, A/B
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "camera or lockstep or free_running or c3_shape or baseline_config or large_arena or informed or wide_cluster" 2>&1 | tail -3
PS_ENGINE_LIB=$GRAFT_REPO_ROOT/puckswarm_b200/csrc/libpuckswarm_b200_diag.so python tools/gpu_phase_stats.py 4096 2>&1 | grep -E "^sense|robot-steps" | tee gpurun_out/r2_phase_stats_sense.log
bash tools/gpu_ab.sh "PS_X=0" 2>&1 | tee gpurun_out/r2_ab_sense.log
AB_WARMUP=5 AB_STEPS=20 bash tools/gpu_ab.sh "PS_X=0" 2>&1 | tee -a gpurun_out/r2_ab_sense.log
