cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" "PS_ENGINE_LIB=$GRAFT_REPO_ROOT/puckswarm_b200/csrc/libps_post8.so" 2>&1 | tee gpurun_out/r2_ab_post8.log
AB_WARMUP=5 AB_STEPS=20 bash tools/gpu_ab.sh "PS_X=0" "PS_ENGINE_LIB=$GRAFT_REPO_ROOT/puckswarm_b200/csrc/libps_post8.so" 2>&1 | tee -a gpurun_out/r2_ab_post8.log
