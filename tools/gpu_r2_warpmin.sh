cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AB_WARMUP=5 AB_STEPS=20 bash tools/gpu_ab.sh "PS_X=0" "PS_WARP_MIN_CONTACTS=15" "PS_WARP_MIN_CONTACTS=13" "PS_WARP_MIN_CONTACTS=11" "PS_WARP_MIN_CONTACTS=9" "PS_WARP_MIN_CONTACTS=13 PS_WORKER_BLOCKS_PER_SM=5" "PS_WARP_MIN_CONTACTS=7" 2>&1 | tee gpurun_out/r2_ab_warpmin.log
AB_STEPS=40 bash tools/gpu_ab.sh "PS_WARP_MIN_CONTACTS=15" "PS_WARP_MIN_CONTACTS=13" 2>&1 | tee -a gpurun_out/r2_ab_warpmin.log
