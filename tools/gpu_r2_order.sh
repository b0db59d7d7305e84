cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AB_WARMUP=5 AB_STEPS=20 bash tools/gpu_ab.sh "PS_X=0" "PS_MERGE_BATCH=0" "PS_NO_PRIORITY=1" "PS_MERGE_BATCH=0 PS_NO_PRIORITY=1" "PS_WORKER_BLOCKS_PER_SM=3" "PS_WORKER_BLOCKS_PER_SM=5" 2>&1 | tee gpurun_out/r2_ab_order.log
AB_STEPS=40 bash tools/gpu_ab.sh "PS_MERGE_BATCH=0" "PS_NO_PRIORITY=1" 2>&1 | tee -a gpurun_out/r2_ab_order.log
