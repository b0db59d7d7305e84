# Round 2: one lane-batch kernel per bucket with its own slot size (PS_BUCKET_SLOTS=1): parity, then A/B in both regimes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PS_BUCKET_SLOTS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "world_step or free_running or c3_shape or dense_arena or lockstep" 2>&1 | tail -2
AB_WARMUP=5 AB_STEPS=20 bash tools/gpu_ab.sh "PS_X=0" "PS_BUCKET_SLOTS=1" "PS_BUCKET_SLOTS=1 PS_B1_BLOCKS_PER_SM=3 PS_B2_BLOCKS_PER_SM=3 PS_B3_BLOCKS_PER_SM=3" "PS_BUCKET_SLOTS=1 PS_B3_BLOCKS_PER_SM=1 PS_B1_BLOCKS_PER_SM=1" "PS_BUCKET_SLOTS=1 PS_WORKER_BLOCKS_PER_SM=3" 2>&1 | tee gpurun_out/r2_ab_bucketslots.log
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" "PS_BUCKET_SLOTS=1" "PS_BUCKET_SLOTS=1 PS_B1_BLOCKS_PER_SM=3 PS_B2_BLOCKS_PER_SM=3 PS_B3_BLOCKS_PER_SM=3" 2>&1 | tee -a gpurun_out/r2_ab_bucketslots.log
