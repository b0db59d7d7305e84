# Round 2: arena groups whose island kernels are sized for the group (half the workers for half the arenas), so that a group's
# island phase leaves shared memory and warp slots to the other group's wide kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" "PS_GROUPS=2" "PS_GROUPS=2 PS_WORKER_BLOCKS_PER_SM=2 PS_BATCH_BLOCKS_PER_SM=1 PS_BIG_BLOCKS_PER_SM=2" "PS_GROUPS=2 PS_WORKER_BLOCKS_PER_SM=2 PS_BATCH_BLOCKS_PER_SM=1" "PS_GROUPS=2 PS_WORKER_BLOCKS_PER_SM=3 PS_BATCH_BLOCKS_PER_SM=1" "PS_GROUPS=2 PS_WORKER_BLOCKS_PER_SM=2" "PS_GROUPS=3 PS_WORKER_BLOCKS_PER_SM=2 PS_BATCH_BLOCKS_PER_SM=1 PS_BIG_BLOCKS_PER_SM=2" "PS_GROUPS=4 PS_WORKER_BLOCKS_PER_SM=1 PS_BATCH_BLOCKS_PER_SM=1 PS_BATCH_THREADS=32 PS_BIG_BLOCKS_PER_SM=1" 2>&1 | tee gpurun_out/r2_ab_groups2.log
