# Round 2: is the island phase bound by shared-memory residency? (three island kernels alone: 3.2-3.8 ms, together: 5.4-5.9 ms)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" "PS_BATCH_THREADS=32" "PS_WORKER_BLOCKS_PER_SM=2" "PS_WORKER_BLOCKS_PER_SM=2 PS_BATCH_THREADS=32" "PS_BIG_BLOCKS_PER_SM=1" "PS_BIG_BLOCKS_PER_SM=2 PS_WORKER_BLOCKS_PER_SM=2 PS_BATCH_THREADS=32" "PS_WORKER_BLOCKS_PER_SM=1 PS_BATCH_THREADS=32 PS_BATCH_BLOCKS_PER_SM=4" "PS_BATCH_BLOCKS_PER_SM=4" 2>&1 | tee gpurun_out/r2_ab_residency.log
