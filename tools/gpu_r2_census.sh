cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/gpu_island_census.py 4096 300 30 2>&1 | tee gpurun_out/r2_island_census.log
python tools/gpu_island_census.py 4096 15 30 2>&1 | tee -a gpurun_out/r2_island_census.log
AB_STEPS=40 bash tools/gpu_ab.sh "PS_WORKER_BLOCKS_PER_SM=6" "PS_WORKER_BLOCKS_PER_SM=8" "PS_ISLAND_G=8 PS_POS_BLOCKS_PER_SM=8" "PS_ISLAND_G=8 PS_POS_BLOCKS_PER_SM=10 PS_WORKER_BLOCKS_PER_SM=2" "PS_WARP_MIN_CONTACTS=13" "PS_GROUPS=2" 2>&1 | tee gpurun_out/r2_ab_residency2.log
