# Round 2: heavy warp-bucket islands a warp each, the lighter ones through the split solver (G lanes per island): parity + A/B
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PS_ISLAND_G=8 PS_ISLAND_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "world_step or free_running or dense_arena or c3_shape or steady" 2>&1 | tail -3
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" "PS_ISLAND_G=8 PS_ISLAND_SPLIT=1" "PS_ISLAND_G=8 PS_ISLAND_SPLIT=1 PS_POS_BLOCKS_PER_SM=8" "PS_ISLAND_G=4 PS_ISLAND_SPLIT=1" "PS_ISLAND_G=4 PS_ISLAND_SPLIT=1 PS_POS_BLOCKS_PER_SM=6" "PS_ISLAND_G=8 PS_ISLAND_SPLIT=1 PS_POS_BLOCKS_PER_SM=8 PS_WORKER_BLOCKS_PER_SM=2" "PS_ISLAND_G=2 PS_ISLAND_SPLIT=1 PS_POS_BLOCKS_PER_SM=3" 2>&1 | tee gpurun_out/r2_ab_split.log
