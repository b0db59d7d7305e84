# Round 2: wide lane slots (islands of 17-27 contacts as lanes instead of a warp each) + slim lane slots: parity, census, A/B
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "world_step or free_running or dense_arena or c3_shape or lockstep or steady" 2>&1 | tail -3
python tools/gpu_island_census.py 4096 300 30 2>&1 | tee gpurun_out/r2_island_census_lanex.log
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" "PS_LANEX_MAX_CONTACTS=0" "PS_BATCHX_BLOCKS_PER_SM=2" "PS_LANEX_MAX_CONTACTS=22" "PS_BATCH_BLOCKS_PER_SM=3" 2>&1 | tee gpurun_out/r2_ab_lanex.log
AB_WARMUP=5 AB_STEPS=20 bash tools/gpu_ab.sh "PS_X=0" "PS_LANEX_MAX_CONTACTS=0" 2>&1 | tee -a gpurun_out/r2_ab_lanex.log
