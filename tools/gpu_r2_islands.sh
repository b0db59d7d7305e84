# Round 2: parity of the split island solver (k_island_vel + k_island_pos) and its A/B against the one-kernel solver
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for G in 4 2 8; do
  echo "== parity PS_ISLAND_G=$G"
  PS_ISLAND_G=$G timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "world_step or free_running or dense_arena or c3_shape or lockstep or contact_cache or large_arena" 2>&1 | tail -3
done
bash tools/gpu_ab.sh "PS_ISLAND_G=0" "PS_ISLAND_G=4" "PS_ISLAND_G=2" "PS_ISLAND_G=8" "PS_ISLAND_G=4 PS_POS_BLOCKS_PER_SM=2" "PS_ISLAND_G=4 PS_GROUPS=2" "PS_ISLAND_G=4 PS_GROUPS=4" "PS_ISLAND_G=8 PS_GROUPS=4" "PS_ISLAND_G=0 PS_GROUPS=2" 2>&1 | tee gpurun_out/r2_ab_islands.log
