# Round 2: k_phys_pre / k_phys_post with shared memory sized by the configuration, CTA width A/B
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "world_step or free_running or dense_arena or c3_shape or steady or large_arena or lockstep" 2>&1 | tail -3
PS_PRE_THREADS=64 PS_POST_THREADS=64 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "world_step or free_running or dense_arena or c3_shape or large_arena" 2>&1 | tail -3
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" "PS_PRE_THREADS=64" "PS_POST_THREADS=64" "PS_PRE_THREADS=64 PS_POST_THREADS=64" "PS_PRE_THREADS=32 PS_POST_THREADS=32" "PS_PRE_SMEM_T=512" "PS_PRE_THREADS=64 PS_PRE_SMEM_T=512" 2>&1 | tee gpurun_out/r2_ab_prepost.log
