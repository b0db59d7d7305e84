cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" "PS_BIG_SMEM_KB=100" "PS_BIG_SMEM_KB=150" "PS_BIG_SMEM_KB=220" "PS_BIG_SMEM_KB=220 PS_BIG_BLOCKS_PER_SM=1" "PS_BIG_SMEM_KB=110 PS_BIG_BLOCKS_PER_SM=2" 2>&1 | tee gpurun_out/r2_ab_bigsmem.log
