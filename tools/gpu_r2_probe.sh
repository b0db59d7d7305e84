cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AB_STEPS=40 bash tools/gpu_ab.sh "PS_DIAG_HEAVY_CAP=3" "PS_DIAG_HEAVY_CAP=3 PS_WORKER_BLOCKS_PER_SM=6" "PS_DIAG_HEAVY_CAP=3 PS_GROUPS=2" 2>&1 | tee gpurun_out/r2_probe_heavycap.log
PS_DIAG_HEAVY_CAP=3 python tools/gpu_island_census.py 4096 300 30 2>&1 | tee -a gpurun_out/r2_probe_heavycap.log
