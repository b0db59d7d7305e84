# Round 2: heaviest warp-slot islands first (XXL / XL buckets): parity + A/B + slowest-island diagnostics
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "world_step or free_running or dense_arena or c3_shape or lockstep or large_arena" 2>&1 | tail -3
bash tools/gpu_ab.sh "PS_X=0" 2>&1 | tee gpurun_out/r2_ab_prio.log
AB_WARMUP=5 AB_STEPS=20 bash tools/gpu_ab.sh "PS_X=0" 2>&1 | tee -a gpurun_out/r2_ab_prio.log
python tools/gpu_island_hist.py 4096 300 2>&1 | head -3 | tee -a gpurun_out/r2_ab_prio.log
