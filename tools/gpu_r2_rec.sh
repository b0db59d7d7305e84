# Round 2: body / constraint records (16-byte shared-memory loads) in the warp-per-island position sweeps: parity + timing
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "world_step or free_running or dense_arena or c3_shape or steady or large_arena" 2>&1 | tail -3
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" 2>&1 | tee gpurun_out/r2_ab_rec.log
python tools/gpu_island_census.py 4096 300 30 2>&1 | grep -i "big islands\|warp-slot islands:" | tee -a gpurun_out/r2_ab_rec.log
