cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "world_step or free_running or c3_shape or dense_arena or large_arena or lockstep" 2>&1 | tail -2
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" 2>&1 | tee gpurun_out/r2_ab_toi.log
AB_WARMUP=5 AB_STEPS=20 bash tools/gpu_ab.sh "PS_X=0" 2>&1 | tee -a gpurun_out/r2_ab_toi.log
