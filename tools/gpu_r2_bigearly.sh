# Round 2: streaming big-island workers (start with stage A): parity + A/B
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "world_step or free_running or dense_arena or c3_shape or lockstep or contact_cache or large_arena or batch_position or fast_mode" 2>&1 | tail -3
bash tools/gpu_ab.sh "PS_BIG_EARLY=0" "PS_BIG_EARLY=1" "PS_BIG_EARLY=1 PS_BIG_WORKERS_PER_SM=2" 2>&1 | tee gpurun_out/r2_ab_bigearly.log
AB_WARMUP=5 AB_STEPS=20 bash tools/gpu_ab.sh "PS_BIG_EARLY=0" "PS_BIG_EARLY=1" 2>&1 | tee -a gpurun_out/r2_ab_bigearly.log
