# Round 2: heavy / light arena split (PS_HEAVY_ARENAS): parity under the split, then A/B in the steady state
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PS_HEAVY_ARENAS=2 PS_TEST_STEADY_ARENAS=8 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "free_running or dense_arena or c3_shape or steady or lockstep or periodic_record or shift_and_shuffle or batch_position" 2>&1 | tail -3
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" "PS_HEAVY_ARENAS=256" "PS_HEAVY_ARENAS=512" "PS_HEAVY_ARENAS=768" "PS_HEAVY_ARENAS=1024" "PS_HEAVY_ARENAS=1536" 2>&1 | tee gpurun_out/r2_ab_heavy.log
