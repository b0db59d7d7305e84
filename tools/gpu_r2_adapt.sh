# Round 2: adaptive lane / warp boundary (13 contacts while the warp kernel has room, 17 once it is loaded): parity + A/B in three regimes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "free_running or c3_shape or dense_arena or scheduling" 2>&1 | tail -2
AB_WARMUP=5 AB_STEPS=20 bash tools/gpu_ab.sh "PS_X=0" "PS_WARP_MIN_CONTACTS=17" 2>&1 | tee gpurun_out/r2_ab_adapt.log
AB_WARMUP=35 AB_STEPS=20 bash tools/gpu_ab.sh "PS_X=0" "PS_WARP_MIN_CONTACTS=17" "PS_WARP_MIN_CONTACTS=13" 2>&1 | tee -a gpurun_out/r2_ab_adapt.log
AB_WARMUP=65 AB_STEPS=20 bash tools/gpu_ab.sh "PS_X=0" "PS_WARP_MIN_CONTACTS=17" "PS_WARP_MIN_CONTACTS=13" 2>&1 | tee -a gpurun_out/r2_ab_adapt.log
AB_STEPS=40 bash tools/gpu_ab.sh "PS_X=0" "PS_WARP_MIN_CONTACTS=17" 2>&1 | tee -a gpurun_out/r2_ab_adapt.log
python tools/gpu_island_census.py 4096 105 30 2>&1 | grep "per arena-tick" | tee -a gpurun_out/r2_ab_adapt.log
python tools/gpu_island_census.py 4096 195 30 2>&1 | grep "per arena-tick" | tee -a gpurun_out/r2_ab_adapt.log
