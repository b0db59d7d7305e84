# Round 2: sense stage after the revert (A/B), FAST-mode parity test and ensemble gate
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fast_mode or camera or c3_shape_against" 2>&1 | tail -3
bash tools/gpu_ab.sh "PS_X=0" 2>&1 | tee gpurun_out/r2_ab_sense2.log
timeout 900 python tools/gpu_fast_mode_gate.py 1000 3000 10 > gpurun_out/r02_fast_mode_gate.json 2> gpurun_out/fast_gate.err; tail -3 gpurun_out/fast_gate.err; cat gpurun_out/r02_fast_mode_gate.json
