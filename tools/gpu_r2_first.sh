# Round 2, first GPU run: the GPU parity suite (with the new steady-state and 1000-episode tests), the driver-style bench
# line (new also.steady_state / parity_check / external_controller / config4), and the sensing-stage phase split
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
(time python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_first.json 2> gpurun_out/r2_bench_first.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench_first.json; tail -5 gpurun_out/r2_bench_first.err
PS_ENGINE_LIB=$GRAFT_REPO_ROOT/puckswarm_b200/csrc/libpuckswarm_b200_diag.so python tools/gpu_phase_stats.py 4096 > gpurun_out/r2_phase_stats.log 2>&1; grep -E "^sense|findnew|syncfix|toi_cycles|cycles per|robot-steps" gpurun_out/r2_phase_stats.log
