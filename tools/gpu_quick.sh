set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python tools/gpu_phase_stats.py 4096
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json; cat gpurun_out/bench_quick.json
