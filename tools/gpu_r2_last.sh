# Last check of the round on the final build: whole GPU suite, smoke(), the driver-style bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driver_style.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench_driver_style.json
