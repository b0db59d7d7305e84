# Runs on the GPU box: GPU parity tests, the bench lines (steady state = default, early regime, reference arm), the launch
# list of two timed think steps of the default configuration, and one full ncu capture of a whole think step in the steady
# state (1024 arenas, one arena group; 8 launches = one of each kernel, the report must stay under the 64 MiB copy-back limit), each ncu run after a plain run of the same command.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_early.json 2>> gpurun_out/bench.err; echo "early rc=$?"; cut -c1-200 gpurun_out/bench_early.json
python bench.py --impl reference > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
CMD="python bench.py --steps 2 --warmup 100 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1750 --launch-count 150 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
CMD2="python bench.py --steps 2 --warmup 100 --no-cpu-baseline --no-extras --arenas 1024"
PS_GROUPS=1 $CMD2 > gpurun_out/plain2.log 2>&1 && \
PS_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:'k_sense|k_localmap|k_think|k_phys_pre|k_island_big|k_island_warp|k_island_batch|k_phys_post' --launch-skip 1800 --launch-count 8 -f -o gpurun_out/prof_final $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
