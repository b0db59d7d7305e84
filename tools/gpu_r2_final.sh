# Round 2 evidence run (one B200): GPU parity suite, the driver-style bench line, the default (steady-state) bench line, the
# reference arm, the launch list of a steady-state think step and one full ncu capture of every kernel of a steady-state think
# step (1024 arenas, one arena group; each ncu run after a plain run of the same command). Part 2: tools/gpu_r2_final_early.sh.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driver_style.json 2> gpurun_out/r02_bench.err; echo "bench(driver style) rc=$?"; cut -c1-160 gpurun_out/r02_bench_driver_style.json
python bench.py --no-extras > gpurun_out/r02_bench_steady.json 2>> gpurun_out/r02_bench.err; echo "bench(steady) rc=$?"; cut -c1-160 gpurun_out/r02_bench_steady.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err; echo "ref rc=$?"; cut -c1-160 gpurun_out/r02_bench_reference.json
CMD="python bench.py --steps 2 --warmup 100 --no-cpu-baseline --no-extras --no-parity --no-steady --no-e2e"
$CMD > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1750 --launch-count 150 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
echo "ncu list rc=$?"
CMD2="python bench.py --steps 2 --warmup 100 --no-cpu-baseline --no-extras --no-parity --no-steady --no-e2e --arenas 1024"
PS_GROUPS=1 $CMD2 > gpurun_out/r02_plain2.log 2>&1 && \
PS_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:'k_sense|k_localmap|k_think|k_phys_pre|k_island_big|k_island_warp|k_island_batch|k_phys_post' --launch-skip 1800 --launch-count 8 -f -o gpurun_out/r02_prof $CMD2 > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r02_prof.ncu-rep
