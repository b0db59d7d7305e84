# Runs on the GPU box: bench line (+ reference arm), launch list (ncu gpu__time_duration) and one full ncu capture of the kernels.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-400
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --arenas 1024"
PS_GROUPS=1 $CMD2 > gpurun_out/plain2.log 2>&1 && \
PS_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:'k_sense|k_localmap|k_think|k_phys_pre|k_island_warp|k_island_batch|k_phys_post' -s 16 -c 10 -f -o gpurun_out/prof_final $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
