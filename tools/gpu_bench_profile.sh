# Runs on the GPU box: the bench lines (default = steady state, early regime, reference arm), then the launch list
# (ncu gpu__time_duration) of the timed region of the same command with fewer timed steps.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-260 gpurun_out/bench.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_early.json 2>> gpurun_out/bench.err; echo "early rc=$?"
cut -c1-260 gpurun_out/bench_early.json
python bench.py --impl reference > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
cut -c1-260 gpurun_out/bench_reference.json
CMD="python bench.py --steps 2 --warmup 100 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 7300 --launch-count 500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
