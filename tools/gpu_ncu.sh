# One ncu --set full capture of the hot kernels on a reduced batch (A arenas), after the same command ran clean without ncu.
set -x
cd $GRAFT_REPO_ROOT
A=${1:-1024}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --arenas $A"
$CMD > gpurun_out/plain_ncu.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_sense|k_localmap|k_think|k_phys_pre|k_phys_post' -s 10 -c 10 -f -o gpurun_out/prof2 $CMD > gpurun_out/ncu_full2.log 2>&1
echo "ncu rc=$?"
tail -5 gpurun_out/ncu_full2.log
