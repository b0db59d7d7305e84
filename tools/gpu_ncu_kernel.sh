# One full ncu capture of the kernels matching $1 (after a plain run of the same command), one arena group of 1024 arenas.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
K=${1:-k_island_warp}; SKIP=${2:-12}; CNT=${3:-2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --arenas 1024"
PS_GROUPS=1 $CMD > gpurun_out/plain_k.log 2>&1 || { tail -5 gpurun_out/plain_k.log; exit 1; }
PS_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c $CNT -f -o gpurun_out/prof_k $CMD > gpurun_out/ncu_k.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_k.log
