# Runs on the GPU box: one full ncu capture in the steady state (after 100 think steps): every kernel of one tick plus one
# sense / think, one arena group of 1024 arenas, after a plain run of the same command.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD2="python bench.py --steps 2 --warmup 100 --no-cpu-baseline --no-extras --arenas 1024"
PS_GROUPS=1 $CMD2 > gpurun_out/plain2.log 2>&1 && \
PS_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:'k_sense|k_localmap|k_think|k_phys_pre|k_island_big|k_island_warp|k_island_batch|k_phys_post' --launch-skip 1800 --launch-count 9 -f -o gpurun_out/prof_final $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
