# Round 2 evidence, part 2 (its own call: two full ncu reports exceed the 64 MiB a call may bring back): full ncu capture of one think step right after the reset
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD3="python bench.py --steps 2 --warmup 12 --no-cpu-baseline --no-extras --no-parity --no-steady --no-e2e --arenas 1024"
PS_GROUPS=1 $CMD3 > gpurun_out/r02_plain3.log 2>&1 && \
PS_GROUPS=1 ncu --set full --clock-control none -k regex:'k_sense|k_localmap|k_think|k_phys_pre|k_island_big|k_island_warp|k_island_batch|k_phys_post' --launch-skip 216 --launch-count 8 -f -o gpurun_out/r02_prof_early $CMD3 > gpurun_out/r02_ncu_full_early.log 2>&1
echo "ncu early rc=$?"; ls -la gpurun_out/r02_prof_early.ncu-rep
