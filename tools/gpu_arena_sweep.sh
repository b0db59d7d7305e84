# Arenas-per-GPU sweep (one B200, one arena group): the island chains are amortised over the batch
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { python bench.py --arenas $1 --steps $3 --warmup $2 --no-cpu-baseline --no-extras --no-parity --no-steady --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('arenas', $1, 'warmup', $2, 'value', round(d['value']), 'ms_per_step', round(d['ms_per_step'],2), {k: round(v,2) for k,v in r['world_step_stage_ms_per_tick'].items()}, {k: round(r['kernel_ms_total'][k]/max(1,r['kernel_launch_counts'][k]),2) for k in ('sense','think','isl_big','isl_warp','isl_batch')})"; }
for a in 1024 2048 4096 8192 16384 32768; do run $a 5 20; done 2>&1 | tee gpurun_out/r2_arena_sweep.log
for a in 1024 2048 4096 8192 16384 32768; do run $a 100 20; done 2>&1 | tee -a gpurun_out/r2_arena_sweep.log
