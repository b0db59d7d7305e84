# A/B of engine builds / environments (run on the GPU box): every argument is one variant's env assignments.
# AB_WARMUP / AB_STEPS choose the regime (default: steady state, 100 warm-up + 60 timed think steps), device-resident only.
cd $GRAFT_REPO_ROOT
run() { python bench.py --steps ${AB_STEPS:-60} --warmup ${AB_WARMUP:-100} --no-cpu-baseline --no-extras --no-parity --no-steady --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in r['world_step_stage_ms_per_tick'].items()}, {k: round(r['kernel_ms_total'][k]/max(1,r['kernel_launch_counts'][k]),2) for k in ('sense','think','physics','isl_big','isl_warp','isl_batch','isl_lanex')})"; }
for e in "$@"; do
  echo "variant: $e"; env $e bash -c "$(declare -f run); run"
done
