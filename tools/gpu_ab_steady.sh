# A/B of engine builds / environments in the steady state (run on the GPU box): every further argument is one variant's env assignments
cd $GRAFT_REPO_ROOT
run() { python bench.py --steps ${AB_STEPS:-60} --warmup 100 --no-cpu-baseline --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), r['world_step_stage_ms_per_tick'], {k: round(r['kernel_ms_total'][k]/max(1,r['kernel_launch_counts'][k]),2) for k in ('sense','think','physics')})"; }
for e in "$@"; do
  echo "variant: $e"; env $e bash -c "$(declare -f run); run"
done
