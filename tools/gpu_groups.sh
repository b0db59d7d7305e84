cd $GRAFT_REPO_ROOT
for g in 1 2 3 4 6; do
  echo "groups=$g"; PS_GROUPS=$g python bench.py --steps 40 --warmup 100 --no-cpu-baseline --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], r['world_step_stage_ms_per_tick'], {k: r['kernel_ms_total'][k]/max(1,r['kernel_launch_counts'][k]) for k in ('sense','think','physics')})"
done
