cd $GRAFT_REPO_ROOT
for g in 4 16 32; do
  echo "groups=$g"; PS_GROUPS=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['world_step_stage_ms_per_tick'])"
done
