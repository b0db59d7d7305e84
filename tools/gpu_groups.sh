cd $GRAFT_REPO_ROOT
for cfg in "4 0" "4 1" "6 1" "8 1" "12 1"; do
  set -- $cfg
  echo "groups=$1 merge=$2"; PS_GROUPS=$1 PS_MERGE_BATCH=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], r['world_step_stage_ms_per_tick'], {k: r['kernel_ms_total'][k]/max(1,r['kernel_launch_counts'][k]) for k in ('sense','think','physics')})"
done
