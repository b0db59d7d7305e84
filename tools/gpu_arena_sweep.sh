cd $GRAFT_REPO_ROOT
for cfg in "4096 1" "16384 1"; do
  set -- $cfg
  echo "arenas=$1 groups=$2"; PS_GROUPS=$2 python bench.py --arenas $1 --steps 10 --warmup 4 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], r['world_step_stage_ms_per_tick'], {k: r['kernel_ms_total'][k]/max(1,r['kernel_launch_counts'][k]) for k in r['kernel_ms_total']})"
done
