cd $GRAFT_REPO_ROOT
if [ -z "$NOTEST" ]; then python -m pytest tests -m gpu -q -x 2>&1 | tail -2; fi
for cfg in "20 5" "40 100"; do
  set -- $cfg
  echo "steps=$1 warmup=$2"; python bench.py --steps $1 --warmup $2 --no-cpu-baseline --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['world_step_stage_ms_per_tick'], {k: r['kernel_ms_total'][k]/max(1,r['kernel_launch_counts'][k]) for k in ('sense','think','physics')})"
done
