cd $GRAFT_REPO_ROOT
for cfg in "4 4" "8 2" "16 2" "32 2"; do
  set -- $cfg
  echo "groups=$1 worker_blocks_per_sm=$2"; PS_GROUPS=$1 PS_WORKER_BLOCKS_PER_SM=$2 PS_BATCH_BLOCKS_PER_SM=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['world_step_stage_ms_per_tick'], d['roofline']['kernel_ms_total']['sense']/d['roofline']['kernel_launch_counts']['sense'])"
done
