cd $GRAFT_REPO_ROOT
for w in 6 9 12 16; do
  echo "warp_min_contacts=$w"; PS_WARP_MIN_CONTACTS=$w python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], r['world_step_stage_ms_per_tick'])"
done
