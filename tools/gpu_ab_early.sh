# A/B of engine environments in the early regime (ticks 15-75; run on the GPU box): every argument is one variant's env assignments
cd $GRAFT_REPO_ROOT
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), r['world_step_stage_ms_per_tick'])"; }
for e in "$@"; do
  echo "variant: $e"; env $e bash -c "$(declare -f run); run"
done
