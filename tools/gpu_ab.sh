cd $GRAFT_REPO_ROOT
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['world_step_stage_ms_per_tick'], {k: r['kernel_ms_total'][k]/max(1,r['kernel_launch_counts'][k]) for k in ('sense','think','physics')})"; }
echo "A: $A_ENV"; env $A_ENV bash -c "$(declare -f run); run"
echo "B: $B_ENV"; env $B_ENV bash -c "$(declare -f run); run"
echo "A: $A_ENV"; env $A_ENV bash -c "$(declare -f run); run"
echo "B: $B_ENV"; env $B_ENV bash -c "$(declare -f run); run"
