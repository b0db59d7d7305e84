cd $GRAFT_REPO_ROOT
for n in 2 4 8; do
echo "handles=$n"; PS_E2E_HANDLES=$n python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
