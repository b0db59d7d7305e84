cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['also'], d['cpu_baseline'])
print({k:v for k,v in d['roofline'].items() if k in ('achieved','frac','traffic','algorithmic_bytes_per_arena_tick')})
print(d['roofline_sense'])
"
