# Launch list of one arena group (1024 arenas) after a plain run: per-kernel mean / max duration (ncu gpu__time_duration).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --arenas 1024"
PS_GROUPS=1 $CMD > gpurun_out/plain_ll.log 2>&1 || { tail -5 gpurun_out/plain_ll.log; exit 1; }
PS_GROUPS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1g.csv $CMD > gpurun_out/ncu_ll.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/launches_1g.csv') if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size'); ui = h.index('Metric Unit')
d = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1e3
    elif r[ui] == 'ms': v *= 1e3
    d.setdefault((r[ki][:40], r[gi]), []).append(v)
b=[v for k,v in d.items() if 'k_island_batch' in k[0]]
if b:
    b=b[0]; n=len(b)//4*4; b=b[n//2:n]
    for j,name in enumerate(['bucket4 (13-15)','bucket3 (6-12)','bucket2 (3-5)','bucket1 (1-2)']):
        x=b[j::4]; print('   k_island_batch %-16s mean %9.1f us max %9.1f us' % (name, sum(x)/len(x), max(x)))
for k, v in d.items():
    v2 = v[len(v)//2:]   # second half: after warm-up
    print('%-42s %-14s n=%3d mean %9.1f us  max %9.1f us' % (k[0], k[1], len(v2), sum(v2)/len(v2), max(v2)))
PY
