# Launch list in the steady state (after 100 think steps = 300 ticks), one arena group of 1024 arenas.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 100 --no-cpu-baseline --no-extras --arenas ${ARENAS:-1024}"
PS_GROUPS=1 $CMD > gpurun_out/plain_lls.log 2>&1 || { tail -5 gpurun_out/plain_lls.log; exit 1; }
PS_GROUPS=1 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1801 --launch-count 54 --csv --log-file gpurun_out/launches_steady.csv $CMD > gpurun_out/ncu_lls.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/launches_steady.csv') if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size'); ui = h.index('Metric Unit')
d = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1e3
    elif r[ui] == 'ms': v *= 1e3
    d.setdefault((r[ki][:40], r[gi]), []).append(v)
for k, v in d.items():
    print('%-42s %-14s n=%3d mean %9.1f us  max %9.1f us' % (k[0], k[1], len(v), sum(v)/len(v), max(v)))
PY
