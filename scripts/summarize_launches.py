"""Summarise an ncu `--csv` launch list (long format: one row per launch x metric) per kernel:
launches, total time, share, time-weighted tensor-pipe activity and DRAM bytes.
    python scripts/summarize_launches.py gpurun_out/r1b_bench_launches.csv [--md]"""
import csv, sys, re, collections
path = sys.argv[1]
rows = []
with open(path, newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
per = collections.OrderedDict()
for r in rd:
    key = r['ID']
    d = per.setdefault(key, {'name': r['Kernel Name']})
    try:
        v = float(r['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    unit = r['Metric Unit']
    m = r['Metric Name']
    if m == 'gpu__time_duration.sum':
        v = v / 1e3 if unit in ('ns', 'nsecond') else (v * 1e3 if unit in ('ms', 'msecond') else v)   # -> us
    if m.startswith('dram__bytes'):
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
        v *= mult
    d[m] = v
agg = collections.OrderedDict()
for d in per.values():
    n = re.sub(r'\(.*$', '', d['name'])
    n = n.replace('void mxf::', '').replace('void at::native::', 'at::')
    a = agg.setdefault(n, [0, 0.0, 0.0, 0.0, 0.0])
    t = d.get('gpu__time_duration.sum', 0.0)
    a[0] += 1
    a[1] += t
    a[2] += t * d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0)
    a[3] += d.get('dram__bytes_read.sum', 0.0)
    a[4] += d.get('dram__bytes_write.sum', 0.0)
tot = sum(a[1] for a in agg.values())
print('| kernel | launches | total us | share | tensor pipe % (time-weighted) | DRAM rd MB | DRAM wr MB |')
print('|---|---:|---:|---:|---:|---:|---:|')
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print('| `%s` | %d | %.1f | %.1f%% | %.1f | %.1f | %.1f |' % (n[:70], a[0], a[1], 100 * a[1] / tot, a[2] / a[1] if a[1] else 0,
                                                          a[3] / 1e6, a[4] / 1e6))
print('total %.1f us over %d launches' % (tot, sum(a[0] for a in agg.values())))
