#!/bin/bash
# End-of-round verification on one GPU: GPU test tier, smoke, and every bench workload (JSON lines -> gpurun_out/final/).
O=gpurun_out/final
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/summary.txt
tail -3 $O/pytest_gpu.log >> $O/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
tail -1 $O/smoke.log >> $O/summary.txt
timeout 300 python bench.py --out $O/bench_headline.json > /dev/null 2> $O/bench_headline.err; echo "headline rc=$?" >> $O/summary.txt
for w in c1 c2 c3 c4 c5; do
  timeout 300 python bench.py --workload $w --out $O/bench_$w.json > /dev/null 2> $O/bench_$w.err; echo "$w rc=$?" >> $O/summary.txt
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 --out $O/bench_reference.json > /dev/null 2> $O/bench_reference.err; echo "reference rc=$?" >> $O/summary.txt
python - <<'PY' >> $O/summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/final/bench_*.json')):
    d = json.load(open(f))
    print(f.split('/')[-1], 'value %.1f' % d.get('value', float('nan')), 'ms/step %.4f' % d.get('ms_per_step', float('nan')),
          'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'),
          'roofline', (d.get('roofline') or {}).get('frac'), 'no_flush', d.get('value_no_flush'))
PY
cat $O/summary.txt
