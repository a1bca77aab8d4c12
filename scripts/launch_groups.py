"""Launch list grouped by (kernel, grid): usage launch_groups.py launches.csv [min_us_to_list]"""
import csv, re, sys
from collections import defaultdict
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = []
for r in csv.DictReader(lines):
    if r['Metric Name'] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('p4c::', '')
    rows.append((int(r['ID']), name, r['Grid Size'], float(r['Metric Value'].replace(',', '')) / 1e3))
agg = defaultdict(lambda: [0, 0.0])
for i, n, g, t in rows:
    agg[(n, g)][0] += 1
    agg[(n, g)][1] += t
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot/1e3:.3f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1]/1e3:8.3f} ms {v[0]:5d} x {v[1]/v[0]:9.2f} us  {k[0][:44]:44s} grid {k[1]}")
if len(sys.argv) > 2:
    th = float(sys.argv[2])
    for i, n, g, t in rows:
        if t > th:
            print(i, n[:45], g, round(t, 1))
