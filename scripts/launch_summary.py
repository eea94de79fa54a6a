"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name']); v = float(row['Metric Value'].replace(',', ''))
    if row['Metric Unit'] == 'ns':
        v /= 1000.0
    agg[name][0] += 1; agg[name][1] += v; tot += v
print("total us %.1f launches %d" % (tot, sum(a[0] for a in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print("%-44s n=%4d total=%9.1f avg=%8.2f share=%5.1f%%" % (k.replace('void ', '').strip()[:44], n, t, t / n, 100 * t / tot))
