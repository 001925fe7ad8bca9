#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: agg_launches.py file.csv [last_n_launches]"""
import collections, csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
    if row.get('Metric Name') == 'gpu__time_duration.sum':
        rows.append((int(row['ID']), row['Kernel Name'], float(row['Metric Value'].replace(',', ''))))
if len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
tot = sum(x[2] for x in rows)
print("launches %d total %.1f us" % (len(rows), tot / 1000))
agg = collections.OrderedDict()
for i, name, t in rows:
    short = re.sub(r'\(.*', '', name)[:72]
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += t
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-74s %4d %9.1f us %5.1f%%" % (k, c, t / 1000, 100 * t / tot))
