#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict(); tot = 0.0; n = 0
for row in csv.DictReader(lines):
    name = re.sub(r'^.*::', '', re.sub(r'\(.*', '', row['Kernel Name']))
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1e3 if u == 'us' else v / 1e6 if u == 'ns' else v * 1e3 if u == 's' else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v; n += 1
print("launches %d  total %.3f ms" % (n, tot))
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-44s n=%4d %10.3f ms %5.1f%%" % (k, c, v, 100 * v / tot))
