"""Pivot an `ncu --metrics ... --csv --log-file` dump: one row per launch, one column per metric.  usage: ncu_metrics_table.py file.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
tab = collections.OrderedDict()
for r in rows[1:]:
    key = (r[ix['ID']], r[ix['Kernel Name']][:60])
    tab.setdefault(key, {})[r[ix['Metric Name']]] = (r[ix['Metric Value']], r[ix['Metric Unit']])
for key, m in tab.items():
    print(key)
    for k, (v, u) in m.items():
        print(f'    {k:90s} {v:>18s} {u}')
