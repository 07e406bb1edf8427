"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name count, total
time and share.  usage: python tools/summarize_ncu.py launches.csv [skip_first_n] > summary.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = OrderedDict()
n = 0
for r in rows[1:]:
    if r[ki] == 'Kernel Name':
        continue
    n += 1
    if n <= skip:
        continue
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').strip()
    name = re.sub(r'<.*', '', name) if not name.startswith('k_') else name
    t = float(r[vi].replace(',', ''))
    t = t / 1e3 if r[ui].startswith('ns') or r[ui] == 'nsecond' else (t if r[ui].startswith('us') else t * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print('kernel,launches,total_us,share')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%s,%d,%.1f,%.4f' % (k, a[0], a[1], a[1] / tot))
print('TOTAL,%d,%.1f,1.0' % (sum(a[0] for a in agg.values()), tot))
