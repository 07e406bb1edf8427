"""Per-launch DRAM traffic of the library's kernels from an ncu CSV holding dram__bytes_read.sum / dram__bytes_write.sum
(one search unit of `bench.py --profile-only`), keyed by the profiler-scope names bench.py uses for its roofline line.

    python tools/ncu_traffic.py dram.csv > profiles/ncu_traffic.json
Values: mean (read + write) bytes per launch of that kernel over the captured launches."""
import csv
import json
import re
import sys
from collections import OrderedDict

NAMES = [(r'k_ws_expand|k_um_expand', 'expand'), (r'k_ws_project|k_um_project', 'project'), (r'k_ws_dc|k_um_dc', 'dc'),
         (r'k_ws_dx|k_um_dx', 'dx'), (r'k_um_wgrad<\(int\)0|k_um_wgrad<0', 'wgrad_w3'), (r'k_um_wgrad<\(int\)1|k_um_wgrad<1', 'wgrad_w1'),
         (r'k_um_wgrad<\(int\)2|k_um_wgrad<2', 'xcov'), (r'k_dws_fwd<\(int\)3|k_dws_fwd<3|k_dw_fwd<\(int\)3|k_dw_fwd<3', 'dw_fwd_k3'),
         (r'k_dws_fwd<\(int\)5|k_dws_fwd<5|k_dw_fwd<\(int\)5|k_dw_fwd<5', 'dw_fwd_k5'),
         (r'k_dws_bwd<\(int\)3|k_dws_bwd<3|k_dw_bwd<\(int\)3|k_dw_bwd<3', 'dw_bwd_k3'),
         (r'k_dws_bwd<\(int\)5|k_dws_bwd<5|k_dw_bwd<\(int\)5|k_dw_bwd<5', 'dw_bwd_k5'), (r'k_dxfin', 'dxfin'), (r'k_b2b', 'b2b')]


def scope(kernel):
    for rx, n in NAMES:
        if re.search(rx, kernel):
            return n
    return None


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ui, idi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('ID')
    mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    per = OrderedDict()
    for r in rows[1:]:
        if r[ki] == 'Kernel Name' or not r[mi].startswith('dram__bytes_'):
            continue
        n = scope(r[ki])
        if n is None:
            continue
        d = per.setdefault(n, {})
        d[r[idi]] = d.get(r[idi], 0.0) + float(r[vi].replace(',', '')) * mult.get(r[ui], 1.0)
    out = OrderedDict((n, sum(d.values()) / len(d)) for n, d in per.items() if d)
    out['_launches'] = {n: len(d) for n, d in per.items()}
    out['_source'] = 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one search unit of bench.py --profile-only'
    json.dump(out, sys.stdout, indent=1)


if __name__ == '__main__':
    main()
