"""Compare two tools/time_mixedops.py JSON outputs per shape for selected kernels: python tools/cmp_kernels.py old.json new.json [names]"""
import json, sys
a = json.load(open(sys.argv[1])); b = json.load(open(sys.argv[2]))
names = sys.argv[3].split(',') if len(sys.argv) > 3 else ['expand', 'project', 'dc', 'dx']
def kmap(r):
    ks = r.get('kernels', r.get('rows', []))
    return {k['name']: k['ms'] / max(k.get('launches', 1), 1) * (2 if k['name'] == 'um_prep_w' else 1) for k in ks}
print('%-40s' % 'shape' + ''.join('%22s' % n for n in names))
for ra, rb in zip(a['rows'], b['rows']):
    ka, kb = kmap(ra), kmap(rb)
    print('%-40s' % ra.get('block', '?')[:40] + ''.join('%10.3f ->%9.3f' % (ka.get(n, 0), kb.get(n, 0)) for n in names))
