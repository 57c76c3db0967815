"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the
per-launch durations in order.   python tools/ncu_launches.py gpurun_out/launches.csv [out.md]"""
import csv, io, sys, collections

def load(path):
    lines = open(path, errors='replace').read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))

def short(name):
    n = name.split('(')[0]
    for pre in ('void ', 'dl::'):
        n = n.replace(pre, '')
    return n[:70]

rows = [r for r in load(sys.argv[1]) if r.get('Metric Name') == 'gpu__time_duration.sum']
def us(r):
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    return v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
tot = collections.OrderedDict()
for r in rows:
    k = short(r['Kernel Name'])
    t = tot.setdefault(k, [0, 0.0])
    t[0] += 1; t[1] += us(r)
out = ['| kernel | launches | total us | avg us | share |', '|---|---|---|---|---|']
total = sum(v[1] for v in tot.values())
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    out.append('| %s | %d | %.1f | %.1f | %.1f%% |' % (k, n, t, t / n, 100 * t / total))
out.append('')
out.append('per-launch (first 120 launches of dl kernels), us: grid | block | name')
n = 0
for r in rows:
    k = short(r['Kernel Name'])
    if 'kernel' in k and ('igemm' in k or 'stem' in k or 'frontend' in k or 'pool' in k or 'znorm' in k or 'cosine' in k):
        out.append('%8.1f  %s %s  %s' % (us(r), r.get('Grid Size', ''), r.get('Block Size', ''), k))
        n += 1
        if n >= 120:
            break
text = '\n'.join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(text + '\n')
