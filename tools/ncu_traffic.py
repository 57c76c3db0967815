"""Per-launch duration + DRAM bytes of one extraction step from an ncu CSV log
(`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`), as text + a json of the trunk's traffic
that bench.py cites as `roofline.traffic`.
    python tools/ncu_traffic.py gpurun_out/step_traffic.csv profiles/rXX_step_traffic.txt profiles/rXX_trunk_traffic.json <steps> <commit>"""
import collections, csv, io, json, sys

src, out_txt, out_json, steps, commit = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
lines = open(src, errors='replace').read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))
byid = collections.OrderedDict()
for r in rows:
    d = byid.setdefault(r['ID'], {'name': r['Kernel Name']})
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    if r['Metric Name'] == 'gpu__time_duration.sum':
        d['us'] = v / 1e3 if u.startswith('n') else (v if u.startswith('u') else v * 1e3)
    else:
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        d['rd' if 'read' in r['Metric Name'] else 'wr'] = v * mult


def short(n):
    n = n.split('(')[0].replace('void ', '').replace('dl::', '')
    return n[:72]


launches = [d for d in byid.values() if 'us' in d]
per_step = len(launches) // steps
last = launches[-per_step:]                       # the last step: warm caches of packed weights, steady allocator
i0 = next(i for i, d in enumerate(last) if 'stem_conv3d' in d['name'] or 'stem2_conv3d' in d['name'])
i1 = next(i for i, d in enumerate(last) if 'frame_pool' in d['name'] or 'temporal_mean' in d['name'])
trunk = last[i0 + 1:i1]                        # the ResNet-18 body: everything between the stem and the frame pool
txt = ['one AV extraction step, B=64, commit %s: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,'
       'dram__bytes_write.sum --clock-control none, per launch (last of %d steps)' % (commit, steps), '']
for d in last:
    txt.append('%8.1f us %8.1f MB read %8.1f MB write  %s' % (d['us'], d.get('rd', 0) / 1e6, d.get('wr', 0) / 1e6, short(d['name'])))
tb = sum(d.get('rd', 0) + d.get('wr', 0) for d in trunk)
txt += ['', 'launches per step: %d; sum of durations %.1f us' % (per_step, sum(d['us'] for d in last)),
        'trunk conv launches (between the stem and the frame pool): %d, %.1f us, %.3f GB of DRAM traffic' %
        (len(trunk), sum(d['us'] for d in trunk), tb / 1e9)]
open(out_txt, 'w').write('\n'.join(txt) + '\n')
json.dump({'trunk_dram_bytes_per_step': tb, 'trunk_conv_launches': len(trunk), 'batch': 64, 'source': out_txt,
           'commit': commit, 'trunk_us_under_ncu': sum(d['us'] for d in trunk)}, open(out_json, 'w'))
print('\n'.join(txt[-3:]))
