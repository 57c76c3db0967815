"""SASS evidence of the Blackwell-native paths: per kernel of libdeeplip_b200.so, counts of the tcgen05 / TMEM / TMA
mnemonics and a few lines around the first site of each.   python tools/sass_excerpt.py > profiles/rXX_sass_excerpts.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'deeplip_b200', 'libdeeplip_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
dem = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
PAT = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTMAPF', 'SYNCS', 'HMMA', 'UTCATOMSWS']
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
    elif cur is not None:
        kernels[cur].append(line)
commit = subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
print('# SASS excerpts of libdeeplip_b200.so (built at commit %s, `cuobjdump -sass`, sm_100a)\n' % commit)
print('`tcgen05.mma` shows as `UTCHMMA` (`.2CTA` for cta_group::2), `tcgen05.ld/st` as `LDTM`/`STTM`, TMA as `UTMALDG` / `UTMASTG`, '
      '`tcgen05.commit` as `UTCBAR`; there is no legacy `HMMA` anywhere.\n')
print('| kernel | ' + ' | '.join(PAT) + ' |')
print('|---|' + '---|' * len(PAT))
first = {}
for k, lines in kernels.items():
    cnt = {p: 0 for p in PAT}
    for i, l in enumerate(lines):
        for p in PAT:
            if re.search(r'\b' + p + r'\b', l) or (' ' + p + '.') in l:
                cnt[p] += 1
                first.setdefault((k, p), i)
    if cnt['UTCHMMA'] or cnt['UTMALDG'] or cnt['UTMASTG'] or cnt['LDTM']:
        name = re.sub(r'\(.*', '', dem(k)).replace('void ', '').replace('dl::', '')
        print('| `%s` | ' % name[:80] + ' | '.join(str(cnt[p]) for p in PAT) + ' |')
print()


def excerpt(kpat, p, before=2, after=5):
    for (k, pp), i in first.items():
        if pp == p and kpat in dem(k):
            name = re.sub(r'\(.*', '', dem(k)).replace('void ', '').replace('dl::', '')
            print('### first `%s` site in `%s`\n```' % (p, name[:90]))
            for l in kernels[k][max(0, i - before):i + after]:
                l = re.sub(r'/\* 0x[0-9a-f]+ \*/', '', l).rstrip()
                if l.strip():
                    print(re.sub(r'^\s+', '  ', l)[:150])
            print('```\n')
            return


excerpt('igemm2_conv_kernel<256, true, 1, 2, true>', 'UTCHMMA', 3, 12)
excerpt('igemm2_conv_kernel<256, true, 1, 2, true>', 'UTMASTG')
excerpt('igemm2_conv_kernel<256, false, 1', 'UTMALDG', 2, 6)
excerpt('igemm2_conv_kernel<128, true, 1, 1, false>', 'LDTM', 1, 4)
excerpt('conv3x3_halo_kernel', 'UTCHMMA', 2, 10)
excerpt('stem_conv3d_kernel', 'UTCHMMA', 2, 8)
excerpt('stem_conv3d_kernel', 'STTM', 1, 3)
