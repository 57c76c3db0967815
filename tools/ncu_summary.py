"""Key metrics + hottest SASS lines of one ncu report.  python tools/ncu_summary.py rep.ncu-rep [out.md] [launch index in the report]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hdr, units, vals = rows[0], rows[1], rows[2 + which]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__cycles_elapsed.max',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']
out = ['## %s' % vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '## kernel', '']
for h, u, v in zip(hdr, units, vals):
    if h in want:
        out.append('- `%s` = %s %s' % (h, v, u))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(which), '--launch-count', '1'], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
sh = srows[1]; data = srows[2:]
ix = {h: i for i, h in enumerate(sh)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data) or 1
stalls = [h for h in sh if h.startswith('stall_') and 'Not Issued' not in h]
agg = sorted(((s, sum(f(r, s) for r in data)) for s in stalls), key=lambda kv: -kv[1])[:8]
out += ['', 'stall reasons (all warps, samples): ' + ', '.join('%s %.0f%%' % (s[6:], 100 * v / tot) for s, v in agg), '',
        'hottest SASS (share of samples, top stalls):', '```']
for i, r in sorted(enumerate(data), key=lambda ir: -f(ir[1], '# Samples'))[:28]:
    st = sorted(((s[6:], f(r, s)) for s in stalls), key=lambda kv: -kv[1])[:2]
    out.append('%5d %5.1f%%  %-64s %s' % (i, 100 * f(r, '# Samples') / tot, r[ix['Source']].strip()[:64],
                                       ' '.join('%s=%d' % (a, b) for a, b in st)))
out.append('```')
text = '\n'.join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], 'a').write(text + '\n\n')
