"""Digest of an ncu report (run here, no GPU): headline counters per kernel from the raw page and the
instruction mix / hottest instructions from the source page.  python tools/ncu_digest.py <rep> [n_hot]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
n_hot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'smsp__inst_issued.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_active.avg.per_cycle_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get('Kernel Name', '')[:100])
    for k in keys:
        if k in d: print('   %-70s %s' % (k, d[k]))
    st = [(float(d[k].replace(',', '')), k) for k in hdr if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and d[k]]
    print('   stalls per issue: ' + ', '.join('%s %.2f' % (k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), v) for v, k in sorted(st, reverse=True) if v > 0.04))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# the source page holds one table per kernel: header lines start with "Address"
tables, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; tables.append(cur)
    elif r and r[0] == 'Address' and cur is not None: cur['hdr'] = r
    elif cur is not None and 'hdr' in cur and len(r) >= len(cur['hdr']): cur['rows'].append(r)
for tb in tables:
    h = tb['hdr']; ia, isrc, iex, ism = h.index('Address'), h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    byop, samp, tot, ts = collections.Counter(), collections.Counter(), 0, 0
    ins = []
    for r in tb['rows']:
        s = r[isrc].strip(); ex = int(r[iex]); sm = int(r[ism])
        parts = s.split()
        op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
        byop[op] += ex; samp[op] += sm; tot += ex; ts += sm
        ins.append((int(r[ia], 16), s, ex, sm))
    print('\n== %s\n   %d warp instructions, %d samples' % (tb['name'][:100], tot, ts))
    print('   ' + '  '.join('%s %.1f%%(%.1f%%s)' % (op, 100.0 * c / tot, 100.0 * samp[op] / max(ts, 1)) for op, c in byop.most_common(16)))
    base = ins[0][0]
    for a, s, ex, sm in sorted(ins, key=lambda t: -t[3])[:n_hot]:
        print('   %5x %8d %6d  %s' % (a - base, ex, sm, s[:80]))
