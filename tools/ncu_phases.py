"""Split one kernel's SASS into phases at marker instructions and report instructions executed and
stall samples per phase:  python tools/ncu_phases.py report.ncu-rep <kernel regex>"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat,
                      '--launch-skip', (sys.argv[3] if len(sys.argv) > 3 else '0'), '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO('\n'.join(raw.splitlines()[1:]))))
h = rows[0]
isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
data = []
for r in rows[1:]:
    try:
        data.append((r[isrc], int(r[isamp] or 0), int(r[iex] or 0), r))
    except (ValueError, IndexError):
        pass
data = data[:len(data) // 2] if len(data) > 9000 else data   # report lists the kernel twice
tot_s = sum(d[1] for d in data); tot_e = sum(d[2] for d in data)
# phases: cut at BAR.SYNC, first/last SHFL, LDGSTS
cuts = [0]
prev_kind = None
for i, (src, s, e, r) in enumerate(data):
    kind = None
    if 'BAR.SYNC' in src: kind = 'BAR'
    elif src.startswith('SHFL') or ' SHFL' in src[:12]: kind = 'SHFL'
    elif 'LDGSTS' in src: kind = 'LDGSTS'
    elif 'WARPSYNC' in src: kind = 'WS'
    if kind in ('BAR',) or (kind == 'SHFL' and prev_kind != 'SHFL') or (kind == 'LDGSTS' and prev_kind != 'LDGSTS'):
        cuts.append(i)
    if kind in ('SHFL', 'LDGSTS', 'BAR'):
        prev_kind = kind
cuts.append(len(data))
print('total executed %d samples %d' % (tot_e, tot_s))
for a, b in zip(cuts[:-1], cuts[1:]):
    if b - a < 8: continue
    e = sum(d[2] for d in data[a:b]); s = sum(d[1] for d in data[a:b])
    if s < 0.004 * tot_s: continue
    st = sorted(((sum(int(d[3][i] or 0) for d in data[a:b]), n) for i, n in stall_cols), reverse=True)[:4]
    print('[%5d,%5d) first=%-34s exec %5.1f%% samples %5.1f%%  %s' % (a, b, data[a][0][:34], 100.0 * e / tot_e, 100.0 * s / tot_s,
          ' '.join('%s=%.0f%%' % (n.replace('stall_', ''), 100.0 * v / max(1, s)) for v, n in st)))
