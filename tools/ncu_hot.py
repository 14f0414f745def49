"""Hot spots of one kernel from an ncu report (SASS view):
   python tools/ncu_hot.py report.ncu-rep <kernel regex> [top] [launch-skip]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
skip = sys.argv[4] if len(sys.argv) > 4 else '0'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat,
                      '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
lines = raw.splitlines()
rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
h = rows[0]
ia, isrc, isamp, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
data = []
for r in rows[1:]:
    if len(r) <= iex:
        continue
    try:
        data.append((int(r[isamp] or 0), int(r[iex] or 0), r[isrc], r, len(data)))
    except ValueError:
        pass
tot_s = sum(d[0] for d in data)
tot_e = sum(d[1] for d in data)
print('instructions (static)', len(data), 'executed', tot_e, 'samples', tot_s)
print('--- stall totals')
for i, n in stall_cols:
    v = sum(int(d[3][i] or 0) for d in data)
    if v:
        print('  %-28s %6.2f%%' % (n, 100.0 * v / max(1, tot_s)))
print('--- top instructions by samples')
for s, e, src, r, idx in sorted(data, reverse=True)[:top]:
    why = max(stall_cols, key=lambda c: int(r[c[0]] or 0))[1]
    print('  #%5d %5.2f%%  exec %9d  %-12s %s' % (idx, 100.0 * s / max(1, tot_s), e, why, src[:90]))
print('--- opcode mix by executed instructions')
mix = {}
for s, e, src, r, idx in data:
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    op = op.split('.')[0]
    m = mix.setdefault(op, [0, 0])
    m[0] += e
    m[1] += s
for op, (e, s) in sorted(mix.items(), key=lambda kv: -kv[1][0])[:22]:
    print('  %-10s exec %5.2f%%  samples %5.2f%%' % (op, 100.0 * e / max(1, tot_e), 100.0 * s / max(1, tot_s)))
