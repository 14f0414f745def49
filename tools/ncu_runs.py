"""Collapse one kernel's SASS (ncu source page) into runs of equal execution count:
   python tools/ncu_runs.py report.ncu-rep <kernel regex> [launch-skip] [frames]
prints [first, last] instruction index, warp-instructions executed per frame, samples share and the opcode mix."""
import csv, io, subprocess, sys
from collections import Counter
rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else '0'
frames = float(sys.argv[4]) if len(sys.argv) > 4 else 112916.0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat,
                      '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO('\n'.join(raw.splitlines()[1:]))))
h = rows[0]
isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
data = []
for r in rows[1:]:
    try:
        data.append((r[isrc], int(r[isamp] or 0), int(r[iex] or 0)))
    except (ValueError, IndexError):
        pass
if len(data) > 2 and data[:len(data) // 2] == data[len(data) // 2:]:
    data = data[:len(data) // 2]
tot_s = sum(d[1] for d in data); tot_e = sum(d[2] for d in data)
print('static %d executed %d (%.0f per frame) samples %d' % (len(data), tot_e, tot_e / frames, tot_s))
def op(src):
    t = src.split()
    o = t[1] if t[0].startswith('@') else t[0]
    return o.split('.')[0]
i = 0
while i < len(data):
    j = i
    while j + 1 < len(data) and abs(data[j + 1][2] - data[i][2]) <= 0.02 * max(1, data[i][2]):
        j += 1
    e = sum(d[2] for d in data[i:j + 1]); s = sum(d[1] for d in data[i:j + 1])
    if e / frames >= 3 or s > 0.003 * tot_s:
        mix = Counter(op(d[0]) for d in data[i:j + 1])
        print('[%5d,%5d] n=%4d x%8.3f/frame = %7.1f instr/frame (%4.1f%%) samples %4.1f%%  %s' % (
            i, j, j - i + 1, data[i][2] / frames, e / frames, 100.0 * e / tot_e, 100.0 * s / tot_s,
            ' '.join('%s:%d' % kv for kv in mix.most_common(7))))
    i = j + 1
