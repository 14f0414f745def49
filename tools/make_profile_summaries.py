"""Regenerate the committed profile summaries from a gpurun capture:

    python tools/make_profile_summaries.py gpurun_out/prof_X.ncu-rep TAG [launches.csv]

writes profiles/TAG_ncu_summary.txt (tracked metrics + stall reasons per captured launch),
profiles/TAG_*_sass_phases.txt (warp-instructions per frame by SASS run), profiles/TAG_traffic.json (per
kernel: DRAM bytes per launch and the pipe utilisations bench.py quotes) and, with a launch list
(`ncu --metrics gpu__time_duration.sum` of bench.py), profiles/TAG_bench_launch_shares.txt.

A report with every kernel and --import-source exceeds gpurun's 64 MiB return limit, so round 2 captured the
Griffin-Lim kernels (`-k regex:gl_step_kernel`, tag r2b) and the feature kernels (`-k regex:stft_feature_kernel`,
tag r2c) separately and merged the two TAG_traffic.json files into profiles/r2_traffic.json."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
launches = sys.argv[3] if len(sys.argv) > 3 else None
here = os.path.dirname(os.path.abspath(__file__))

if launches:
    rows = list(csv.reader(open(launches)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'second': 1e6}.get(r[mu], 1.0)
        agg.setdefault(r[kn], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    out = ['# ncu launch list of `python bench.py --workload gl256 --steps 2 --warmup 1` (first launches)',
           '# gpu__time_duration.sum per launch, --clock-control none; cold-cache / serialised: compare SHARES', '']
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append('%6.2f%%  n=%4d  avg %9.1f us  %s' % (100 * sum(v) / tot, len(v), sum(v) / len(v), k[:120]))
    open('profiles/%s_bench_launch_shares.txt' % tag, 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out))

summary = subprocess.run([sys.executable, os.path.join(here, 'ncu_summary.py'), rep], capture_output=True, text=True).stdout
open('profiles/%s_ncu_summary.txt' % tag, 'w').write(summary)

raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
names = [r[h.index('Kernel Name')] for r in rows[2:]]


def first_index(pattern):
    for i, n in enumerate(names):
        if pattern in n:
            return i
    return None


# per-frame SASS phase tables: (file suffix, kernel substring, frames of that launch, title)
for suffix, pat, frames, title in (
        ('gl', 'gl_step_kernel<float, StaticGeom<1102, 275, 2048>, 8, 0, 0, 0>', 112916, 'gl_step_kernel<float, model geometry> (iteration launch)'),
        ('gl1024', 'gl_step_kernel<float, NativeGeom1024<1024, 256>, 8, 0, 0, 0>', 121282, 'gl_step_kernel<float, native n_fft 1024> (iteration launch)'),
        ('feat_f32', 'stft_feature_kernel<float, StaticGeom<1102, 275, 2048>, 8, 1>', 112916, 'stft_feature_kernel<float, model geometry, fused dB mode>'),
        ('feat_f64', 'stft_feature_kernel<double, StaticGeom<1102, 275, 2048>, 4, 1>', 112916, 'stft_feature_kernel<double, model geometry, fused dB mode>'),
        ('stats_f64', 'stft_feature_kernel<double, NativeGeom1024<1024, 256>, 4, 1>', 121282, 'stft_feature_kernel<double, native n_fft 1024, fused statistics mode>')):
    idx = first_index(pat)
    if idx is None:
        continue
    # launch-skip counts launches matching the regex: position of this launch among same-pattern launches
    short = pat.split('<')[0]
    skip = sum(1 for n in names[:idx] if short in n)
    body = subprocess.run([sys.executable, os.path.join(here, 'ncu_runs.py'), rep, short, str(skip), str(frames)],
                          capture_output=True, text=True).stdout
    open('profiles/%s_%s_sass_phases.txt' % (tag, suffix), 'w').write('# %s: warp-instructions per frame by SASS run\n%s' % (title, body))

BY = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
US = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}
col = {k: h.index(k) for k in (
    'Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
    'sm__warps_active.avg.pct_of_peak_sustained_active') if k in h}
agg = collections.OrderedDict()
for r in rows[2:]:
    name = r[col['Kernel Name']].replace('void ', '').replace('sstts::', '').split('(')[0]
    rd, wr, tm = col['dram__bytes_read.sum'], col['dram__bytes_write.sum'], col['gpu__time_duration.sum']

    def f(key):
        return float(r[col[key]]) if key in col else None

    agg.setdefault(name, []).append({
        'dram_bytes_per_launch': float(r[rd]) * BY[units[rd]] + float(r[wr]) * BY[units[wr]],
        'gpu_time_us_under_ncu': float(r[tm]) * US.get(units[tm], 1),
        'fp32_pipe_active_pct': f('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
        'fp64_pipe_active_pct': f('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
        'issue_slots_active_pct': f('smsp__issue_active.avg.pct_of_peak_sustained_active'),
        'lsu_pipe_active_pct': f('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
        'dram_throughput_pct': f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        'registers_per_thread': f('launch__registers_per_thread'),
        'warps_active_pct': f('sm__warps_active.avg.pct_of_peak_sustained_active')})
traffic = collections.OrderedDict()
for k, v in agg.items():
    traffic[k] = {m: (sum(x[m] for x in v) / len(v) if v[0][m] is not None else None) for m in v[0]}
    traffic[k]['launches'] = len(v)
for k in list(traffic):
    if 'NativeGeom1024<1024, 256>, 4, 1>' in k and k.startswith('stft_feature_kernel<double'):
        traffic['statistics'] = dict(traffic[k], name=k)       # the default (float64) statistics kernel, by role
traffic['_source'] = ('ncu --set full (+ fp64 / lsu / alu pipe metrics) --clock-control none, PROF_ONCE=1 python tools/prof_run.py 3 '
                      '(BASELINE configs[1]/[2] shapes: 256 clips, 112,916 frames at hop 275 / 121,282 at hop 256, whole '
                      'batch per launch); profiles/%s_ncu_summary.txt' % tag)
json.dump(traffic, open('profiles/%s_traffic.json' % tag, 'w'), indent=1)
for k, v in traffic.items():
    print(k, v)
