"""Regenerate profiles/ from a gpurun capture:
   python tools/make_profile_summaries.py gpurun_out/prof_X.ncu-rep profiles/r1_bench_launches.csv TAG
writes profiles/r1_bench_launch_shares.txt, profiles/TAG_ncu_summary.txt, profiles/TAG_*_sass_phases.txt,
profiles/r1_traffic.json."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
here = os.path.dirname(os.path.abspath(__file__))
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
out = ['# ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (first 400 launches), final round-1 kernels',
       '# gpu__time_duration.sum per launch, --clock-control none; cold-cache / serialised: compare SHARES', '']
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.append('%6.2f%%  n=%4d  avg %9.1f us  %s' % (100 * sum(v) / tot, len(v), sum(v) / len(v), k[:110]))
open('profiles/r1_bench_launch_shares.txt', 'w').write('\n'.join(out) + '\n')

summary = subprocess.run([sys.executable, os.path.join(here, 'ncu_summary.py'), rep], capture_output=True, text=True).stdout
open('profiles/%s_ncu_summary.txt' % tag, 'w').write(summary)
for name, pat, skip, title in (('gl', 'gl_step_kernel', '1', 'gl_step_kernel<float, model geometry> (iteration launch)'),
                               ('feat_f32', 'stft_feature_kernel', '1', 'stft_feature_kernel<float, ..., kDbFeatures>'),
                               ('feat_f64', 'stft_feature_kernel', '3', 'stft_feature_kernel<double, ..., kDbFeatures>')):
    body = subprocess.run([sys.executable, os.path.join(here, 'ncu_runs.py'), rep, pat, skip], capture_output=True, text=True).stdout
    open('profiles/%s_%s_sass_phases.txt' % (tag, name), 'w').write('# %s: warp-instructions per frame by SASS run\n%s' % (title, body))

raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
kn, r_, w_, t_ = (h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum'),
                  h.index('gpu__time_duration.sum'))
f_ = h.index('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active')
i_ = h.index('smsp__issue_active.avg.pct_of_peak_sustained_active')
d_ = h.index('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')
BY = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
US = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}
agg = collections.OrderedDict()
for r in rows[2:]:
    name = r[kn].replace('void ', '').replace('sstts::', '').split('(')[0]
    agg.setdefault(name, []).append((float(r[r_]) * BY[units[r_]] + float(r[w_]) * BY[units[w_]], float(r[t_]) * US.get(units[t_], 1),
                                     float(r[f_]), float(r[i_]), float(r[d_])))
traffic = {k: {'dram_bytes_per_launch': sum(x[0] for x in v) / len(v), 'launches': len(v),
               'gpu_time_us_under_ncu': sum(x[1] for x in v) / len(v),
               'fp32_pipe_active_pct': sum(x[2] for x in v) / len(v), 'issue_slots_active_pct': sum(x[3] for x in v) / len(v),
               'dram_throughput_pct': sum(x[4] for x in v) / len(v)} for k, v in agg.items()}
traffic['_source'] = ('ncu --set full --clock-control none -k regex:gl_step_kernel|stft_feature_kernel|gl_finalize|random_phase -c 20, '
                      'python tools/prof_run.py 3 (BASELINE configs[1]/[2] shapes: 256 clips, 112,916 frames, whole batch per '
                      'launch); profiles/%s_ncu_summary.txt' % tag)
json.dump(traffic, open('profiles/r1_traffic.json', 'w'), indent=1)
print('\n'.join(out))
for k, v in traffic.items():
    print(k, v)
