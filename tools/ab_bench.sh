#!/bin/bash
# A/B several builds of libsstts on the GPU box: tools/ab_bench.sh lib1.so lib2.so ...
for lib in "$@"; do
  echo "== $lib"
  SSTTS_LIB=$PWD/$lib python bench.py --workload gl256 --steps 5 --warmup 3 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print('GL %.0f audio-s/s  %.3f ms/launch  frac %.3f | e2e %.0f' % (d['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value']))
        f = d['features']
        print('feat f64 %.0f (%.2f ms)  f32 %.0f (%.2f ms)' % (f['f64']['value'], f['f64']['ms_per_step'], f['f32_fast']['value'], f['f32_fast']['ms_per_step']))
"
done
