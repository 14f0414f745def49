"""Host-side timeline of one pipelined spectrograms_to_wavs call (BASELINE configs[2] shape, pageable inputs):
when each sub-batch was packed + uploaded (helper thread), how long the launching thread waited for it and when
its launches were enqueued.   python tools/e2e_trace.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from single_speaker_tts_b200 import _runtime                      # noqa: E402
from single_speaker_tts_b200.audio import synthesis               # noqa: E402
from single_speaker_tts_b200.synthetic import make_clips          # noqa: E402

WIN, HOP, NFFT = 1102, 275, 2048
if os.environ.get('E2E_HEAD'):      # size of the first sub-batch in frames (default: like the others)
    _runtime._GL_CHUNK_HEAD = int(os.environ['E2E_HEAD'])
clips = make_clips(256, seed=1, pool=16)
fb = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, want_spec=True, precision='f32', keep_on_device=True)
mag = fb.spec.abs().contiguous().cpu().numpy()
off = np.concatenate([[0], np.cumsum(fb.frames)])
mags = [mag[off[i]:off[i + 1]].T for i in range(256)]
for _ in range(8):
    synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, seed=1)
import single_speaker_tts_b200 as pkg
mag_pin = pkg.pinned_empty(mag.shape); mag_pin[:] = mag
mags_pin = [mag_pin[off[i]:off[i + 1]].T for i in range(256)]
for rep, mm in enumerate((mags, mags_pin)):
    for _ in range(2):
        synthesis.spectrograms_to_wavs(mm, WIN, HOP, NFFT, 50, seed=1)
    _runtime._trace = []
    t0 = time.perf_counter()
    synthesis.spectrograms_to_wavs(mm, WIN, HOP, NFFT, 50, seed=1)
    t1 = time.perf_counter()
    tr, _runtime._trace = _runtime._trace, None
    print('call %.2f ms (%s inputs)' % ((t1 - t0) * 1e3, 'pinned' if rep == 1 else 'pageable'))
    gpu = [r for r in tr if r[0] == 'gpu']
    tr = [r for r in tr if r[0] != 'gpu']
    base = gpu[0][2]
    for _, k, g0, g1 in gpu:
        print('  gpu      %2d  start %6.2f  end %6.2f  (%.2f ms)   [relative to the first launch]' % (
            k, base.elapsed_time(g0), base.elapsed_time(g1), g0.elapsed_time(g1)))
    for what, k, a, b in sorted(tr, key=lambda r: r[2]):
        print('  %-8s %2d  start %6.2f  end %6.2f  (%.2f ms)' % (what, k, (a - t0) * 1e3, (b - t0) * 1e3, (b - a) * 1e3))
