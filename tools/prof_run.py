"""Small fixed workload for ncu: BASELINE configs[1]/[2] shapes (256 ragged clips), Griffin-Lim with
a few iterations and one feature pass per precision, launched through the C ABI.

    python tools/prof_run.py [n_iter] [n_utts]
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from single_speaker_tts_b200 import _lib, _runtime            # noqa: E402
from single_speaker_tts_b200.synthetic import make_clips      # noqa: E402

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_utts = int(sys.argv[2]) if len(sys.argv) > 2 else 256
REPS = 1 if os.environ.get('PROF_ONCE') else 2      # under ncu every launch is replayed anyway
WIN, HOP, NFFT = 1102, 275, 2048
clips = make_clips(n_utts, seed=1, pool=16)
mags = []
fb = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, want_spec=True, precision='f32', keep_on_device=True)
mag = fb.spec.abs().contiguous().cpu().numpy()
off = np.concatenate([[0], np.cumsum(fb.frames)])
mags = [mag[off[i]:off[i + 1]].T for i in range(n_utts)]
_runtime._GL_CHUNK_FRAMES = 10 ** 9      # whole batch in one launch sequence, like bench.py's device-resident `value`
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(REPS):
    e0.record()
    wavs = _runtime.griffin_lim_batch(mags, WIN, HOP, NFFT, n_iter, seed=3)[0]
    e1.record(); torch.cuda.synchronize()
    print('gl e2e n_iter', n_iter, 'ms', e0.elapsed_time(e1))
for prec in ('f32', 'f64'):
    for rep in range(REPS):
        e0.record()
        r = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, sampling_rate=22050, n_mels=80, fmin=0, fmax=8000,
                                         reduction=5, want_lin=True, want_mel=True, normalize=(35.66, 100.0, 6.02, 99.89),
                                         precision=prec, keep_on_device=True)
        e1.record(); torch.cuda.synchronize()
        print('features', prec, 'ms', e0.elapsed_time(e1))
for prec in ('f64', 'f32'):
    for rep in range(REPS):
        e0.record()
        r = _runtime.stft_features_batch(clips, 1024, 256, 1024, sampling_rate=22050, n_mels=80, fmin=0, fmax=11025,
                                         want_minmax=True, precision=prec, keep_on_device=True)
        e1.record(); torch.cuda.synchronize()
        print('statistics', prec, 'ms', e0.elapsed_time(e1))
# Griffin-Lim at n_fft 1024 / hop 256 (native half-warp transform)
fb1 = _runtime.stft_features_batch(clips, 1024, 256, 1024, want_spec=True, precision='f32', keep_on_device=True)
mag1 = fb1.spec.abs().contiguous().cpu().numpy()
off1 = np.concatenate([[0], np.cumsum(fb1.frames)])
mags1 = [mag1[off1[i]:off1[i + 1]].T for i in range(n_utts)]
for rep in range(REPS):
    e0.record()
    _runtime.griffin_lim_batch(mags1, 1024, 256, 1024, n_iter, seed=3)
    e1.record(); torch.cuda.synchronize()
    print('gl n_fft 1024 n_iter', n_iter, 'ms', e0.elapsed_time(e1))
print('ok')
