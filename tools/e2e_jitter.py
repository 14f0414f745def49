"""Per-call wall time of spectrograms_to_wavs (pinned inputs), 30 calls:  python tools/e2e_jitter.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import single_speaker_tts_b200 as pkg                             # noqa: E402
from single_speaker_tts_b200 import _runtime                      # noqa: E402
from single_speaker_tts_b200.audio import synthesis               # noqa: E402
from single_speaker_tts_b200.synthetic import make_clips          # noqa: E402

WIN, HOP, NFFT = 1102, 275, 2048
clips = make_clips(256, seed=1, pool=16)
fb = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, want_spec=True, precision='f32', keep_on_device=True)
mag_dev = fb.spec.abs().contiguous()
mag = pkg.pinned_empty(tuple(mag_dev.shape))
torch.from_numpy(mag).copy_(mag_dev)
off = np.concatenate([[0], np.cumsum(fb.frames)])
mags = [mag[off[i]:off[i + 1]].T for i in range(256)]
del fb, mag_dev
ts = []
for i in range(30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    w = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, seed=1234)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
    del w
print(' '.join('%.1f' % t for t in ts))
print('median %.2f  min %.2f  max %.2f' % (np.median(ts), min(ts), max(ts)))

# bench.py style: no device-wide synchronisation between calls, one event after every call
ev = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
ev[0].record()
for i in range(12):
    synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, seed=1234)
    ev[i + 1].record()
torch.cuda.synchronize()
print('events, no sync between calls:', ' '.join('%.1f' % ev[i].elapsed_time(ev[i + 1]) for i in range(12)))
import gc
gc.disable()
ev[0].record()
for i in range(12):
    synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, seed=1234)
    ev[i + 1].record()
torch.cuda.synchronize()
print('same, gc disabled:', ' '.join('%.1f' % ev[i].elapsed_time(ev[i + 1]) for i in range(12)))
gc.enable()
big = [torch.empty(600_000_000, dtype=torch.uint8, device='cuda') for _ in range(4)]   # bench holds ~2.4 GB of device buffers
ev[0].record()
for i in range(12):
    synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, seed=1234)
    ev[i + 1].record()
torch.cuda.synchronize()
print('with 2.4 GB held:', ' '.join('%.1f' % ev[i].elapsed_time(ev[i + 1]) for i in range(12)))
import time
for i in range(6):
    t0 = time.perf_counter()
    w = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, seed=1234)
    t1 = time.perf_counter()
    del w
    t2 = time.perf_counter()
    print('call %.1f ms, dropping the result %.1f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
