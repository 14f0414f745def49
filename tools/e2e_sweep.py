"""Sweep of the sub-batch schedule of spectrograms_to_wavs (first sub-batch size in frames, growth factor) on the
BASELINE configs[2] shape with pageable inputs: wall-clock per call.   python tools/e2e_sweep.py"""
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
clips = make_clips(256, seed=1, pool=16)
fb = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, want_spec=True, precision='f32', keep_on_device=True)
mag = fb.spec.abs().contiguous().cpu().numpy()
off = np.concatenate([[0], np.cumsum(fb.frames)])
mags = [mag[off[i]:off[i + 1]].T for i in range(256)]
audio_s = sum(HOP * (t - 1) for t in fb.frames) / 22050
for first, growth, head, cap in ((10000, 1, None, None), (10000, 1, 2500, None), (10000, 1, 4000, None), (10000, 1, 1500, None), (2500, 2, None, 10000), (8000, 1, 2500, None), (12000, 1, 3000, None), (10000, 1, None, None)):
    _runtime._GL_CHUNK_FRAMES, _runtime._GL_CHUNK_GROWTH, _runtime._GL_CHUNK_HEAD, _runtime._GL_CHUNK_CAP = first, growth, head, cap
    n_sub = len(_runtime._split_by_frames(list(fb.frames), first, growth, head=head, cap=cap))
    for _ in range(8):
        synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, seed=1)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, seed=1)
        ts.append((time.perf_counter() - t0) * 1e3)
    print('first %6d growth %.1f head %s cap %s -> %2d sub-batches: median %.2f ms  min %.2f ms  (%.0f audio-s/s)' % (
        first, growth, head, cap, n_sub, sorted(ts)[len(ts) // 2], min(ts), audio_s / (sorted(ts)[len(ts) // 2] / 1e3)))
