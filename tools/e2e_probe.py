"""Where the end-to-end time of spectrograms_to_wavs goes (BASELINE configs[2] shape):
   python tools/e2e_probe.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from single_speaker_tts_b200 import _hostio, _runtime            # noqa: E402
from single_speaker_tts_b200.synthetic import make_clips         # noqa: E402

WIN, HOP, NFFT = 1102, 275, 2048
clips = make_clips(256, seed=1, pool=16)
fb = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, want_spec=True, precision='f32', keep_on_device=True)
mag = fb.spec.abs().contiguous().cpu().numpy()
off = np.concatenate([[0], np.cumsum(fb.frames)])
mags = [mag[off[i]:off[i + 1]].T for i in range(256)]
dev = torch.device('cuda', 0)


def wall(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print('upload_rows all (pack + H2D) ms', wall(lambda: _hostio.upload_rows([m.T for m in mags], 1025, torch.float32, dev, slot='probe')))
big = torch.empty((int(off[-1]), 1025), dtype=torch.float32, pin_memory=True)
big.numpy()[:] = mag
d = torch.empty_like(big, device=dev)
print('H2D only from pinned ms', wall(lambda: d.copy_(big, non_blocking=True)))
stage = np.empty_like(mag)
def pack():
    o = 0
    for m in mags:
        stage[o:o + m.shape[1]] = m.T
        o += m.shape[1]
print('pack single thread (numpy copies) ms', wall(pack))
w = torch.empty(125_000_000, dtype=torch.float32, device=dev)
wh = torch.empty(125_000_000, dtype=torch.float32, pin_memory=True)
t = wall(lambda: wh.copy_(w, non_blocking=True))
print('D2H 500 MB into a fixed pinned buffer: %.2f ms = %.1f GB/s' % (t, 0.5 / t * 1e3))
t = wall(lambda: w.copy_(wh, non_blocking=True))
print('H2D 500 MB from a fixed pinned buffer: %.2f ms = %.1f GB/s' % (t, 0.5 / t * 1e3))
s2 = torch.cuda.Stream()
def duplex():
    wh.copy_(w, non_blocking=True)
    with torch.cuda.stream(s2):
        d.copy_(big, non_blocking=True)
t = wall(duplex)
print('D2H 500 MB + H2D 463 MB concurrently: %.2f ms' % t)
def dl():
    r = _hostio.download(w)
    torch.cuda.current_stream().synchronize()
    return r
print('_hostio.download 500 MB (+ sync, result dropped): %.2f ms' % wall(dl))
for first, growth in ((10 ** 9, 1), (24000, 1), (16000, 1), (10000, 1), (6000, 1), (6000, 2)):
    _runtime._GL_CHUNK_FRAMES, _runtime._GL_CHUNK_GROWTH = first, growth
    n = len(_runtime._split_by_frames(fb.frames, first, growth))
    print('e2e first=%d growth=%d (%d sub-batches): %.2f ms' % (
        first, growth, n, wall(lambda: _runtime.griffin_lim_batch(mags, WIN, HOP, NFFT, 50, seed=3))))
_runtime._GL_CHUNK_FRAMES, _runtime._GL_CHUNK_GROWTH = 10000, 1
for it in (0, 50):
    t = wall(lambda: _runtime.griffin_lim_batch(mags, WIN, HOP, NFFT, it, seed=3))
    print('e2e n_iter=%d: %.2f ms' % (it, t))

from single_speaker_tts_b200.audio import synthesis, features        # noqa: E402
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, seed=1234)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, seed=1234)
    e1.record()
    torch.cuda.synchronize()
    print('bench-style e2e (events, 5 calls): %.2f ms / call' % (e0.elapsed_time(e1) / 5))
consts = (35.66, 100.0, 6.02, 99.89)
for chunk in (1 << 40, 6 << 20, 3 << 20, 3 << 19):
    _runtime._FEAT_CHUNK_SAMPLES = chunk
    t = wall(lambda: features.features_batch(clips, NFFT, HOP, WIN, 22050, 80, 0, 8000, *consts, reduction=5))
    print('features e2e chunk=%d samples: %.2f ms' % (chunk, t))
