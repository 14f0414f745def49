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
w = torch.empty(30_000_000, dtype=torch.float32, device=dev)
print('D2H 120 MB ms', wall(lambda: _hostio.download(w)))
for first, growth in ((10 ** 9, 1), (24000, 1), (16000, 1), (10000, 1), (6000, 1), (6000, 2)):
    _runtime._GL_CHUNK_FRAMES, _runtime._GL_CHUNK_GROWTH = first, growth
    n = len(_runtime._split_by_frames(fb.frames, first, growth))
    print('e2e first=%d growth=%d (%d sub-batches): %.2f ms' % (
        first, growth, n, wall(lambda: _runtime.griffin_lim_batch(mags, WIN, HOP, NFFT, 50, seed=3))))
_runtime._GL_CHUNK_FRAMES, _runtime._GL_CHUNK_GROWTH = 10000, 1
for it in (0, 50):
    t = wall(lambda: _runtime.griffin_lim_batch(mags, WIN, HOP, NFFT, it, seed=3))
    print('e2e n_iter=%d: %.2f ms' % (it, t))
