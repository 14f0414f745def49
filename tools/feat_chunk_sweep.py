"""features_batch end to end vs sub-batch size (pinned and pageable inputs):  python tools/feat_chunk_sweep.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import single_speaker_tts_b200 as pkg                             # noqa: E402
from single_speaker_tts_b200 import _runtime                      # noqa: E402
from single_speaker_tts_b200.audio import features                # noqa: E402
from single_speaker_tts_b200.synthetic import make_clips          # noqa: E402

clips = make_clips(256, seed=1, pool=16)
pin = pkg.pinned_empty((sum(len(c) for c in clips),))
pin[:] = np.concatenate(clips)
off = np.concatenate([[0], np.cumsum([len(c) for c in clips])])
pclips = [pin[off[i]:off[i + 1]] for i in range(256)]
consts = (35.66, 100.0, 6.02, 99.89)


def wall(fn, n=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for chunk, head in ((6 << 20, None), (6 << 20, 1 << 20), (6 << 20, 1 << 19), (6 << 20, 2 << 20), (4 << 20, 1 << 20), (3 << 20, 1 << 20),
                    (8 << 20, 1 << 20), (6 << 20, None)):
    _runtime._FEAT_CHUNK_SAMPLES, _runtime._FEAT_CHUNK_HEAD = chunk, head
    a = wall(lambda: features.features_batch(pclips, 2048, 275, 1102, 22050, 80, 0, 8000, *consts, reduction=5))
    b = wall(lambda: features.features_batch(clips, 2048, 275, 1102, 22050, 80, 0, 8000, *consts, reduction=5))
    print('chunk %10d samples, first %s: pinned %.2f ms, pageable %.2f ms' % (chunk, head, a, b))
