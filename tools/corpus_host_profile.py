"""cProfile of one rank's corpus pass (configs[3] shape, 1024 clips):  python tools/corpus_host_profile.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from single_speaker_tts_b200 import distributed                         # noqa: E402
from single_speaker_tts_b200.audio.features import features_batch       # noqa: E402
from single_speaker_tts_b200.synthetic import ClipPlan                  # noqa: E402

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
plan = ClipPlan(n_total, seed=3, kind='ljspeech', pool=32)
mine = list(range(n_total))
wavs = plan.clips(mine)
audio = float(plan.lengths.sum()) / 22050


def one_pass():
    t0 = time.perf_counter()
    mean4, mn4, mx4, _ = distributed.corpus_decibel_statistics(wavs, mine, n_total, 22050, batch_clips=512)
    t1 = time.perf_counter()
    lin_max, lin_ref, mel_max, mel_ref = mean4
    for s in range(0, len(wavs), 256):
        features_batch(wavs[s:s + 256], 2048, 275, 1102, 22050, 80, 0, 8000, lin_ref, lin_max, mel_ref, mel_max, reduction=5)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


one_pass()
torch.cuda.synchronize()
a, b = one_pass()
print('statistics %.1f ms, precalc %.1f ms, %.0f audio-s/s' % (a * 1e3, b * 1e3, audio / (a + b)))
pr = cProfile.Profile()
pr.enable()
one_pass()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
