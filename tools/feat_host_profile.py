"""cProfile of the host side of features_batch (256 clips):  python tools/feat_host_profile.py"""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from single_speaker_tts_b200 import _runtime                      # noqa: E402
from single_speaker_tts_b200.audio import features, synthesis      # noqa: E402
from single_speaker_tts_b200.synthetic import make_clips          # noqa: E402

clips = make_clips(256, seed=1, pool=16)
consts = (35.66, 100.0, 6.02, 99.89)
_runtime._FEAT_CHUNK_SAMPLES = int(sys.argv[1]) if len(sys.argv) > 1 else 3 << 20


def run():
    for _ in range(10):
        features.features_batch(clips, 2048, 275, 1102, 22050, 80, 0, 8000, *consts, reduction=5)


run()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
