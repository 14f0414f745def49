"""Small workload touching every kernel (both staging variants via SSTTS_GL_STAGING, the native n_fft 1024 path in
its generic and fused modes, the fused dB-feature mode, trim, glue kernels) for compute-sanitizer:

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from single_speaker_tts_b200 import _runtime                                  # noqa: E402
from single_speaker_tts_b200.audio import effects, features, synthesis         # noqa: E402
from single_speaker_tts_b200.datasets import statistics                        # noqa: E402
from single_speaker_tts_b200.datasets.dataset_helper import LJSpeechDatasetHelper  # noqa: E402
from single_speaker_tts_b200.synthetic import speech_like_clip                 # noqa: E402

rng = np.random.default_rng(3)
clips = [speech_like_clip(int(n), rng) for n in (300, 5000, 22050, 275 * 40 + 3, 256 * 33 + 1)]
res = _runtime.stft_features_batch(clips, 2048, 275, 1102, want_spec=True)
mags = [np.abs(res.rows(res.spec, i)).T for i in range(len(clips))]
w = synthesis.spectrograms_to_wavs(mags, 1102, 275, 2048, 3, seed=5, return_mse=True)
w1 = synthesis.spectrogram_to_wav(mags[2], 1102, 275, 2048, 2)
w2 = synthesis.model_outputs_to_wavs([np.clip(m.T / (m.max() + 1e-9), 0, 1) for m in mags[:2]], 6.02, 99.89, 1.3, 1102, 275, 2048, 2,
                                     seed=1, normalize_peak=True)
f = LJSpeechDatasetHelper.features_from_wavs(clips, 22050)
f32 = LJSpeechDatasetHelper.features_from_wavs(clips, 22050, precision='f32')
st = statistics.decibel_statistics_batch(clips, 22050)
st32 = statistics.decibel_statistics_batch(clips, 22050, precision='f32')
s1024 = features.linear_scale_spectrogram(clips[2], 1024)
m1024 = features.mel_scale_spectrogram(clips[2], 1024, 22050, 80, 0, 8000, 200, 800, 1)
m2048 = features.mel_scale_spectrogram(clips[1], 2048, 22050, 80, 0, 8000, 275, 1102, 2)
ts = effects.time_stretch(clips[1], 1.25)
tr = effects.trim_batch(clips)
mf = features.calculate_mfccs(np.log(np.maximum(1e-5, m2048)), 22050, 13)
print('sanitize workload ok', len(w[0]), f[0][0].shape, st.shape, s1024.shape, ts.shape)
