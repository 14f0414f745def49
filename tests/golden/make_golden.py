"""Generate the committed golden vectors under tests/golden/ with the CPU oracle.

Run in the development container (needs /root/reference for the real model-output fixture):

    python tests/golden/make_golden.py

The reference pins no numerical result on this path and cannot be executed here (librosa and
TensorFlow 1.8 are not installable), so these vectors are produced by ``oracle/`` -- the numpy
restatement of the reference -- and record its output for seeded inputs.  ``tests/`` compares
both the oracle (regression) and the CUDA path (parity) against them; the GPU box has no
/root/reference, so everything the tests need is stored in the .npz files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import librosa_compat as lc            # noqa: E402
from oracle import reference_audio as ra           # noqa: E402
from single_speaker_tts_b200.synthetic import make_clips, speech_like_clip  # noqa: E402

FIXTURE = ('/root/reference/visualization/data/ljspeech/v1.1/post-processing/'
           'ljspeech-linear-spec-post-215k.npz')
WIN, HOP, NFFT = 1102, 275, 2048


def gl_fixture():
    """Crop of the reference's dumped model output (tacotron/model.py:573-596 writes
    ``linear_spec`` (1, 1025, 1000, 1) in [0, 1]) pushed through the inference recipe
    (tacotron/inference.py:94-101,175) and 50 Griffin-Lim iterations."""
    spec = np.load(FIXTURE)['linear_spec'][0, :, :, 0].T        # (1000, 1025), what session.run returns
    crop = np.ascontiguousarray(spec[180:340]).astype(np.float32)  # 160 frames of speech
    mag = ra.inference_postprocess(crop)                          # (1025, 160) float32
    angles = np.exp(2j * np.pi * np.random.RandomState(20260).rand(*mag.shape))
    wav, mse = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, 50, angles=angles)
    np.savez_compressed(os.path.join(HERE, 'gl_fixture.npz'), model_output=crop, seed=20260,
                        n_iter=50, wav=wav.astype(np.float32), mse=np.float64(mse))
    print('gl_fixture', crop.shape, wav.shape, mse)


def gl_synthetic():
    clips = [c[:n] for c, n in zip(make_clips(3, seed=5), (9000, 3000, 14000))]
    out = {}
    for i, c in enumerate(clips):
        mag = np.abs(lc.stft(c, NFFT, HOP, WIN))
        angles = np.exp(2j * np.pi * np.random.RandomState(100 + i).rand(*mag.shape))
        wav, mse = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, 50, angles=angles)
        out['clip%d' % i] = c
        out['wav%d' % i] = wav.astype(np.float32)
        out['mse%d' % i] = np.float64(mse)
    np.savez_compressed(os.path.join(HERE, 'gl_synthetic.npz'), n_iter=50, seed0=100, **out)
    print('gl_synthetic', [out['wav%d' % i].shape for i in range(3)])


def features():
    clips = [c[:n] for c, n in zip(make_clips(3, seed=9), (22050, 7000, 300))]
    out = {}
    for i, c in enumerate(clips):
        mel, lin = ra.load_audio_from_wav(c, 22050, trim=False)
        out['clip%d' % i] = c
        out['mel%d' % i] = mel
        out['lin%d' % i] = lin
        out['stats%d' % i] = ra.decibel_statistics(c, 22050)
    out['corpus_stats'] = ra.collect_decibel_statistics_from_wavs(clips, 22050)
    np.savez_compressed(os.path.join(HERE, 'features.npz'), **out)
    print('features', [out['lin%d' % i].shape for i in range(3)], out['corpus_stats'])


def gl_config0():
    """BASELINE configs[0] exactly: one synthetic 5 s clip (110,250 samples -> T = 401 frames),
    50 iterations, initial phase from ``RandomState(0)``."""
    clip = speech_like_clip(5 * 22050, np.random.default_rng(0))
    mag = np.abs(lc.stft(clip, NFFT, HOP, WIN))
    assert mag.shape == (1025, 401)
    angles = np.exp(2j * np.pi * np.random.RandomState(0).rand(*mag.shape))
    wav, mse = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, 50, angles=angles)
    np.savez_compressed(os.path.join(HERE, 'gl_config0.npz'), clip=clip, seed=0, n_iter=50,
                        wav=wav.astype(np.float32), mse=np.float64(mse))
    print('gl_config0', mag.shape, wav.shape, mse)


def gl_fixture_full():
    """The reference's dumped model output at FULL size -- T = 1000 frames, what
    tacotron/inference.py:75-101 really feeds (decoder.maximum_iterations, tacotron/params/model.py:108) --
    through the inference recipe and 50 iterations."""
    spec = np.load(FIXTURE)['linear_spec'][0, :, :, 0].T        # (1000, 1025)
    out = np.ascontiguousarray(spec).astype(np.float32)
    mag = ra.inference_postprocess(out)                           # (1025, 1000)
    angles = np.exp(2j * np.pi * np.random.RandomState(215).rand(*mag.shape))
    wav, mse = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, 50, angles=angles)
    assert wav.shape == (274725,)
    np.savez_compressed(os.path.join(HERE, 'gl_fixture_full.npz'), model_output=out, seed=215,
                        n_iter=50, wav=wav.astype(np.float32), mse=np.float64(mse))
    print('gl_fixture_full', out.shape, wav.shape, mse)


def gl_100_iterations():
    """BASELINE configs[4] runs 100 iterations: two ragged utterances, reported separately because the
    FP32 drift grows with the iteration count."""
    clips = [c[:n] for c, n in zip(make_clips(2, seed=44), (2 * 22050, 77000))]
    out = {}
    for i, c in enumerate(clips):
        mag = np.abs(lc.stft(c, NFFT, HOP, WIN))
        angles = np.exp(2j * np.pi * np.random.RandomState(400 + i).rand(*mag.shape))
        wav, mse = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, 100, angles=angles)
        out['clip%d' % i] = c
        out['wav%d' % i] = wav.astype(np.float32)
        out['mse%d' % i] = np.float64(mse)
    np.savez_compressed(os.path.join(HERE, 'gl_100it.npz'), n_iter=100, seed0=400, **out)
    print('gl_100it', [out['wav%d' % i].shape for i in range(2)])


if __name__ == '__main__':
    only = sys.argv[1:]
    for fn in (gl_fixture, gl_synthetic, features, gl_config0, gl_fixture_full, gl_100_iterations):
        if not only or fn.__name__ in only:
            fn()
