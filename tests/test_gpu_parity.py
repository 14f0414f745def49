"""Parity of the CUDA path (through the public API -> C ABI -> sm_100a kernels) with the oracle.

Tolerances are the ones BASELINE.json:north_star states:
  * Griffin-Lim waveform: relative L2 <= 1e-3 after 50 iterations in FP32;
  * normalised spectrograms: max abs error <= 1e-4;
  * statistics reduction (host float64, listing order): bit-exact given the per-clip rows.
Run with ``pytest -m gpu`` on the B200 box; nothing here reads /root/reference.
"""
import numpy as np
import pytest
import torch

from oracle import librosa_compat as lc
from oracle import reference_audio as ra
from single_speaker_tts_b200 import _lib, _runtime
from single_speaker_tts_b200.audio import features, synthesis
from single_speaker_tts_b200.datasets import statistics
from single_speaker_tts_b200.datasets.dataset_helper import (BlizzardNancyDatasetHelper, CMUDatasetHelper,
                                                            LJSpeechDatasetHelper, PAVOQUEDatasetHelper)
from single_speaker_tts_b200.synthetic import make_clips, speech_like_clip

pytestmark = pytest.mark.gpu

WIN, HOP, NFFT = 1102, 275, 2048
GL_TOL = 1e-3      # north_star: waveform relative L2 after 50 iterations, FP32
NORM_TOL = 1e-4    # north_star: normalised spectrogram max abs error


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(1e-30, np.linalg.norm(b)))


def _case(frames, seed=5):
    rng = np.random.default_rng(seed)
    mags, angs = [], []
    for i, T in enumerate(frames):
        x = speech_like_clip(HOP * (T - 1) + 5, rng)
        m = np.abs(lc.stft(x, NFFT, HOP, WIN))
        mags.append(m)
        angs.append(np.exp(2j * np.pi * np.random.RandomState(i).rand(*m.shape)))
    return mags, angs


def test_native_library_is_loaded():
    lib = _lib.load()
    assert lib.sstts_device_count() >= 1
    assert torch.cuda.get_device_capability(0)[0] >= 10, 'kernels are built for sm_100a only'


@pytest.mark.parametrize('precision,tol', [('f32', 5e-6), ('f64', 5e-7)])
def test_griffin_lim_small_ragged(precision, tol):
    frames = [1, 2, 5, 9, 13, 30, 77]
    mags, angs = _case(frames)
    wavs, mses = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 2, angles=angs, precision=precision,
                                                return_mse=True)
    for T, m, a, w, mse in zip(frames, mags, angs, wavs, mses):
        assert w.dtype == np.float32 and w.shape == (HOP * (T - 1),)
        if T == 1:
            continue
        ref, rmse = ra.griffin_lim_v2(m, WIN, HOP, NFFT, 2, angles=a, batched_fft=True)
        assert rel_l2(w, ref) < tol
        assert abs(mse - rmse) / rmse < 1e-4


def test_griffin_lim_50_iterations_golden_synthetic(golden_dir):
    g = np.load(golden_dir + '/gl_synthetic.npz')
    clips = [g['clip%d' % i] for i in range(3)]
    mags = [np.abs(lc.stft(c, NFFT, HOP, WIN)) for c in clips]
    angs = [np.exp(2j * np.pi * np.random.RandomState(int(g['seed0']) + i).rand(*m.shape)) for i, m in enumerate(mags)]
    wavs, mses = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, int(g['n_iter']), angles=angs, return_mse=True)
    for i, w in enumerate(wavs):
        assert rel_l2(w, g['wav%d' % i]) <= GL_TOL
        assert abs(mses[i] - float(g['mse%d' % i])) / float(g['mse%d' % i]) < 1e-2


def test_griffin_lim_50_iterations_real_model_output(golden_dir):
    """The reference's dumped model output through tacotron/inference.py:94-101,175 + 50 iterations."""
    g = np.load(golden_dir + '/gl_fixture.npz')
    mag = ra.inference_postprocess(g['model_output'])
    ang = np.exp(2j * np.pi * np.random.RandomState(int(g['seed'])).rand(*mag.shape))
    wav, mse = synthesis.griffin_lim_v2(mag, WIN, HOP, NFFT, int(g['n_iter']), angles=ang)
    assert wav.shape == g['wav'].shape
    assert rel_l2(wav, g['wav']) <= GL_TOL
    assert abs(mse - float(g['mse'])) / float(g['mse']) < 1e-2


def test_griffin_lim_baseline_config0_401_frames(golden_dir):
    """BASELINE configs[0] exactly: one 5 s clip, T = 401, 50 iterations, RandomState(0) phase."""
    g = np.load(golden_dir + '/gl_config0.npz')
    mag = np.abs(lc.stft(g['clip'], NFFT, HOP, WIN))
    assert mag.shape == (1025, 401)
    ang = np.exp(2j * np.pi * np.random.RandomState(int(g['seed'])).rand(*mag.shape))
    wav, mse = synthesis.griffin_lim_v2(mag, WIN, HOP, NFFT, int(g['n_iter']), angles=ang)
    err = rel_l2(wav, g['wav'])
    print('config0 T=401 50 it: rel-L2 %.3e, mse rel err %.3e' % (err, abs(mse - float(g['mse'])) / float(g['mse'])))
    assert wav.shape == (110000,) and err <= GL_TOL
    assert abs(mse - float(g['mse'])) / float(g['mse']) < 1e-2


def test_griffin_lim_full_1000_frame_model_output(golden_dir):
    """What tacotron/inference.py:75-101 really feeds: the full (1000, 1025) model output
    (decoder.maximum_iterations = 1000, tacotron/params/model.py:108), 50 iterations, both through the
    per-item signature with the oracle's magnitudes and through the fused model-output glue."""
    g = np.load(golden_dir + '/gl_fixture_full.npz')
    mag = ra.inference_postprocess(g['model_output'])
    ang = np.exp(2j * np.pi * np.random.RandomState(int(g['seed'])).rand(*mag.shape))
    wav, mse = synthesis.griffin_lim_v2(mag, WIN, HOP, NFFT, int(g['n_iter']), angles=ang)
    err = rel_l2(wav, g['wav'])
    print('fixture T=1000 50 it: rel-L2 %.3e' % err)
    assert wav.shape == (274725,) and err <= GL_TOL
    assert abs(mse - float(g['mse'])) / float(g['mse']) < 1e-2
    # fused de-normalisation (device float32 exp2) instead of the host recipe: same waveform within budget
    wavs, _ = _runtime.griffin_lim_batch([g['model_output']], WIN, HOP, NFFT, int(g['n_iter']),
                                         angles=[ang], denormalize=(6.02, 99.89, 1.3))
    assert rel_l2(wavs[0], g['wav']) <= GL_TOL


def test_griffin_lim_100_iterations_reported_separately(golden_dir):
    """BASELINE configs[4] runs 100 iterations.  north_star's 1e-3 is stated for 50; FP32 drift grows with
    the iteration count, so the 100-iteration error is measured and reported on its
    own, against the looser bound below; the float64 instantiation separates drift from algorithm."""
    g = np.load(golden_dir + '/gl_100it.npz')
    clips = [g['clip%d' % i] for i in range(2)]
    mags = [np.abs(lc.stft(c, NFFT, HOP, WIN)) for c in clips]
    angs = [np.exp(2j * np.pi * np.random.RandomState(int(g['seed0']) + i).rand(*m.shape)) for i, m in enumerate(mags)]
    n_iter = int(g['n_iter'])
    assert n_iter == 100
    w32 = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, n_iter, angles=angs)
    w64 = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, n_iter, angles=angs, precision='f64')
    for i in range(2):
        e32, e64 = rel_l2(w32[i], g['wav%d' % i]), rel_l2(w64[i], g['wav%d' % i])
        print('100 it, utterance %d (T=%d): rel-L2 f32 %.3e, f64 %.3e' % (i, mags[i].shape[1], e32, e64))
        assert e32 <= 5e-3      # reported bound for 100 iterations in FP32 (the stated 1e-3 is for 50)
        assert e64 <= GL_TOL    # the same kernels in float64 stay inside the 50-iteration budget


def test_utterances_sampled_from_the_full_256_batch_match_the_oracle():
    """BASELINE configs[2] at full size: ONE batched call over the 256 ragged utterances (112,916 frames,
    sub-batch pipeline, tile tables at full size), 50 iterations, device-seeded phase; four utterances
    (shortest, longest, two in between) are then checked against the oracle run with the very same
    initial phasors (read back with sstts_random_phase_at)."""
    import ctypes
    clips = make_clips(256, seed=1, pool=16)
    frames = np.array([1 + len(c) // HOP for c in clips])
    order = np.argsort(frames)
    picks = [int(order[0]), int(order[85]), int(order[170]), int(order[-1])]
    fb = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, want_spec=True, precision='f64')
    mags = [np.abs(fb.rows(fb.spec, i)).T for i in range(256)]
    for i in picks:                                   # the oracle's own |STFT| for the checked ones
        mags[i] = np.abs(lc.stft(clips[i], NFFT, HOP, WIN))
    seed = 20261018
    wavs = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, seed=seed)
    lib = _lib.load()
    foff = np.concatenate([[0], np.cumsum(frames)])
    for i in picks:
        T = int(frames[i])
        ph = torch.empty((T * 1025, 2), dtype=torch.float32, device='cuda')
        _lib.check(lib.sstts_random_phase_at(ctypes.c_uint64(seed), int(foff[i]) * 1025, T * 1025,
                                             ctypes.c_void_p(ph.data_ptr()),
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        ang = ph.cpu().numpy().view(np.complex64).reshape(T, 1025).T
        ref = ra.spectrogram_to_wav(mags[i], WIN, HOP, NFFT, 50, angles=ang, batched_fft=True)
        err = rel_l2(wavs[i], ref)
        print('utterance %d of the 256-batch (T=%d): rel-L2 %.3e' % (i, T, err))
        assert wavs[i].shape == ref.shape == (HOP * (T - 1),) and err <= GL_TOL


def test_single_frame_spectrogram_follows_the_reference():
    """T = 1: the reference's per-item functions raise ValueError from the reflect padding of the empty
    re-analysis signal when n_iter >= 1 (audio/synthesis.py:96-106) and return (empty, None) for
    n_iter = 0; the batched extension skips such items (empty waveform, mse None) instead of failing
    the whole batch."""
    m = np.ones((1025, 1), np.float32)
    with pytest.raises(ValueError):
        synthesis.griffin_lim_v2(m, WIN, HOP, NFFT, 2)
    with pytest.raises(ValueError):
        synthesis.spectrogram_to_wav(m, WIN, HOP, NFFT, 2)
    w, mse = synthesis.griffin_lim_v2(m, WIN, HOP, NFFT, 0)
    assert w.shape == (0,) and w.dtype == np.float32 and mse is None
    wavs, mses = synthesis.spectrograms_to_wavs([m, np.ones((1025, 3), np.float32)], WIN, HOP, NFFT, 2, seed=1,
                                                return_mse=True)
    assert wavs[0].shape == (0,) and mses[0] is None and wavs[1].shape == (2 * HOP,) and mses[1] > 0


def test_griffin_lim_dropin_uses_numpy_global_rng():
    x = speech_like_clip(6000, np.random.default_rng(4))
    mag = np.abs(lc.stft(x, NFFT, HOP, WIN))
    np.random.seed(123)
    ref, rmse = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, 5, batched_fft=True)
    np.random.seed(123)
    wav, mse = synthesis.griffin_lim_v2(mag, WIN, HOP, NFFT, 5)
    assert rel_l2(wav, ref) < 1e-5 and abs(mse - rmse) / rmse < 1e-4
    np.random.seed(123)
    wav2 = synthesis.spectrogram_to_wav(mag, WIN, HOP, NFFT, 5)
    np.random.seed(123)
    wav3 = synthesis.spectrogram_to_wav(mag, WIN, HOP, NFFT, 5)
    assert np.array_equal(wav2, wav3)                  # deterministic
    assert rel_l2(wav, wav2) < 1e-6                    # the mse-reporting kernel variant agrees
    w0, m0 = synthesis.griffin_lim_v2(mag, WIN, HOP, NFFT, 0)
    assert m0 is None and w0.shape == ref.shape


def test_griffin_lim_batch_invariance_and_layouts():
    mags, angs = _case([40, 9, 100])
    batch = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 6, angles=angs)
    alone = synthesis.spectrograms_to_wavs([mags[2]], WIN, HOP, NFFT, 6, angles=[angs[2]])[0]
    assert np.array_equal(batch[2], alone)
    # float64, C-ordered (1025, T) input gives the same result as the float32 F-ordered view
    m64 = np.ascontiguousarray(mags[0].astype(np.float64))
    w64 = synthesis.spectrograms_to_wavs([m64], WIN, HOP, NFFT, 6, angles=[angs[0]])[0]
    assert np.array_equal(w64, batch[0])


def test_griffin_lim_sub_batch_pipeline_is_invisible(monkeypatch):
    """Large batches run as a pipeline of sub-batches (copy stream + events): neither the explicit-
    phase nor the seeded-device-phase results may depend on where the batch is split."""
    from single_speaker_tts_b200 import _runtime
    mags, angs = _case([40, 9, 100, 17, 64, 3])
    whole = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 4, angles=angs)
    whole_seeded, mse_whole = _runtime.griffin_lim_batch(mags, WIN, HOP, NFFT, 4, seed=99, return_mse=True)
    monkeypatch.setattr(_runtime, '_GL_CHUNK_FRAMES', 30)
    monkeypatch.setattr(_runtime, '_GL_CHUNK_GROWTH', 2)
    assert len(_runtime._split_by_frames([m.shape[1] for m in mags], 30, 2)) >= 3
    split = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 4, angles=angs)
    split_seeded, mse_split = _runtime.griffin_lim_batch(mags, WIN, HOP, NFFT, 4, seed=99, return_mse=True)
    for a, b, c, d in zip(whole, split, whole_seeded, split_seeded):
        assert np.array_equal(a, b) and np.array_equal(c, d)
    assert mse_whole == mse_split


def test_griffin_lim_dynamic_geometry_and_zero_bins():
    win, hop = 1024, 256
    x = speech_like_clip(hop * 40 + 3, np.random.default_rng(8))
    m = np.abs(lc.stft(x, NFFT, hop, win))
    a = np.exp(2j * np.pi * np.random.RandomState(1).rand(*m.shape))
    w = synthesis.spectrogram_to_wav(m, win, hop, NFFT, 3, angles=a)
    assert rel_l2(w, ra.spectrogram_to_wav(m, win, hop, NFFT, 3, angles=a)) < 5e-6
    mag = np.zeros((1025, 12), np.float32)
    mag[40, :] = 1.0
    ang = np.exp(2j * np.pi * np.random.RandomState(3).rand(1025, 12))
    w = synthesis.spectrogram_to_wav(mag, WIN, HOP, NFFT, 2, angles=ang)
    assert np.isfinite(w).all() and rel_l2(w, ra.spectrogram_to_wav(mag, WIN, HOP, NFFT, 2, angles=ang)) < 1e-4
    wz = synthesis.spectrogram_to_wav(np.zeros((1025, 9), np.float32), WIN, HOP, NFFT, 2)
    assert np.array_equal(wz, np.zeros(HOP * 8, np.float32))


def test_device_random_phase_is_seeded_and_unit():
    mags, _ = _case([20, 33])
    a = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 3, seed=7)
    b = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 3, seed=7)
    c = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 3, seed=8)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert not np.array_equal(a[0], c[0])
    np.random.seed(5); d = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 3)
    np.random.seed(5); e = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 3)
    assert all(np.array_equal(x, y) for x, y in zip(d, e))
    # the phase drawn inside the synthesis launch is the stream sstts_random_phase writes: feeding
    # that buffer back as explicit initial phasors reproduces the seeded result, and the oracle with
    # the same phasors agrees
    import ctypes
    lib = _lib.load()
    n_el = sum(m.shape[1] for m in mags) * 1025
    ph = torch.empty((n_el, 2), dtype=torch.float32, device='cuda')
    _lib.check(lib.sstts_random_phase(ctypes.c_uint64(7), n_el, ctypes.c_void_p(ph.data_ptr()),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    ph = ph.cpu().numpy().view(np.complex64).reshape(-1, 1025)
    assert np.abs(np.abs(ph) - 1).max() < 3e-6            # SFU sine / cosine
    assert abs(np.angle(ph).mean()) < 0.02 and abs(np.angle(ph).std() - np.pi / np.sqrt(3)) < 0.02   # uniform phase
    offs = np.concatenate([[0], np.cumsum([m.shape[1] for m in mags])])
    angs = [ph[offs[i]:offs[i + 1]].T for i in range(len(mags))]
    f = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 3, angles=angs)
    assert all(np.array_equal(x, y) for x, y in zip(a, f))
    ref = ra.spectrogram_to_wav(mags[1], WIN, HOP, NFFT, 3, angles=angs[1], batched_fft=True)
    assert rel_l2(a[1], ref) < 5e-6


def test_features_golden(golden_dir):
    f = np.load(golden_dir + '/features.npz')
    clips = [f['clip%d' % i] for i in range(3)]
    out = LJSpeechDatasetHelper.features_from_wavs(clips, sampling_rate=22050, trim_silence=False)
    for i, (mel, lin) in enumerate(out):
        assert mel.dtype == np.float32 and lin.dtype == np.float32
        assert mel.shape == f['mel%d' % i].shape and lin.shape == f['lin%d' % i].shape
        assert np.abs(mel - f['mel%d' % i]).max() <= NORM_TOL
        assert np.abs(lin - f['lin%d' % i]).max() <= NORM_TOL
    rows = statistics.decibel_statistics_batch(clips, 22050)
    for i in range(3):
        assert np.abs(rows[i] - f['stats%d' % i]).max() < 1e-3
    assert np.abs(statistics.collect_decibel_statistics_from_wavs(clips, 22050, batch_clips=2) - f['corpus_stats']).max() < 1e-3
    assert np.array_equal(statistics.reduce_decibel_statistics(rows),
                          statistics.collect_decibel_statistics_from_wavs(clips, 22050))


def test_features_ragged_edge_cases_vs_oracle():
    rng = np.random.default_rng(7)
    lens = [1, 2, 274, 275, 276, 1500, 5000, 30000]
    wavs = [speech_like_clip(max(n, 8), rng)[:n] for n in lens]
    wavs.append(np.zeros(4000, np.float32))                       # digital silence -> -100 dB floor
    out = LJSpeechDatasetHelper.features_from_wavs(wavs, sampling_rate=22050, trim_silence=False)
    for w, (mel, lin) in zip(wavs, out):
        mel_ref, lin_ref = ra.load_audio_from_wav(w, 22050, trim=False)
        assert mel.shape == mel_ref.shape and lin.shape == lin_ref.shape
        assert np.abs(mel - mel_ref).max() <= NORM_TOL and np.abs(lin - lin_ref).max() <= NORM_TOL
    # with the trim step of datasets/lj_speech.py:119
    x = np.concatenate([np.zeros(3000, np.float32), wavs[-2], np.zeros(3000, np.float32)])
    mel, lin = LJSpeechDatasetHelper.features_from_wavs([x], sampling_rate=22050)[0]
    mel_ref, lin_ref = ra.load_audio_from_wav(x, 22050, trim=True)
    assert mel.shape == mel_ref.shape and np.abs(lin - lin_ref).max() <= NORM_TOL


def test_features_public_functions_vs_oracle():
    x = speech_like_clip(12345, np.random.default_rng(11))
    S = features.linear_scale_spectrogram(x, NFFT, HOP, WIN)
    Sr = ra.linear_scale_spectrogram(x, NFFT, HOP, WIN)
    assert S.dtype == np.complex64 and S.shape == Sr.shape and S.flags['F_CONTIGUOUS']
    assert np.abs(S - Sr).max() / np.abs(Sr).max() < 2e-7
    S2 = features.linear_scale_spectrogram(x, 1024)                 # defaults: win = n_fft, hop = win // 4
    assert np.abs(S2 - ra.linear_scale_spectrogram(x, 1024)).max() / np.abs(Sr).max() < 2e-7
    for power in (1, 2):
        M = features.mel_scale_spectrogram(x, NFFT, 22050, 80, 0, 8000, HOP, WIN, power)
        Mr = ra.mel_scale_spectrogram(x, NFFT, 22050, 80, 0, 8000, HOP, WIN, power)
        assert M.dtype == np.float64 and M.shape == Mr.shape
        assert np.abs(M - Mr).max() / np.abs(Mr).max() < 1e-6
    st = statistics.decibel_statistics(x, 22050)
    assert np.abs(st - ra.decibel_statistics(x, 22050)).max() < 1e-3
    with pytest.raises(ValueError):
        features.linear_scale_spectrogram(np.zeros((2, 100), np.float32), NFFT, HOP, WIN)


def test_features_sub_batch_pipeline_and_fused_mode_are_invisible(monkeypatch):
    """features_batch runs large batches as a pipeline of sub-batches on side streams and picks the
    fused dB-feature kernel mode: per-clip results must equal the single-batch call bit for bit, and
    the fused mode must agree with the generic kernel mode to float32 rounding."""
    rng = np.random.default_rng(21)
    clips = [speech_like_clip(int(n), rng) for n in (9000, 300, 22050, 5000, 14000, 275, 31000)]
    consts = (35.66, 100.0, 6.02, 99.89)
    whole = features.features_batch(clips, NFFT, HOP, WIN, 22050, 80, 0, 8000, *consts, reduction=5)
    monkeypatch.setattr(_runtime, '_FEAT_CHUNK_SAMPLES', 10000)
    assert len(_runtime._split_by_frames([len(c) for c in clips], 10000)) >= 3
    split = features.features_batch(clips, NFFT, HOP, WIN, 22050, 80, 0, 8000, *consts, reduction=5)
    for (m0, l0), (m1, l1) in zip(whole, split):
        assert np.array_equal(m0, m1) and np.array_equal(l0, l1)
    for prec in ('f64', 'f32'):
        kw = dict(sampling_rate=22050, n_mels=80, fmin=0, fmax=8000, reduction=5, want_lin=True, want_mel=True,
                  normalize=consts, precision=prec)
        fused = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, **kw)
        generic = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, force_generic=True, **kw)
        assert np.abs(fused.lin_db - generic.lin_db).max() < 2e-6
        assert np.abs(fused.mel_db - generic.mel_db).max() < 2e-6
        raw_f = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, **dict(kw, normalize=None))
        raw_g = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, force_generic=True, **dict(kw, normalize=None))
        assert np.abs(raw_f.lin_db - raw_g.lin_db).max() < 3e-4      # dB, i.e. 2e-6 of the 135 dB range
        assert np.abs(raw_f.mel_db - raw_g.mel_db).max() < 3e-4


def test_features_f32_fast_mode_error_is_bounded():
    """precision='f32' is the documented fast mode: bins > ~100 dB under the frame peak differ."""
    x = speech_like_clip(40000, np.random.default_rng(12))
    mel_ref, lin_ref = ra.load_audio_from_wav(x, 22050, trim=False)
    mel, lin = LJSpeechDatasetHelper.features_from_wavs([x], 22050, trim_silence=False, precision='f32')[0]
    assert np.abs(mel - mel_ref).max() < 1e-5
    assert np.abs(lin - lin_ref).max() < 2e-2 and np.mean(np.abs(lin - lin_ref) > NORM_TOL) < 1e-3


def test_full_size_round_trip_properties():
    """BASELINE configs 1-2 at full size (256 ragged clips): STFT -> (|S|, true phase) -> Griffin-Lim
    must return the input (a consistent spectrogram is a fixed point of the iteration), every output
    obeys the hop * (T - 1) length rule, and the feature rows are in [0, 1] with zero pad rows."""
    clips = make_clips(256, seed=1, pool=8)
    res = LJSpeechDatasetHelper.features_from_wavs(clips, 22050, trim_silence=False)
    for c, (mel, lin) in zip(clips[:256:17], res[:256:17]):
        T = 1 + len(c) // HOP
        assert lin.shape == (-(-T // 5), 5125) and mel.shape == (-(-T // 5), 400)
        assert lin.min() >= 0 and lin.max() <= 1 and mel.min() >= 0 and mel.max() <= 1
        assert np.all(lin.reshape(-1, 1025)[T:] == 0) and np.all(mel.reshape(-1, 80)[T:] == 0)
    from single_speaker_tts_b200 import _runtime
    # whole number of hops: then hop * (T - 1) == len(clip) and the re-analysis reflects at the same place
    sub = [c[:HOP * (len(c) // HOP)] for c in clips[:64]]
    spec = _runtime.stft_features_batch(sub, NFFT, HOP, WIN, want_spec=True, precision='f64')
    mags, angs = [], []
    for i in range(len(sub)):
        S = spec.rows(spec.spec, i).T
        mags.append(np.abs(S))
        angs.append(np.where(np.abs(S) > 0, S / np.maximum(np.abs(S), 1e-30), 1.0))
    for n_iter in (0, 5):
        wavs = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, n_iter, angles=angs)
        for c, w in zip(sub, wavs):
            assert w.shape == (HOP * (len(c) // HOP),)
            assert rel_l2(w, c[:len(w)]) < 2e-5


def test_model_output_glue_matches_inference_recipe(golden_dir):
    """tacotron/inference.py:94-101,175 fused on the device == the oracle recipe + Griffin-Lim."""
    g = np.load(golden_dir + '/gl_fixture.npz')
    out = g['model_output']                                   # (T, 1025) normalised, as session.run returns
    mag = ra.inference_postprocess(out)
    np.random.seed(11)
    ref = synthesis.spectrograms_to_wavs([mag, mag[:, :50]], WIN, HOP, NFFT, 8, seed=99)
    got = synthesis.model_outputs_to_wavs([out, out[:50]], 6.02, 99.89, 1.3, WIN, HOP, NFFT, 8, seed=99)
    for a, b in zip(got, ref):
        assert a.shape == b.shape and rel_l2(a, b) < 1e-4
    with pytest.raises(AssertionError, match='smaller -100 dB'):
        synthesis.model_outputs_to_wavs([np.zeros((20, 1025), np.float32)], 6.02, 120.0, 1.3, WIN, HOP, NFFT, 1)


def test_trim_batch_matches_oracle():
    from single_speaker_tts_b200.audio import effects
    rng = np.random.default_rng(3)
    sp = speech_like_clip(30000, rng)
    wavs = [np.concatenate([np.zeros(4000, np.float32), sp, np.zeros(5000, np.float32)]), sp[:3000].copy(),
            np.concatenate([1e-5 * rng.normal(size=3000).astype(np.float32), sp[:7000], 1e-6 * np.ones(2500, np.float32)])]
    trimmed, bounds = effects.trim_batch(wavs)
    for w, t, b in zip(wavs, trimmed, bounds):
        yr, br = lc.trim(w)
        assert tuple(b) == tuple(br) and np.array_equal(t, yr)
        assert tuple(effects.trim(w)[1]) == tuple(br)


def test_file_level_drivers_precalc_and_statistics(tmp_path):
    """tacotron/dataset_precalc_features.py and dataset_statistics.py paths end to end on wav files:
    pre_compute_features (datasets/dataset_helper.py:326-355) writes <name>.npz with mel_mag_db /
    linear_mag_db, collect_decibel_statistics (datasets/statistics.py:69-98) averages per-file extrema."""
    from scipy.io import wavfile
    from single_speaker_tts_b200.audio.io import load_wav
    rng = np.random.default_rng(21)
    paths = []
    for i, n in enumerate((30000, 12000, 50000)):
        x = np.concatenate([np.zeros(3000, np.float32), speech_like_clip(n, rng), np.zeros(2000, np.float32)])
        path = str(tmp_path / ('clip%d.wav' % i))
        wavfile.write(path, 22050, (x * 32767).astype(np.int16))
        paths.append(path)
    LJSpeechDatasetHelper.pre_compute_features(paths, batch_clips=2)
    for path in paths:
        wav, sr = load_wav(path)
        assert sr == 22050 and wav.dtype == np.float32
        mel_ref, lin_ref = ra.load_audio_from_wav(wav, sr, trim=True)
        data = np.load(path[:-4] + '.npz')
        assert data['mel_mag_db'].shape == mel_ref.shape and data['linear_mag_db'].shape == lin_ref.shape
        assert np.abs(data['mel_mag_db'] - mel_ref).max() <= NORM_TOL
        assert np.abs(data['linear_mag_db'] - lin_ref).max() <= NORM_TOL
        mel, lin = LJSpeechDatasetHelper.load_audio(path.encode())
        assert np.array_equal(mel, data['mel_mag_db']) and np.array_equal(lin, data['linear_mag_db'])
    stats = statistics.collect_decibel_statistics(paths, batch_clips=2)
    ref = ra.collect_decibel_statistics_from_wavs([load_wav(p)[0] for p in paths], 22050)
    assert np.abs(stats - ref).max() < 1e-3


def test_sibling_corpus_helpers_match_their_reference_recipes():
    """datasets/blizzard_nancy.py, cmu_slt.py (LJSpeech recipe, other constants) and pavoque.py
    (zeroed low bins + spectrogram-based row slicing) against the oracle restatements."""
    rng = np.random.default_rng(31)
    clips = [speech_like_clip(int(n), rng) for n in (15000, 30011)]
    for helper, consts in ((BlizzardNancyDatasetHelper, ra.BlizzardNancyConstants), (CMUDatasetHelper, ra.CMUConstants)):
        got = helper.features_from_wavs(clips, sampling_rate=22050)
        for c, (mel, lin) in zip(clips, got):
            mel_ref, lin_ref = ra.load_audio_from_wav(c, 22050, constants=consts)
            assert mel.shape == mel_ref.shape and lin.shape == lin_ref.shape
            assert np.abs(mel - mel_ref).max() < NORM_TOL and np.abs(lin - lin_ref).max() < NORM_TOL
    got = PAVOQUEDatasetHelper.features_from_wavs(clips, sampling_rate=22050)
    for c, (mel, lin) in zip(clips, got):
        mel_ref, lin_ref = ra.pavoque_load_audio_from_wav(c, 22050)
        assert mel.shape == mel_ref.shape and lin.shape == lin_ref.shape and mel.dtype == np.float32
        assert np.abs(mel - mel_ref).max() < NORM_TOL and np.abs(lin - lin_ref).max() < NORM_TOL


def test_reconstruction_error_study_matches_oracle():
    """collect_reconstruction_error (datasets/statistics.py:146-187): per-clip MSE after n iterations;
    compared through griffin_lim_batch with explicit initial phases (the seeded device generator has
    no numpy equivalent), plus the seeded batch entry point for determinism."""
    from single_speaker_tts_b200 import _runtime as rt
    rng = np.random.default_rng(33)
    clips = [speech_like_clip(int(n), rng) for n in (9000, 14000)]
    win, hop = 1102, 275
    for c in clips:
        mag = np.abs(lc.stft(c, NFFT, hop, win))
        ang = np.exp(2j * np.pi * np.random.RandomState(5).rand(*mag.shape))
        ref = ra.reconstruction_error(c, 22050, 4, angles=ang)
        _, mses = rt.griffin_lim_batch([mag], win, hop, NFFT, 4, angles=[ang], return_mse=True)
        assert abs(mses[0] - ref) / ref < 1e-4
    a = statistics.reconstruction_errors_from_wavs(clips, 22050, 3, seed=11)
    b = statistics.reconstruction_errors_from_wavs(clips, 22050, 3, seed=11)
    assert a == b and all(np.isfinite(v) and v > 0 for v in a)


def test_pinned_inputs_take_the_zero_copy_upload_and_give_identical_results():
    import single_speaker_tts_b200 as pkg
    from single_speaker_tts_b200 import _hostio
    mags, angs = _case([40, 9, 100])
    pinned = []
    for m in mags:
        buf = pkg.pinned_empty((m.shape[1], m.shape[0]))      # (T, bins) C-ordered, like the model output
        buf[:] = m.T
        pinned.append(buf.T)                                   # the reference's (bins, T) view
    assert _hostio._direct_sources([p.T for p in pinned], torch.float32) is not None
    assert _hostio._direct_sources([np.ascontiguousarray(m.T) for m in mags], torch.float32) is None
    a = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 3, seed=5)
    b = synthesis.spectrograms_to_wavs(pinned, WIN, HOP, NFFT, 3, seed=5)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    rng = np.random.default_rng(4)
    clips = [speech_like_clip(int(n), rng) for n in (7000, 12000)]
    pclips = []
    for c in clips:
        buf = pkg.pinned_empty((len(c),))
        buf[:] = c
        pclips.append(buf)
    consts = (35.66, 100.0, 6.02, 99.89)
    fa = features.features_batch(clips, NFFT, HOP, WIN, 22050, 80, 0, 8000, *consts, reduction=5)
    fb = features.features_batch(pclips, NFFT, HOP, WIN, 22050, 80, 0, 8000, *consts, reduction=5)
    for (m0, l0), (m1, l1) in zip(fa, fb):
        assert np.array_equal(m0, m1) and np.array_equal(l0, l1)


def test_griffin_lim_n_fft_1024_time_stretch_geometry():
    """audio/effects.py:71-86 reconstructs with n_fft = win = 1024, hop = 256, 25 iterations."""
    n_fft, win, hop = 1024, 1024, 256
    rng = np.random.default_rng(44)
    x = speech_like_clip(hop * 60 + 11, rng)
    m = np.abs(lc.stft(x, n_fft, hop, win))
    a = np.exp(2j * np.pi * np.random.RandomState(9).rand(*m.shape))
    w, mse = synthesis.griffin_lim_v2(m, win, hop, n_fft, 25, angles=a)
    ref, rmse = ra.griffin_lim_v2(m, win, hop, n_fft, 25, angles=a, batched_fft=True)
    assert w.shape == ref.shape == (hop * (m.shape[1] - 1),)
    assert rel_l2(w, ref) < GL_TOL and abs(mse - rmse) / rmse < 1e-3
    w512 = synthesis.spectrogram_to_wav(np.abs(lc.stft(x, 512, 128, 512)), 512, 128, 512, 3,
                                        angles=np.exp(2j * np.pi * np.random.RandomState(2).rand(257, 1 + len(x) // 128)))
    r512 = ra.spectrogram_to_wav(np.abs(lc.stft(x, 512, 128, 512)), 512, 128, 512, 3,
                                 angles=np.exp(2j * np.pi * np.random.RandomState(2).rand(257, 1 + len(x) // 128)))
    assert rel_l2(w512, r512) < 1e-5


@pytest.mark.parametrize('precision,tol', [('f32', 1e-5), ('f64', 1e-6)])
def test_griffin_lim_native_1024_transform_vs_embedded_and_oracle(precision, tol, monkeypatch):
    """n_fft 1024 Griffin-Lim runs on the native 512-point complex transform (two frames per warp, 16-frame
    tiles); SSTTS_GL_NATIVE1024=0 keeps the embedding in the 2048-point transform.  Both must match the
    oracle on a ragged batch with single-frame, two-frame, odd and multi-tile utterances, for the
    statistics / time-stretch geometry (compile-time) and a run-time one (shorter window)."""
    rng = np.random.default_rng(77)
    for win, hop in ((1024, 256), (800, 200)):
        frames = [1, 2, 3, 7, 16, 17, 33, 50, 121]
        mags, angs = [], []
        for i, T in enumerate(frames):
            x = speech_like_clip(hop * (T - 1) + 5, rng)
            m = np.abs(lc.stft(x, 1024, hop, win))
            mags.append(m)
            angs.append(np.exp(2j * np.pi * np.random.RandomState(i).rand(*m.shape)))
        out = {}
        for native in ('1', '0'):
            monkeypatch.setenv('SSTTS_GL_NATIVE1024', native)
            wavs, mses = _runtime.griffin_lim_batch(mags, win, hop, 1024, 4, angles=angs, return_mse=True,
                                                    precision=precision)
            out[native] = wavs
            for T, m, a, w, mse in zip(frames, mags, angs, wavs, mses):
                assert w.shape == (hop * (T - 1),)
                if T == 1:
                    continue
                ref, rmse = ra.griffin_lim_v2(m, win, hop, 1024, 4, angles=a, batched_fft=True)
                assert rel_l2(w, ref) < tol, (native, win, hop, T)
                assert abs(mse - rmse) / rmse < 1e-4
        # two different kernels: equal up to rounding, not bit for bit (float64 results round to the same float32)
        if precision == 'f32':
            assert any(not np.array_equal(a, b) for a, b in zip(out['1'], out['0']))


def test_time_stretch_matches_the_reference_recipe():
    """audio/effects.py:46-86: interpolated |STFT| (the part of the phase vocoder the reference keeps)
    and the 25-iteration n_fft-1024 reconstruction, with the same numpy-seeded initial phase."""
    from single_speaker_tts_b200.audio import effects
    rng = np.random.default_rng(51)
    x = speech_like_clip(256 * 50 + 33, rng)
    S = lc.stft(x, 1024, 256, 1024)
    for rate in (0.8, 1.0, 1.25, 2.3):
        got = effects.stretch_magnitude(S, rate)
        ref = np.abs(lc.phase_vocoder(S, rate))
        assert got.shape == ref.shape and got.dtype == np.float32
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    for rate in (0.8, 1.25):
        np.random.seed(3)
        w = effects.time_stretch(x, rate)
        T_out = len(np.arange(0, S.shape[1], rate))
        np.random.seed(3)
        ang = np.exp(2j * np.pi * np.random.rand(513, T_out))
        ref = ra.time_stretch(x, rate, angles=ang)
        assert w.shape == ref.shape == (256 * (T_out - 1),) and w.dtype == np.float32
        assert rel_l2(w, ref) < GL_TOL
    with pytest.raises(ValueError):
        effects.time_stretch(x, 0.0)


def test_calculate_mfccs_matches_oracle():
    rng = np.random.default_rng(61)
    x = speech_like_clip(9000, rng)
    mel = ra.mel_scale_spectrogram(x, NFFT, 22050, 80, 0, 8000, HOP, WIN, 1)
    logmel = 10.0 * np.log10(np.maximum(1e-10, mel))
    for n_mfcc in (13, 20, 80):
        got = features.calculate_mfccs(logmel, 22050, n_mfcc)
        ref = ra.calculate_mfccs(logmel, 22050, n_mfcc)
        assert got.shape == ref.shape == (n_mfcc, mel.shape[1]) and got.dtype == np.float64
        assert np.abs(got - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())


def test_concurrent_callers_get_the_serial_results():
    """tacotron/serve.py:69-72 calls the per-item function from a ThreadPool(6), the TF input pipeline
    calls load_audio from >= 4 queue-runner threads (tacotron/params/training.py:14): the library and the
    host runtime must be re-entrant (locked plan / table caches, thread-local staging and streams)."""
    from concurrent.futures import ThreadPoolExecutor
    mags, angs = _case([30, 5, 55, 9, 41, 2])      # spans (hence dynamic shared memory sizes) differ per caller
    rng = np.random.default_rng(71)
    clips = [speech_like_clip(int(n), rng) for n in (5000, 9000, 7000, 12000, 3000, 8000)]

    def synth(i):
        return synthesis.spectrogram_to_wav(mags[i], WIN, HOP, NFFT, 4, angles=angs[i])

    def feats(i):
        return LJSpeechDatasetHelper.features_from_wavs([clips[i]], sampling_rate=22050)[0]

    serial_w = [synth(i) for i in range(6)]
    serial_f = [feats(i) for i in range(6)]
    for _ in range(10):
        with ThreadPoolExecutor(max_workers=6) as ex:
            fw = [ex.submit(synth, i) for i in range(6)]
            ff = [ex.submit(feats, i) for i in range(6)]
            got_w = [f.result() for f in fw]
            got_f = [f.result() for f in ff]
        for a, b in zip(serial_w, got_w):
            assert np.array_equal(a, b)
        for (m0, l0), (m1, l1) in zip(serial_f, got_f):
            assert np.array_equal(m0, m1) and np.array_equal(l0, l1)


def test_pcm16_clips_are_decoded_on_the_device(tmp_path):
    """N2: int16 clips (the LJSpeech file format) are uploaded as 2-byte samples and converted on the
    device exactly like load_wav converts them on the host -- identical features, with and without the
    trim step, also for a mixed int16 / float32 list and through the file-level driver."""
    from scipy.io import wavfile
    from single_speaker_tts_b200.audio import io as aio
    rng = np.random.default_rng(81)
    pcm = [np.round(speech_like_clip(int(n), rng) * 30000).astype(np.int16) for n in (9000, 15000, 4000)]
    flt = [p.astype(np.float32) / 32768.0 for p in pcm]
    for trim_silence in (False, True):
        a = LJSpeechDatasetHelper.features_from_wavs(pcm, 22050, trim_silence=trim_silence)
        b = LJSpeechDatasetHelper.features_from_wavs(flt, 22050, trim_silence=trim_silence)
        c = LJSpeechDatasetHelper.features_from_wavs([pcm[0], flt[1], pcm[2]], 22050, trim_silence=trim_silence)
        for (m0, l0), (m1, l1), (m2, l2) in zip(a, b, c):
            assert np.array_equal(m0, m1) and np.array_equal(l0, l1)
            assert np.array_equal(m0, m2) and np.array_equal(l0, l2)
    paths = []
    for i, p in enumerate(pcm):
        path = str(tmp_path / ('u%d.wav' % i))
        wavfile.write(path, 22050, p)
        paths.append(path)
    assert aio.load_wav_pcm16(paths[0])[0].dtype == np.int16
    LJSpeechDatasetHelper.pre_compute_features(paths)
    for path, (mel, lin) in zip(paths, LJSpeechDatasetHelper.features_from_wavs(flt, 22050)):
        z = np.load(path[:-4] + '.npz')
        assert np.array_equal(z['mel_mag_db'], mel) and np.array_equal(z['linear_mag_db'], lin)
    s1 = statistics.collect_decibel_statistics(paths)
    s2 = statistics.collect_decibel_statistics_from_wavs(flt, 22050)
    assert np.array_equal(s1, s2)


def test_peak_normalisation_matches_save_wav_norm():
    """save_wav(..., norm=True) (audio/io.py:33-53, used at tacotron/inference.py:199) divides by the
    peak; the device version must give the same float32 values (true division), utterance by utterance,
    also across sub-batches."""
    mags, angs = _case([40, 9, 100, 17])
    plain = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 3, angles=angs)
    normed = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 3, angles=angs, normalize_peak=True)
    for w, n in zip(plain, normed):
        ref = w / np.max(np.abs(w))
        assert n.dtype == np.float32 and np.array_equal(n, ref.astype(np.float32))
        assert abs(float(np.max(np.abs(n))) - 1.0) < 1e-6
    z = synthesis.spectrograms_to_wavs([np.zeros((1025, 6), np.float32)], WIN, HOP, NFFT, 1, seed=1, normalize_peak=True)
    assert np.array_equal(z[0], np.zeros(HOP * 5, np.float32))


def test_corpus_pass_equals_the_two_separate_drivers():
    """BASELINE configs[3] on one rank: distributed.corpus_pass (one upload, statistics kernel, reduction,
    pre-calculation on the resident clips, pipelined downloads) must produce exactly the statistics of
    collect_decibel_statistics_from_wavs and, with them as constants, exactly the features of features_batch
    -- also when the shard is split into several device batches."""
    from single_speaker_tts_b200 import distributed
    rng = np.random.default_rng(91)
    clips = [speech_like_clip(int(n), rng) for n in (9000, 300, 22050, 5000, 14000, 275, 31000, 1200)]
    idx = list(range(len(clips)))
    want_stats = statistics.collect_decibel_statistics_from_wavs(clips, 22050)
    lin_max, lin_ref, mel_max, mel_ref = want_stats
    want = features.features_batch(clips, NFFT, HOP, WIN, 22050, 80, 0, 8000, lin_ref, lin_max, mel_ref, mel_max, reduction=5)
    for chunk in (1 << 30, 20000):
        got = {}

        def sink(indices, part):
            for j, i in enumerate(indices):
                got[i] = (part.rows(part.mel_db, j, padded=True).reshape(-1, 400).copy(),
                          part.rows(part.lin_db, j, padded=True).reshape(-1, 5125).copy())

        mean4, n_rows = distributed.corpus_pass(clips, idx, len(clips), 22050, NFFT, HOP, WIN, 80, 0, 8000,
                                                reduction=5, chunk_samples=chunk, sink=sink)
        assert np.array_equal(mean4, want_stats)
        assert n_rows == sum(m.shape[0] * 5 for m, _ in want)
        for i, (mel, lin) in enumerate(want):
            assert np.array_equal(got[i][0], mel) and np.array_equal(got[i][1], lin)


def test_bulk_copy_staging_variant_matches_the_default(tmp_path):
    """SSTTS_GL_STAGING=bulk runs the iteration kernel whose interior tiles are staged with cp.async.bulk +
    mbarrier (read once per process, hence a subprocess): same waveforms as the default staging up to the
    rounding of the folded window x normalisation table, and inside the oracle tolerance after 50 iterations."""
    import os
    import subprocess
    import sys
    mags, angs = _case([40, 9, 100, 333])
    np.savez(str(tmp_path / 'case.npz'), **{'m%d' % i: m for i, m in enumerate(mags)}, **{'a%d' % i: a for i, a in enumerate(angs)})
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from single_speaker_tts_b200.audio import synthesis\n"
        "z = np.load(%r)\n"
        "mags = [z['m%%d' %% i] for i in range(4)]; angs = [z['a%%d' %% i] for i in range(4)]\n"
        "w = synthesis.spectrograms_to_wavs(mags, 1102, 275, 2048, 50, angles=angs)\n"
        "np.savez(%r, *w)\n" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / 'case.npz'),
                                 str(tmp_path / 'out.npz')))
    env = dict(os.environ, SSTTS_GL_STAGING='bulk')
    subprocess.run([sys.executable, '-c', code], check=True, env=env, timeout=600)
    got = np.load(str(tmp_path / 'out.npz'))
    ref = synthesis.spectrograms_to_wavs(mags, WIN, HOP, NFFT, 50, angles=angs)
    for i, (m, a) in enumerate(zip(mags, angs)):
        w = got['arr_%d' % i]
        assert w.shape == ref[i].shape and rel_l2(w, ref[i]) < 1e-4
    orc = ra.spectrogram_to_wav(mags[2], WIN, HOP, NFFT, 50, angles=angs[2], batched_fft=True)
    assert rel_l2(got['arr_2'], orc) <= GL_TOL


def test_random_supported_geometries_vs_oracle():
    """Run-time geometries across the supported set (the wrappers accept a subset of what librosa
    accepts -- everything inside that subset must match the oracle): STFT / dB features for n_fft 512, 1024
    (native transform) and 2048 with assorted window / hop lengths, Griffin-Lim for those with
    win / 5 <= hop <= win; geometries outside raise ValueError before touching the device."""
    rng = np.random.default_rng(123)
    x = speech_like_clip(9000, rng)
    cases = [(2048, 2048, 512), (2048, 1500, 375), (2048, 800, 800), (2048, 1102, 1400), (1024, 1024, 256), (1024, 800, 200),
             (1024, 512, 512), (1024, 1000, 1023), (512, 512, 128), (512, 400, 100), (512, 256, 300)]
    for n_fft, win, hop in cases:
        S = features.linear_scale_spectrogram(x, n_fft, hop, win)
        Sr = ra.linear_scale_spectrogram(x, n_fft, hop, win)
        assert S.shape == Sr.shape, (n_fft, win, hop)
        assert np.abs(S - Sr).max() / np.abs(Sr).max() < 2e-7, (n_fft, win, hop)
        M = features.mel_scale_spectrogram(x, n_fft, 22050, 40, 50, 7000, hop, win, 1)
        Mr = ra.mel_scale_spectrogram(x, n_fft, 22050, 40, 50, 7000, hop, win, 1)
        assert np.abs(M - Mr).max() / np.abs(Mr).max() < 1e-6, (n_fft, win, hop)
        if hop <= win and -(-win // hop) <= 5:
            m = np.abs(Sr)
            a = np.exp(2j * np.pi * np.random.RandomState(n_fft + win + hop).rand(*m.shape))
            w, mse = synthesis.griffin_lim_v2(m, win, hop, n_fft, 3, angles=a)
            ref, rmse = ra.griffin_lim_v2(m, win, hop, n_fft, 3, angles=a, batched_fft=True)
            assert w.shape == ref.shape and rel_l2(w, ref) < 1e-5, (n_fft, win, hop)
            assert abs(mse - rmse) / rmse < 1e-4
        else:
            with pytest.raises(ValueError, match='Griffin-Lim'):
                synthesis.spectrogram_to_wav(np.abs(Sr), win, hop, n_fft, 1)
    for bad in ((4096, 4096, 1024), (2048, 1101, 275), (1024, 1024, 2000)):
        with pytest.raises(ValueError):
            features.linear_scale_spectrogram(x, *bad[:1], bad[2], bad[1])
