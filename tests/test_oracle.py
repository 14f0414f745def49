"""Pin the CPU oracle (oracle/) -- the checker every parity test relies on.

The reference holds no test or golden vector for the audio path and cannot run here (librosa /
TensorFlow 1.8 missing), so the oracle is pinned against independent implementations
(torch.stft / torch.istft, torchaudio's filterbank, numpy.fft), against the identities the
reference documents, and against the committed golden vectors (regression).
"""
import numpy as np
import pytest
import torch

from oracle import librosa_compat as lc
from oracle import reference_audio as ra
from single_speaker_tts_b200.synthetic import make_clips, speech_like_clip

WIN, HOP, NFFT = 1102, 275, 2048


def _torch_window(win, n_fft):
    w = torch.zeros(n_fft, dtype=torch.float64)
    lpad = (n_fft - win) // 2
    w[lpad:lpad + win] = torch.hann_window(win, periodic=True, dtype=torch.float64)
    return w


@pytest.mark.parametrize('n_fft,win,hop', [(2048, 1102, 275), (1024, 1024, 256)])
def test_stft_matches_torch(n_fft, win, hop):
    x = speech_like_clip(30000, np.random.default_rng(0))
    S = lc.stft(x, n_fft, hop, win)
    assert S.dtype == np.complex64 and S.flags['F_CONTIGUOUS']
    assert S.shape == (n_fft // 2 + 1, 1 + len(x) // hop)
    St = torch.stft(torch.from_numpy(x).double(), n_fft, hop, n_fft, window=_torch_window(win, n_fft),
                    center=True, pad_mode='reflect', return_complex=True).numpy()
    assert np.abs(S - St).max() / np.abs(St).max() < 2e-7  # complex64 storage


def test_istft_matches_torch_and_length_rule():
    x = speech_like_clip(20000, np.random.default_rng(1))
    S = lc.stft(x, NFFT, HOP, WIN).astype(np.complex128)
    y = lc.istft(S, HOP, WIN)
    T = S.shape[1]
    assert y.dtype == np.float32 and y.shape == (HOP * (T - 1),)     # readme wavs: 200 * (T - 1)
    yt = torch.istft(torch.from_numpy(S), NFFT, HOP, NFFT, window=_torch_window(WIN, NFFT),
                     center=True).numpy()
    assert np.abs(yt - y).max() < 5e-7
    assert np.abs(y - x[:len(y)]).max() < 5e-7                        # COLA round trip
    yb = lc.istft(S, HOP, WIN, batched_fft=True)
    assert np.abs(yb - y).max() < 1e-7


def test_per_frame_fft_matches_numpy():
    x = speech_like_clip(5000, np.random.default_rng(2))
    S = lc.stft(x, NFFT, HOP, WIN)
    xp = np.pad(x, NFFT // 2, mode='reflect')
    w = lc.padded_window(WIN, NFFT)
    for t in (0, 3, S.shape[1] - 1):
        ref = np.fft.rfft(w * xp[t * HOP:t * HOP + NFFT])
        assert np.abs(S[:, t] - ref).max() / np.abs(ref).max() < 2e-7


def test_window_is_periodic_hann_centred():
    w = lc.padded_window(WIN, NFFT)
    assert w.shape == (NFFT,) and w[:473].max() == 0 and w[473 + WIN:].max() == 0
    tw = torch.hann_window(WIN, periodic=True, dtype=torch.float64).numpy()
    assert np.abs(w[473:473 + WIN] - tw).max() < 1e-15
    wss = lc.window_sumsquare(50, HOP, WIN, NFFT)
    assert wss.dtype == np.float32 and wss[NFFT // 2:-NFFT // 2].min() > 1.2


def test_mel_filterbank_matches_torchaudio():
    ta = pytest.importorskip('torchaudio')
    for n_fft, fmax in ((2048, 8000.0), (1024, 11025.0)):
        mb = lc.mel_filterbank(22050, n_fft, 80, 0, fmax)
        fb = ta.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, fmax, 80, 22050, norm='slaney',
                                           mel_scale='htk').numpy().T
        assert mb.shape == (80, n_fft // 2 + 1)
        assert np.abs(mb - fb).max() < 1e-5 * np.abs(mb).max()      # torchaudio builds it in float32
    mb = lc.mel_filterbank(22050, 2048, 80, 0, 8000)
    assert (mb != 0).sum() == 1459 and np.nonzero(mb.any(0))[0].max() == 743


def test_conversion_identities():
    # audio/conversion.py:19,26-27
    assert ra.magnitude_to_decibel(np.array([1.0]))[0] == 0.0
    assert abs(ra.magnitude_to_decibel(np.array([1e-5]))[0] + 100.0) < 1e-9
    assert abs(ra.magnitude_to_decibel(np.array([0.0]))[0] + 100.0) < 1e-9
    # tacotron/params/model.py:13-24 through audio/conversion.py:136
    assert ra.ms_to_samples(50.0, 22050) == 1102 and ra.ms_to_samples(12.5, 22050) == 275
    db = np.linspace(-100, 40, 29)
    n = ra.normalize_decibel(db, 35.66, 100.0)
    assert n.min() >= 0 and n.max() <= 1
    inside = (n > 0) & (n < 1)
    assert np.abs(ra.inv_normalize_decibel(n, 35.66, 100.0)[inside] - db[inside]).max() < 1e-10
    with pytest.raises(AssertionError):
        ra.decibel_to_magnitude(np.array([-100.5]))
    x = np.float32(0.25)
    assert ra.magnitude_to_decibel(np.array([x], dtype=np.float32)).dtype == np.float32


def test_reduction_padding_shapes():
    mel = np.ones((7, 80), np.float32)
    lin = np.ones((7, 1025), np.float32)
    m2, l2 = ra.apply_reduction_padding(mel, lin, 5)
    assert m2.shape == (2, 400) and l2.shape == (2, 5125)
    assert m2.reshape(-1, 80)[7:].max() == 0 and l2.reshape(-1, 1025)[7:].max() == 0


def test_trim_removes_silence():
    x = np.concatenate([np.zeros(4000, np.float32), speech_like_clip(12000, np.random.default_rng(3)),
                        np.zeros(5000, np.float32)])
    y, (s, e) = lc.trim(x)
    assert 0 < s <= 4000 + 4096 and len(x) - 5000 - 4096 <= e < len(x) and len(y) == e - s


def test_griffin_lim_properties():
    x = speech_like_clip(4000, np.random.default_rng(4))
    mag = np.abs(lc.stft(x, NFFT, HOP, WIN))
    np.random.seed(7)
    w1, mse1 = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, 3)
    np.random.seed(7)
    ang = np.exp(2j * np.pi * np.random.rand(*mag.shape))
    w2, mse2 = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, 3, angles=ang)
    assert np.array_equal(w1, w2) and mse1 == mse2          # global-RNG semantics (synthesis.py:85)
    assert w1.dtype == np.float32 and w1.shape == (HOP * (mag.shape[1] - 1),)
    w0, mse0 = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, 0, angles=ang)
    assert mse0 is None
    # the spectral error must fall with iterations
    _, mse10 = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, 10, angles=ang)
    assert mse10 < mse1


def test_golden_vectors_regression(golden_dir):
    g = np.load(golden_dir + '/gl_synthetic.npz')
    c = g['clip1']
    mag = np.abs(lc.stft(c, NFFT, HOP, WIN))
    ang = np.exp(2j * np.pi * np.random.RandomState(int(g['seed0']) + 1).rand(*mag.shape))
    wav, mse = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, int(g['n_iter']), angles=ang)
    assert np.linalg.norm(wav - g['wav1']) / np.linalg.norm(g['wav1']) < 1e-5
    f = np.load(golden_dir + '/features.npz')
    mel, lin = ra.load_audio_from_wav(f['clip1'], 22050, trim=False)
    assert np.abs(mel - f['mel1']).max() < 1e-6 and np.abs(lin - f['lin1']).max() < 1e-6
    assert np.abs(ra.decibel_statistics(f['clip1'], 22050) - f['stats1']).max() < 1e-4
    clips = [f['clip%d' % i] for i in range(3)]
    assert np.abs(ra.collect_decibel_statistics_from_wavs(clips, 22050) - f['corpus_stats']).max() < 1e-4


def test_fixture_recipe(golden_dir):
    g = np.load(golden_dir + '/gl_fixture.npz')
    mag = ra.inference_postprocess(g['model_output'])
    assert mag.shape == (1025, 160) and mag.dtype == np.float32 and mag.min() > 0
    assert g['wav'].shape == (HOP * 159,)                    # 275 * (T - 1)


def test_real_shape_goldens_regression(golden_dir):
    """The goldens at the shapes the reference really runs (BASELINE configs[0]: T = 401; the full
    1000-frame model output of tacotron/params/model.py:108; 100 iterations as in configs[4]) are
    what this oracle produces today, and obey the known answers: T = 1 + N // hop, length hop * (T - 1)."""
    g = np.load(golden_dir + '/gl_config0.npz')
    mag = np.abs(lc.stft(g['clip'], NFFT, HOP, WIN))
    assert mag.shape == (1025, 401) and g['wav'].shape == (110000,)
    ang = np.exp(2j * np.pi * np.random.RandomState(int(g['seed'])).rand(*mag.shape))
    wav, mse = ra.griffin_lim_v2(mag, WIN, HOP, NFFT, int(g['n_iter']), angles=ang, batched_fft=True)
    assert np.linalg.norm(wav - g['wav']) / np.linalg.norm(g['wav']) < 1e-5
    assert abs(mse - float(g['mse'])) / float(g['mse']) < 1e-5
    f = np.load(golden_dir + '/gl_fixture_full.npz')
    assert f['model_output'].shape == (1000, 1025) and f['wav'].shape == (274725,)     # SURVEY 8c known answer
    mag = ra.inference_postprocess(f['model_output'])
    assert mag.shape == (1025, 1000) and mag.dtype == np.float32 and mag.min() > 0
    h = np.load(golden_dir + '/gl_100it.npz')
    assert int(h['n_iter']) == 100
    for i in range(2):
        assert h['wav%d' % i].shape == (HOP * (len(h['clip%d' % i]) // HOP),)


def test_single_frame_spectrogram_raises_like_the_reference():
    """audio/synthesis.py:96-106 with T = 1: istft returns an empty signal and the stft's reflect
    padding of an empty array raises ValueError in numpy; n_iter = 0 returns (empty, None)."""
    m = np.ones((1025, 1), np.float32)
    with pytest.raises(ValueError):
        ra.griffin_lim_v2(m, WIN, HOP, NFFT, 1)
    w, mse = ra.griffin_lim_v2(m, WIN, HOP, NFFT, 0)
    assert w.shape == (0,) and mse is None


def test_pavoque_recipe_restatement_shapes_and_floor():
    """datasets/pavoque.py:104-160: zeroed bins sit on the normalised -100 dB floor (0.0 with
    ref 24 / max 100) and the row slice comes from silence_interval_from_spectrogram as written."""
    from oracle import reference_audio as ra
    from single_speaker_tts_b200.synthetic import speech_like_clip
    wav = speech_like_clip(12000, np.random.default_rng(2))
    mel, lin = ra.pavoque_load_audio_from_wav(wav, 22050)
    assert mel.dtype == np.float32 and lin.dtype == np.float32
    assert mel.shape[0] == lin.shape[0] and mel.shape[1] == 400 and lin.shape[1] == 5125
    rows = lin.reshape(-1, 1025)
    assert np.all(rows[:, 0:8] == 0.0)
    spec = np.zeros((5, 7)); spec[2, 3] = 1.0; spec[4, 5] = 1.0
    assert ra.silence_interval_from_spectrogram(spec, 0.5) == (3, 5)
    assert ra.silence_interval_from_spectrogram(spec, 2.0) is None


def test_phase_vocoder_restatement():
    """librosa.core.phase_vocoder (audio/effects.py:77): rate 1 reproduces the magnitudes and the frame
    count; other rates interpolate |D| linearly at the fractional positions; a stationary sinusoid keeps
    its bin and magnitude."""
    from oracle import librosa_compat as lc
    rng = np.random.default_rng(5)
    D = (rng.standard_normal((513, 23)) + 1j * rng.standard_normal((513, 23))).astype(np.complex64)
    same = lc.phase_vocoder(D, 1.0)
    assert same.shape == D.shape and np.allclose(np.abs(same), np.abs(D), rtol=1e-6)
    for rate in (0.5, 1.3, 2.0):
        out = lc.phase_vocoder(D, rate)
        steps = np.arange(0, D.shape[1], rate)
        assert out.shape == (513, len(steps)) and out.dtype == np.complex64
        Dp = np.pad(np.abs(D), [(0, 0), (0, 2)])
        i0 = steps.astype(int)
        a = steps - i0
        assert np.allclose(np.abs(out), (1 - a) * Dp[:, i0] + a * Dp[:, i0 + 1], rtol=2e-6, atol=1e-6)
    sr, f = 22050, 22050 * 40 / 1024.0
    x = np.sin(2 * np.pi * f * np.arange(20000) / sr).astype(np.float32)
    S = lc.stft(x, 1024, 256, 1024)
    St = lc.phase_vocoder(S, 0.7)
    mid = St[:, 10:-10]
    assert np.all(np.argmax(np.abs(mid), axis=0) == 40)
    # the phase of the peak bin advances by 2 pi f hop / sr per output frame (that is what the vocoder keeps)
    dphi = np.angle(mid[40, 1:] / mid[40, :-1])
    expect = np.angle(np.exp(1j * 2 * np.pi * f * 256 / sr))
    assert np.abs(np.angle(np.exp(1j * (dphi - expect)))).max() < 1e-3


def test_mfcc_restatement_equals_scipy_ortho_dct():
    from scipy.fftpack import dct
    from oracle import librosa_compat as lc
    S = np.random.default_rng(8).standard_normal((80, 37))
    assert np.allclose(lc.mfcc(S, 13), dct(S, axis=0, type=2, norm='ortho')[:13], atol=1e-12)
