"""The real kernel sources (csrc/*.cuh) on the CPU SIMT emulator vs the oracle.

No GPU is needed: tests/emu compiles stft_kernels.cuh / fft_warp.cuh / host_plan.h with g++ and
runs thread blocks on OS threads (test infrastructure, never part of the product).  These tests
pin the kernels' index maths -- warp FFT slot order, conjugate-pair shuffles, tile / edge /
reflect logic, gather overlap-add -- before any GPU time is spent; the GPU tier repeats the same
comparisons through the C ABI.
"""
import numpy as np
import pytest

from oracle import librosa_compat as lc
from oracle import reference_audio as ra
from single_speaker_tts_b200.synthetic import speech_like_clip

WIN, HOP, NFFT = 1102, 275, 2048


@pytest.mark.parametrize('prec,tol', [(0, 5e-7), (1, 1e-14)])
@pytest.mark.parametrize('inverse', [False, True])
def test_warp_fft1024(emu, prec, tol, inverse):
    rng = np.random.default_rng(0)
    z = rng.normal(size=1024) + 1j * rng.normal(size=1024)
    got = emu.fft1024(z, inverse=inverse, prec=prec)
    ref = np.fft.ifft(z) * 1024 if inverse else np.fft.fft(z)
    assert np.abs(got - ref).max() / np.abs(ref).max() < tol


def _gl_case(frames, seed=5):
    rng = np.random.default_rng(seed)
    mags, angs = [], []
    for i, T in enumerate(frames):
        x = speech_like_clip(HOP * (T - 1) + 5, rng)
        m = np.abs(lc.stft(x, NFFT, HOP, WIN))
        assert m.shape[1] == T
        mags.append(m)
        angs.append(np.exp(2j * np.pi * np.random.RandomState(i).rand(*m.shape)))
    return mags, angs


@pytest.mark.parametrize('prec,tol', [(1, 5e-7), (0, 5e-6)])
def test_griffin_lim_tiles_and_edges(emu, prec, tol):
    """T = 1 (empty output), 2 (multi-bounce reflect), single tiles, merged 9-frame tile, 8+5 split,
    many tiles; 2 iterations exercise synth, fused step, finalize and the parity buffers."""
    frames = [1, 2, 5, 9, 13, 30]
    mags, angs = _gl_case(frames)
    wavs, mses = emu.griffin_lim(mags, angs, 2, prec=prec, want_mse=True)
    for T, m, a, w, mse in zip(frames, mags, angs, wavs, mses):
        assert w.shape == (HOP * (T - 1),)
        if T == 1:
            continue
        ref, rmse = ra.griffin_lim_v2(m, WIN, HOP, NFFT, 2, angles=a, batched_fft=True)
        assert not np.isnan(w).any()
        assert np.linalg.norm(w - ref) / np.linalg.norm(ref) < tol
        assert abs(mse - rmse) / rmse < 1e-5


def test_griffin_lim_bulk_staging_variant(emu, monkeypatch):
    """The iteration kernel's second staging variant (cp.async.bulk + mbarrier for interior tiles,
    SSTTS_GL_STAGING=bulk): same results as the default staging up to the rounding of the folded
    window x normalisation table, on a case with interior, first, last and reflecting tiles."""
    frames = [2, 9, 30, 45]
    mags, angs = _gl_case(frames)
    monkeypatch.delenv('SSTTS_GL_STAGING', raising=False)
    ref = emu.griffin_lim(mags, angs, 3, prec=0)
    monkeypatch.setenv('SSTTS_GL_STAGING', 'bulk')
    got = emu.griffin_lim(mags, angs, 3, prec=0)
    for T, m, a, w, r in zip(frames, mags, angs, got, ref):
        assert not np.isnan(w).any()
        assert np.linalg.norm(w - r) / np.linalg.norm(r) < 2e-6
        orc = ra.spectrogram_to_wav(m, WIN, HOP, NFFT, 3, angles=a, batched_fft=True)
        assert np.linalg.norm(w - orc) / np.linalg.norm(orc) < 5e-6


def test_griffin_lim_phase_drawn_inside_the_synthesis_launch(emu):
    """angles=None: the first launch draws unit phasors from the counter-based generator (seed fixed
    in the emulator driver).  Deterministic, finite, and a valid Griffin-Lim run: the spectrogram
    error after a few iterations is far below the random-phase starting point."""
    mags, _ = _gl_case([9, 30])
    a = emu.griffin_lim(mags, None, 3, prec=0, want_mse=True)
    b = emu.griffin_lim(mags, None, 3, prec=0, want_mse=True)
    for (w0, w1, m, mse) in zip(a[0], b[0], mags, a[1]):
        assert np.array_equal(w0, w1) and np.isfinite(w0).all() and w0.std() > 0
        assert mse < 0.5 * np.mean(m ** 2)


def test_griffin_lim_zero_iterations_is_istft(emu):
    mags, angs = _gl_case([12])
    w = emu.griffin_lim(mags, angs, 0, prec=1)[0]
    ref = lc.istft(mags[0].astype(np.complex128) * angs[0], HOP, WIN)
    assert np.abs(w - ref).max() < 1e-6 * np.abs(ref).max() + 1e-7


def test_griffin_lim_zero_bins_give_unit_phase(emu):
    """|E| == 0 must yield phase (1, 0) (np.exp(1j*np.angle(0)) == 1, audio/synthesis.py:109)."""
    T = 6
    mag = np.zeros((1025, T), np.float32)
    mag[40, :] = 1.0
    ang = np.exp(2j * np.pi * np.random.RandomState(3).rand(1025, T))
    w = emu.griffin_lim([mag], [ang], 2, prec=0)[0]
    ref = ra.spectrogram_to_wav(mag, WIN, HOP, NFFT, 2, angles=ang)
    assert not np.isnan(w).any()
    assert np.linalg.norm(w - ref) / np.linalg.norm(ref) < 1e-4


def test_griffin_lim_dynamic_geometry(emu):
    """A window / hop other than the model's goes through the run-time geometry kernels."""
    win, hop = 1024, 256
    x = speech_like_clip(hop * 19 + 3, np.random.default_rng(8))
    m = np.abs(lc.stft(x, NFFT, hop, win))
    a = np.exp(2j * np.pi * np.random.RandomState(1).rand(*m.shape))
    w = emu.griffin_lim([m], [a], 2, prec=1, win=win, hop=hop)[0]
    ref = ra.spectrogram_to_wav(m, win, hop, NFFT, 2, angles=a)
    assert np.linalg.norm(w - ref) / np.linalg.norm(ref) < 5e-7


@pytest.mark.parametrize('win,hop', [(2048, 512), (2046, 512), (1500, 375)])
def test_griffin_lim_frame_shift_boundaries(emu, win, hop):
    """The iteration kernel transforms frames at odd span offsets one sample later in their zero-padded buffer
    (aligned sample-pair loads).  win = n_fft leaves no room (the kernel must fall back to scalar loads),
    win = n_fft - 2 is the last window length with room for the shift, 1500 / 375 alternates shifted and
    unshifted frames; all must match the oracle in both precisions."""
    rng = np.random.default_rng(21)
    for T in (2, 9, 19):
        x = speech_like_clip(hop * (T - 1) + 3, rng)
        m = np.abs(lc.stft(x, NFFT, hop, win))
        a = np.exp(2j * np.pi * np.random.RandomState(T).rand(*m.shape))
        for prec, tol in ((1, 1e-6), (0, 1e-5)):
            (w,), (mse,) = emu.griffin_lim([m], [a], 2, prec=prec, win=win, hop=hop, want_mse=True)
            ref, rmse = ra.griffin_lim_v2(m, win, hop, NFFT, 2, angles=a, batched_fft=True)
            assert w.shape == ref.shape and not np.isnan(w).any()
            assert np.linalg.norm(w - ref) / np.linalg.norm(ref) < tol, (win, hop, T, prec)
            assert abs(mse - rmse) / rmse < 1e-5


@pytest.mark.parametrize('n_fft,win,hop,native', [(1024, 1024, 256, '1'), (1024, 1024, 256, '0'), (1024, 800, 200, '1'),
                                                  (512, 400, 100, '1')])
def test_griffin_lim_shorter_transforms(emu, monkeypatch, n_fft, win, hop, native):
    """n_fft 1024 (audio/effects.py:71-86 calls spectrogram_to_wav with 1024 / 256 / 1024): the native path
    (512-point complex transform per half-warp, two frames per warp, 16-frame tiles; odd tiles leave the last
    half-warp without a frame) and, with SSTTS_GL_NATIVE1024=0, the embedding in the 2048-point transform
    (every (2048 / n_fft)-th bin, first period of the inverse) that n_fft 512 always uses."""
    monkeypatch.setenv('SSTTS_GL_NATIVE1024', native)
    rng = np.random.default_rng(12)
    for T in (2, 3, 21, 40):
        x = speech_like_clip(hop * (T - 1) + 7, rng)
        m = np.abs(lc.stft(x, n_fft, hop, win))
        assert m.shape == (1 + n_fft // 2, T)
        a = np.exp(2j * np.pi * np.random.RandomState(T).rand(*m.shape))
        for prec, tol in ((1, 1e-6), (0, 1e-5)):
            (w,), (mse,) = emu.griffin_lim([m], [a], 2, prec=prec, win=win, hop=hop, n_fft=n_fft, want_mse=True)
            ref, rmse = ra.griffin_lim_v2(m, win, hop, n_fft, 2, angles=a, batched_fft=True)
            assert w.shape == ref.shape and not np.isnan(w).any()
            assert np.linalg.norm(w - ref) / np.linalg.norm(ref) < tol
            assert abs(mse - rmse) / rmse < 1e-5


def test_griffin_lim_native_1024_ragged_batch_and_seeded_phase(emu, monkeypatch):
    """The native n_fft 1024 kernels on a ragged batch (single-frame, two-frame, odd and multi-tile
    utterances share one launch sequence) and with the phase drawn inside the synthesis launch."""
    monkeypatch.delenv('SSTTS_GL_NATIVE1024', raising=False)
    rng = np.random.default_rng(5)
    frames = [1, 2, 7, 16, 17, 33]
    mags, angs = [], []
    for i, T in enumerate(frames):
        x = speech_like_clip(256 * (T - 1) + 5, rng)
        m = np.abs(lc.stft(x, 1024, 256, 1024))
        mags.append(m)
        angs.append(np.exp(2j * np.pi * np.random.RandomState(i).rand(*m.shape)))
    wavs, mses = emu.griffin_lim(mags, angs, 3, prec=0, win=1024, hop=256, n_fft=1024, want_mse=True)
    for T, m, a, w, mse in zip(frames, mags, angs, wavs, mses):
        assert w.shape == (256 * (T - 1),)
        if T == 1:
            continue
        ref, rmse = ra.griffin_lim_v2(m, 1024, 256, 1024, 3, angles=a, batched_fft=True)
        assert np.linalg.norm(w - ref) / np.linalg.norm(ref) < 1e-5 and abs(mse - rmse) / rmse < 1e-5
    a = emu.griffin_lim(mags[2:], None, 3, prec=0, win=1024, hop=256, n_fft=1024, want_mse=True)
    b = emu.griffin_lim(mags[2:], None, 3, prec=0, win=1024, hop=256, n_fft=1024, want_mse=True)
    for w0, w1, m, mse in zip(a[0], b[0], mags[2:], a[1]):
        assert np.array_equal(w0, w1) and np.isfinite(w0).all() and w0.std() > 0
        assert mse < 0.5 * np.mean(m ** 2)


@pytest.mark.parametrize('prec,tol_lin', [(1, 5e-7), (0, 2e-3)])
def test_feature_pipeline(emu, prec, tol_lin):
    """load_audio core (datasets/lj_speech.py:124-156): ragged clips incl. N = 1, N < hop, N == hop."""
    rng = np.random.default_rng(7)
    lens = [1, 2, 274, 275, 1500, 5000]
    wavs = [speech_like_clip(max(n, 8), rng)[:n] for n in lens]
    res = emu.stft_features(wavs, prec=prec, r=5, normalize=(35.66, 100.0, 6.02, 99.89))
    for w, r in zip(wavs, res):
        S = lc.stft(w, NFFT, HOP, WIN).T
        assert S.shape[0] == r['T']
        mel_ref, lin_ref = ra.load_audio_from_wav(w, 22050, trim=False)
        assert r['lin'].shape == (mel_ref.shape[0] * 5, 1025)
        assert np.abs(r['lin'] - lin_ref.reshape(-1, 1025)).max() < tol_lin
        assert np.abs(r['mel'] - mel_ref.reshape(-1, 80)).max() < 2e-6
        assert np.abs(r['spec'][:r['T']] - S).max() / np.abs(S).max() < (1e-7 if prec else 1e-6)
        mr = ra.mel_scale_spectrogram(w, NFFT, 22050, 80, 0, 8000, HOP, WIN, 1).T
        assert np.abs(r['melraw'][:r['T']] - mr).max() / np.abs(mr).max() < 1e-6


@pytest.mark.parametrize('prec,tol_lin', [(1, 5e-7), (0, 2e-3)])
@pytest.mark.parametrize('normalize', [(35.66, 100.0, 6.02, 99.89), None])
def test_feature_pipeline_fused_db_mode(emu, prec, tol_lin, normalize):
    """The fused dB-feature kernel mode (lin + mel dB only: the pre-calculation request) against the
    oracle and against the generic mode; clips long enough for interior tiles (16-byte staging with
    every misalignment) next to the short / reflecting ones."""
    rng = np.random.default_rng(17)
    lens = [1, 274, 275, 1500, 5003, 9001, 12345]
    wavs = [speech_like_clip(max(n, 8), rng)[:n] for n in lens]
    fast = emu.stft_features(wavs, prec=prec, r=5, normalize=normalize, fast=True)
    slow = emu.stft_features(wavs, prec=prec, r=5, normalize=normalize)
    for w, f, g in zip(wavs, fast, slow):
        assert not np.isnan(f['lin']).any() and not np.isnan(f['mel']).any()
        scale = 1.0 if normalize is not None else 135.0      # dB instead of (0, 1)
        assert np.abs(f['lin'] - g['lin']).max() < 2e-6 * scale
        assert np.abs(f['mel'] - g['mel']).max() < 2e-6 * scale
        if normalize is not None:
            mel_ref, lin_ref = ra.load_audio_from_wav(w, 22050, trim=False)
            assert np.abs(f['lin'] - lin_ref.reshape(-1, 1025)).max() < tol_lin
            assert np.abs(f['mel'] - mel_ref.reshape(-1, 80)).max() < 3e-6


@pytest.mark.parametrize('prec,tol', [(1, 1e-7), (0, 1e-6)])
def test_decibel_statistics_geometry(emu, prec, tol):
    """datasets/statistics.py:31-51: n_fft 1024 / hop 256 / win 1024, fmax sr//2 -- the feature kernel's
    NATIVE n_fft 1024 path (512-point complex transform on half a warp, two frames per warp, 16-frame
    tiles): single-frame clips, odd and even tile sizes, several tiles per clip."""
    rng = np.random.default_rng(9)
    wavs = [speech_like_clip(max(n, 8), rng)[:n] for n in (1, 255, 300, 256 * 2, 5000, 256 * 16 + 5, 22050)]
    res = emu.stft_features(wavs, prec=prec, n_fft=1024, win=1024, hop=256, fmax=11025.)
    for w, r in zip(wavs, res):
        S = lc.stft(w, 1024, 256, 1024).T
        assert r['spec'].shape == S.shape
        assert np.abs(r['spec'] - S).max() / np.abs(S).max() < tol
        if prec == 1:
            assert np.abs(r['minmax'] - ra.decibel_statistics(w, 22050)).max() < 2e-5
            mr = ra.mel_scale_spectrogram(w, 1024, 22050, 80, 0, 11025, 256, 1024, 1).T
            assert np.abs(r['melraw'] - mr).max() / np.abs(mr).max() < 1e-6
            lin_db = ra.magnitude_to_decibel(np.abs(S))
            assert np.abs(r['lin'] - lin_db).max() < 2e-4


@pytest.mark.parametrize('prec', [1, 0])
def test_statistics_fused_mode(emu, prec):
    """The statistics request (per-clip extrema, optionally linear / mel dB) runs the native path's fused
    float32 epilogue: extrema of the dB values taken as dB of the extrema, mel sums in float32."""
    rng = np.random.default_rng(29)
    wavs = [speech_like_clip(max(n, 8), rng)[:n] for n in (1, 300, 256 * 3, 5000, 256 * 16 + 5, 22050)]
    wavs.append(np.zeros(3000, np.float32))                          # digital silence: the -100 dB floor
    fast = emu.stft_features(wavs, prec=prec, n_fft=1024, win=1024, hop=256, fmax=11025., fast=True)
    slow = emu.stft_features(wavs, prec=prec, n_fft=1024, win=1024, hop=256, fmax=11025.)
    for w, f, g in zip(wavs, fast, slow):
        assert np.abs(f['minmax'] - g['minmax']).max() < 1e-4        # dB
        assert np.abs(f['lin'] - g['lin']).max() < 3e-4 and np.abs(f['mel'] - g['mel']).max() < 3e-4
        if prec == 1:
            assert np.abs(f['minmax'] - ra.decibel_statistics(w, 22050)).max() < 1e-4


def test_native_1024_transform_with_a_shorter_window(emu):
    """n_fft 1024 with win < n_fft (window centred with zero padding) and a hop that is not win / 4:
    the run-time geometry instance of the native path; normalised dB outputs and reduction padding."""
    rng = np.random.default_rng(19)
    wavs = [speech_like_clip(n, rng) for n in (777, 6000, 15000)]
    consts = (35.66, 100.0, 6.02, 99.89)
    res = emu.stft_features(wavs, prec=1, r=5, n_fft=1024, win=800, hop=200, fmax=8000., normalize=consts)
    for w, r in zip(wavs, res):
        S = lc.stft(w, 1024, 200, 800).T
        T = S.shape[0]
        assert r['T'] == T and np.abs(r['spec'][:T] - S).max() / np.abs(S).max() < 1e-7
        lin = ra.normalize_decibel(ra.magnitude_to_decibel(np.abs(S)), consts[0], consts[1])
        mel = ra.mel_scale_spectrogram(w, 1024, 22050, 80, 0, 8000, 200, 800, 1).T
        mel = ra.normalize_decibel(ra.magnitude_to_decibel(np.abs(mel)), consts[2], consts[3])
        assert np.abs(r['lin'][:T] - lin).max() < 2e-6 and np.abs(r['mel'][:T] - mel).max() < 2e-6
        assert np.all(r['lin'][T:] == 0) and np.all(r['mel'][T:] == 0)


def test_mel_basis_matches_oracle(emu):
    for n_fft, fmax in ((2048, 8000), (1024, 11025)):
        mb = emu.mel_basis(22050, n_fft, 80, 0, fmax)
        ref = lc.mel_filterbank(22050, n_fft, 80, 0, fmax)
        assert np.abs(mb - ref).max() < 1e-15 and ((mb != 0) == (ref != 0)).all()


def test_trim_bounds_match_librosa_trim(emu):
    """librosa.effects.trim (datasets/lj_speech.py:119) as a device kernel: same (start, end)."""
    rng = np.random.default_rng(3)
    sp = speech_like_clip(12000, rng)
    wavs = [np.concatenate([np.zeros(4000, np.float32), sp, np.zeros(5000, np.float32)]),
            sp[:3000].copy(), np.zeros(6000, np.float32),
            np.concatenate([1e-5 * rng.normal(size=3000).astype(np.float32), sp[:7000], 1e-6 * np.ones(2500, np.float32)])]
    got = emu.trim_bounds(wavs)
    for w, b in zip(wavs, got):
        assert tuple(b) == tuple(lc.trim(w)[1])


def test_address_sanitizer_run_of_the_emulated_kernels(tmp_path):
    """compute-sanitizer is not available on the GPU pool, so the bounds check of the kernels is an
    AddressSanitizer build of the emulator driver (tests/emu/asan_main.cpp): every global buffer and
    every CTA's shared memory is an exactly-sized heap block."""
    import os
    import shutil
    import subprocess
    import __graft_entry__ as ge
    if os.environ.get('SSTTS_SKIP_ASAN') or shutil.which('g++') is None:
        pytest.skip('g++ / ASAN run disabled')
    exe = ge.build_asan()          # cached in tests/emu/ like the emulator library (rebuilt when a source changes)
    if exe is None:
        pytest.skip('no libasan in this toolchain')
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and 'asan run ok' in run.stdout, (run.stdout + run.stderr)[-3000:]


def test_random_geometries_griffin_lim(emu):
    """Seeded sweep over run-time geometries (n_fft 512 / 1024 / 2048, even n_fft - win, win / 5 <= hop <= win,
    ragged batches with 1- and 2-frame utterances, 0-2 iterations) against the oracle in both precisions --
    the bounded form of the fuzz run that accompanied the late kernel changes (675 cases, no logic error)."""
    rng = np.random.default_rng(2024)
    for _ in range(14):
        n_fft = int(rng.choice([512, 1024, 2048]))
        win = min(int(rng.integers(n_fft // 4, n_fft // 2 + 1)) * 2, n_fft)
        hop = int(rng.integers(-(-win // 5), win + 1))
        frames = [int(v) for v in rng.integers(1, 30, size=int(rng.integers(1, 4)))]
        mags, angs = [], []
        for T in frames:
            x = speech_like_clip(hop * (T - 1) + int(rng.integers(0, hop)), rng)
            m = np.abs(lc.stft(x, n_fft, hop, win))
            mags.append(m)
            angs.append(np.exp(2j * np.pi * rng.random(m.shape)))
        it = int(rng.integers(0, 3))
        for prec, tol in ((1, 5e-6), (0, 1e-4)):       # 2-frame utterances at hop > win / 2 are ill-conditioned
            wavs = emu.griffin_lim(mags, angs, it, prec=prec, win=win, hop=hop, n_fft=n_fft)
            for m, a, w in zip(mags, angs, wavs):
                if m.shape[1] == 1:
                    assert w.shape == (0,)
                    continue
                ref = ra.spectrogram_to_wav(m, win, hop, n_fft, it, angles=a, batched_fft=True)
                assert w.shape == ref.shape
                assert np.linalg.norm(w - ref) / np.linalg.norm(ref) < tol, (n_fft, win, hop, m.shape[1], it, prec)


def test_random_geometries_features(emu):
    """Seeded sweep of the feature kernels over run-time geometries and ragged clip lengths (incl. hop > win and
    clips shorter than a hop): complex STFT and raw mel against the oracle, zero pad rows (1,336 cases in the
    unbounded run, none failing)."""
    rng = np.random.default_rng(4048)
    for _ in range(10):
        n_fft = int(rng.choice([512, 1024, 2048]))
        win = int(rng.integers(n_fft // 4, n_fft // 2 + 1)) * 2
        hop = int(rng.integers(max(8, win // 8), win + 50))
        r = int(rng.choice([1, 5]))
        lens = [int(v) for v in rng.integers(1, 6000, size=int(rng.integers(1, 5)))]
        wavs = [speech_like_clip(max(k, 8), rng)[:k] for k in lens]
        fmax = float(rng.choice([8000., 11025.]))
        for prec in (1, 0):
            res = emu.stft_features(wavs, prec=prec, r=r, n_fft=n_fft, win=win, hop=hop, fmax=fmax)
            for w, o in zip(wavs, res):
                S = lc.stft(w, n_fft, hop, win).T
                T = o['T']
                assert S.shape[0] == T
                assert np.abs(o['spec'][:T] - S).max() / max(np.abs(S).max(), 1e-30) < (2e-7 if prec else 3e-6)
                mr = ra.mel_scale_spectrogram(w, n_fft, 22050, 80, 0, fmax, hop, win, 1).T
                assert np.abs(o['melraw'][:T] - mr).max() / max(np.abs(mr).max(), 1e-30) < (1e-6 if prec else 1e-5)
                assert (o['lin'][T:] == 0).all()
