"""dB statistics of ``datasets/statistics.py`` (reference :11-98) on the sm_100a feature kernel.

The analysis is the reference's hard-coded one: n_fft = 1024, hop = 256, win = 1024, 80 mel bands
up to ``sampling_rate // 2`` (datasets/statistics.py:31-51).  Per clip the kernel reduces the
linear and mel dB spectrograms to ``[min lin, max lin, min mel, max mel]`` on the device; the
corpus value is the float64 MEAN of those 4-vectors in listing order (:86-96), which is what the
reference prints as ``mel_mag_ref_db`` etc.
"""
import numpy as np

from .. import _runtime
from ..audio.io import load_wav, prefetch_batches

STATS_N_FFT = 1024
STATS_HOP = STATS_N_FFT // 4
STATS_WIN = STATS_N_FFT
STATS_N_MELS = 80


def decibel_statistics_batch(wavs, sampling_rate, precision='f64'):
    """(n_clips, 4) float64 ``[min lin dB, max lin dB, min mel dB, max mel dB]`` per clip."""
    res = _runtime.stft_features_batch(list(wavs), STATS_N_FFT, STATS_HOP, STATS_WIN,
                                       sampling_rate=sampling_rate, n_mels=STATS_N_MELS, fmin=0,
                                       fmax=sampling_rate // 2, want_minmax=True,
                                       precision=precision)
    return res.minmax


def decibel_statistics(wav, sampling_rate, precision='f64'):
    """reference datasets/statistics.py:11-66."""
    return decibel_statistics_batch([np.asarray(wav)], sampling_rate, precision=precision)[0]


def reduce_decibel_statistics(per_clip):
    """reference datasets/statistics.py:86-96: float64 running sum in listing order / n_files."""
    stats = np.zeros(4)
    for row in np.asarray(per_clip, dtype=np.float64):
        stats += row
    stats /= len(per_clip)
    return stats


def collect_decibel_statistics_from_wavs(wavs, sampling_rate, batch_clips=512, precision='f64'):
    """Corpus statistics from decoded clips, processed in device batches of ``batch_clips``."""
    wavs = list(wavs)
    rows = []
    for s in range(0, len(wavs), batch_clips):
        rows.append(decibel_statistics_batch(wavs[s:s + batch_clips], sampling_rate,
                                             precision=precision))
    return reduce_decibel_statistics(np.concatenate(rows, axis=0))


def collect_decibel_statistics(path_listing, batch_clips=512, precision='f64'):
    """reference datasets/statistics.py:69-98 -- average (min, max) dB over a list of wav files."""
    rows = []
    for _, loaded in prefetch_batches(path_listing, batch_clips, pcm16=True):   # threaded decode, one batch ahead
        wavs, sr = [], None
        for wav, sr_i in loaded:
            if sr is not None and sr_i != sr:
                # mixed sampling rates: flush what we have, the filterbank depends on sr
                rows.append(decibel_statistics_batch(wavs, sr, precision=precision))
                wavs = []
            sr = sr_i
            wavs.append(wav)
        if wavs:
            rows.append(decibel_statistics_batch(wavs, sr, precision=precision))
    return reduce_decibel_statistics(np.concatenate(rows, axis=0))


def reconstruction_errors_from_wavs(wavs, sampling_rate, n_iters, seed=None):
    """Per-clip Griffin-Lim reconstruction MSE (the loop body of reference
    datasets/statistics.py:156-181) for decoded clips, batched: |STFT| (n_fft 2048, 50 ms / 12.5 ms
    window) on the feature kernel, then ``n_iters`` Griffin-Lim iterations with the fused MSE of
    audio/synthesis.py:112.  ``seed`` keys the device phase generator (None: fresh random)."""
    from ..audio.conversion import ms_to_samples
    n_fft = 2048
    win = ms_to_samples(50.0, sampling_rate)
    hop = ms_to_samples(12.5, sampling_rate)
    wavs = [np.asarray(w) for w in wavs]
    res = _runtime.stft_features_batch(wavs, n_fft, hop, win, want_spec=True)
    mags = [np.abs(res.rows(res.spec, i)).T for i in range(res.n_clips)]
    _, mses = _runtime.griffin_lim_batch(mags, win, hop, n_fft, n_iters, seed=seed, return_mse=True)
    return mses


def collect_reconstruction_error(path_listing, n_iters, batch_clips=64, seed=None):
    """reference datasets/statistics.py:146-187: mean Griffin-Lim spectrogram MSE over a file list."""
    print("Collecting reconstruction statistics for {} files ...".format(len(path_listing)))
    mse_errors = []
    for s in range(0, len(path_listing), batch_clips):
        by_rate = {}
        for path in path_listing[s:s + batch_clips]:
            wav, sr = load_wav(path)
            by_rate.setdefault(sr, []).append(wav)
        for sr, wavs in by_rate.items():
            mse_errors.extend(reconstruction_errors_from_wavs(wavs, sr, n_iters, seed=seed))
    # a clip shorter than one hop has a single frame: the reference's griffin_lim_v2 raises for it
    # (empty re-analysis signal, audio/synthesis.py:96-106); the batched call reports None
    if any(m is None for m in mse_errors):
        raise ValueError('collect_reconstruction_error: a clip is shorter than one hop (single-frame '
                         'spectrogram), Griffin-Lim cannot re-analyse it')
    total_mse = sum(mse_errors) / len(mse_errors)
    print('Dataset MSE with {} iterations: {}'.format(n_iters, total_mse))
    return total_mse
