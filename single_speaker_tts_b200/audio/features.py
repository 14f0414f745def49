"""``audio/features.py`` of the reference, executed by the sm_100a STFT feature kernel.

Signatures, argument order, defaults, output orientation ((bins, T)) and dtypes follow the
reference (audio/features.py:5-6,116); the transform runs in float64 on the device by default so
that every bin matches librosa's float64 ``stft`` (``precision='f32'`` selects the faster
float32 transform, whose bins more than ~100 dB below the frame peak differ).
"""
import numpy as np

from .. import _runtime


def linear_scale_spectrogram(wav, n_fft, hop_length=None, win_length=None, precision='f64'):
    """STFT -- reference audio/features.py:116-145.  (1 + n_fft/2, T) complex64, Fortran order."""
    wav = np.asarray(wav)
    res = _runtime.stft_features_batch([wav], n_fft, hop_length, win_length, want_spec=True,
                                       precision=precision)
    return res.spec.T


def mel_scale_spectrogram(wav, n_fft, sampling_rate, n_mels, fmin, fmax, hop_length, win_length,
                          power, precision='f64'):
    """Mel spectrogram ``mel_basis @ |STFT| ** power`` -- reference audio/features.py:5-86.
    (n_mels, T) float64."""
    wav = np.asarray(wav)
    res = _runtime.stft_features_batch([wav], n_fft, hop_length, win_length,
                                       sampling_rate=sampling_rate, n_mels=n_mels, fmin=fmin,
                                       fmax=fmax, want_mel_raw=True, power=power,
                                       precision=precision)
    return res.mel_raw.T


def calculate_mfccs(mel_spec, sampling_rate, n_mfcc):
    """Mel-frequency cepstral coefficients -- reference audio/features.py:89-113, i.e.
    ``librosa.feature.mfcc(S=mel_spec, sr=sampling_rate, n_mfcc=n_mfcc)``: the orthonormal DCT-II
    over the mel axis of a ``(n_mels, T)`` spectrogram.  Returns ``(n_mfcc, T)`` float64."""
    import ctypes
    import torch
    from .. import _hostio, _lib
    mel_spec = np.asarray(mel_spec)
    if mel_spec.ndim != 2:
        raise ValueError('mel_spec must have shape (n_mels, T)')
    n_mels, n_frames = mel_spec.shape
    lib = _lib.load()
    dev = _runtime.require_cuda()
    with torch.cuda.device(dev):
        mel_dev = _hostio.upload_rows([mel_spec.T], n_mels, torch.float64, dev, slot='mfcc')
        out_dev = torch.empty((n_frames, int(n_mfcc)), dtype=torch.float64, device=dev)
        _lib.check(lib.sstts_dct_project(ctypes.c_void_p(mel_dev.data_ptr()), n_frames, n_mels, int(n_mfcc),
                                         ctypes.c_void_p(out_dev.data_ptr()),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        out = _hostio.download(out_dev)
        torch.cuda.current_stream().synchronize()
    return out.T


def features_batch(wavs, n_fft, hop_length, win_length, sampling_rate, n_mels, fmin, fmax,
                   linear_ref_db, linear_mag_max_db, mel_mag_ref_db, mel_mag_max_db, reduction=1,
                   precision='f64', trim=None):
    """Batched ``load_audio`` core (reference datasets/lj_speech.py:124-156 after decode/trim):
    one STFT per clip, fused |.| -> dB -> normalise for the linear and the mel spectrogram,
    reduction padding and folding.  Returns a list of ``(mel, lin)`` float32 pairs shaped
    ``(ceil(T/r), r * n_mels)`` and ``(ceil(T/r), r * (1 + n_fft/2))``.
    ``trim=(top_db, frame_length, hop_length)`` runs the ``librosa.effects.trim`` step of
    datasets/lj_speech.py:119 on the device first."""
    parts = _runtime.stft_features_parts(list(wavs), n_fft, hop_length, win_length,
                                         sampling_rate=sampling_rate, n_mels=n_mels, fmin=fmin,
                                         fmax=fmax, reduction=reduction, want_lin=True, want_mel=True,
                                         normalize=(linear_ref_db, linear_mag_max_db, mel_mag_ref_db,
                                                    mel_mag_max_db),
                                         precision=precision, trim=trim)
    out = []
    n_bins = 1 + n_fft // 2
    for _, _, res in parts:
        for i in range(res.n_clips):
            mel = res.rows(res.mel_db, i, padded=True).reshape((-1, n_mels * reduction))
            lin = res.rows(res.lin_db, i, padded=True).reshape((-1, n_bins * reduction))
            out.append((mel, lin))
    return out
