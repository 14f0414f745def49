"""Silence trimming used by ``load_audio`` (reference datasets/lj_speech.py:119 calls
``librosa.effects.trim(wav)`` with its defaults top_db=60, ref=np.max, frame_length=2048,
hop_length=512; wrapper at audio/effects.py:188-215).

SURVEY.md section 8f lists this step as the first "next" row (N1) after the STFT hot path.  The
batched feature path runs it on the device (``sstts_trim_bounds``, used by
``features_batch(..., trim=...)`` / ``DatasetHelper.features_from_wavs``); :func:`trim` below is
the single-clip call with the reference's call shape, :func:`trim_batch` the batched one -- both
use the same device kernel (no host implementation: the package has no CPU compute path).
"""
import numpy as np


def trim(y, top_db=60, frame_length=2048, hop_length=512):
    """Trim leading and trailing silence (``librosa.effects.trim`` with the reference's call shape,
    audio/effects.py:188-215): frames whose mean-square energy is more than ``top_db`` below the loudest
    frame.  Returns ``(y[start:end], np.array([start, end]))``.  Runs on the device like everything else
    in this package (:func:`trim_batch` with one clip); there is no host implementation."""
    y = np.asarray(y)
    if y.size == 0:
        return y, np.asarray([0, 0])
    trimmed, bounds = trim_batch([y], top_db=top_db, frame_length=frame_length, hop_length=hop_length)
    return trimmed[0], np.asarray([int(bounds[0][0]), int(bounds[0][1])])


def trim_batch(wavs, top_db=60, frame_length=2048, hop_length=512):
    """Device version for a list of clips: returns ``(list of trimmed views, (n, 2) int64 bounds)``."""
    import ctypes
    import torch
    from .. import _hostio, _lib, _runtime
    lib = _lib.load()
    dev = _runtime.require_cuda()
    n = len(wavs)
    lens = np.asarray([len(w) for w in wavs], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    with torch.cuda.device(dev):
        wav_dev = _hostio.upload_flat(wavs, torch.float32, dev, slot='wav')
        bounds_dev = torch.empty((n, 2), dtype=torch.int64, device=dev)
        start_dev = torch.from_numpy(starts).to(dev)      # keep alive until the kernel has run
        len_dev = torch.from_numpy(lens).to(dev)
        _lib.check(lib.sstts_trim_bounds(ctypes.c_void_p(wav_dev.data_ptr()), n,
                                         ctypes.c_void_p(start_dev.data_ptr()),
                                         ctypes.c_void_p(len_dev.data_ptr()),
                                         float(top_db), int(frame_length), int(hop_length),
                                         ctypes.c_void_p(bounds_dev.data_ptr()),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        bounds = bounds_dev.cpu().numpy()
    return [w[b[0]:b[1]] for w, b in zip(wavs, bounds)], bounds


def time_stretch(wav, rate):
    """reference audio/effects.py:46-86 -- time-stretch by ``rate`` (> 1: faster): STFT (n_fft = win =
    1024, hop 256) -> phase vocoder -> magnitude -> 25 Griffin-Lim iterations (random initial phase
    from numpy's global RNG, like ``spectrogram_to_wav``).  The reference keeps only the MAGNITUDE of
    the stretched spectrogram (:80), i.e. the linear interpolation of |STFT| at the fractional frame
    positions; STFT, interpolation (``sstts_stretch_magnitude``) and Griffin-Lim all run on the
    device.  Returns float32 audio of about ``len(wav) / rate`` samples."""
    import ctypes
    import torch
    from .. import _hostio, _lib, _runtime
    from .synthesis import spectrogram_to_wav
    if rate <= 0.0:
        raise ValueError('The fixed rate used to stretch the signal must be greater 0.')
    n_fft = 1024
    win_len = n_fft
    hop_len = win_len // 4
    reconstr_iters = 25
    lib = _lib.load()
    dev = _runtime.require_cuda()
    n_bins = 1 + n_fft // 2
    with torch.cuda.device(dev):
        res = _runtime.stft_features_batch([np.asarray(wav)], n_fft, hop_len, win_len, want_spec=True,
                                           keep_on_device=True)
        n_frames = res.frames[0]
        n_out = int(lib.sstts_stretch_frames(n_frames, float(rate)))
        spec = torch.view_as_real(res.spec[:n_frames].contiguous())
        mag_dev = torch.empty((n_out, n_bins), dtype=torch.float32, device=dev)
        _lib.check(lib.sstts_stretch_magnitude(ctypes.c_void_p(spec.data_ptr()), n_frames, n_bins, float(rate),
                                               ctypes.c_void_p(mag_dev.data_ptr()),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        mag = _hostio.download(mag_dev)
        torch.cuda.current_stream().synchronize()
    return spectrogram_to_wav(mag.T, win_len, hop_len, n_fft, reconstr_iters)


def stretch_magnitude(stft, rate):
    """``np.abs(librosa.core.phase_vocoder(stft, rate))`` for a host ``(bins, T)`` complex64 STFT
    (device kernel; the building block of :func:`time_stretch`)."""
    import ctypes
    import torch
    from .. import _hostio, _lib, _runtime
    lib = _lib.load()
    dev = _runtime.require_cuda()
    stft = np.asarray(stft)
    n_bins, n_frames = stft.shape
    with torch.cuda.device(dev):
        spec = torch.from_numpy(np.ascontiguousarray(stft.T.astype(np.complex64)).view(np.float32)).to(dev)
        n_out = int(lib.sstts_stretch_frames(n_frames, float(rate)))
        mag_dev = torch.empty((n_out, n_bins), dtype=torch.float32, device=dev)
        _lib.check(lib.sstts_stretch_magnitude(ctypes.c_void_p(spec.data_ptr()), n_frames, n_bins, float(rate),
                                               ctypes.c_void_p(mag_dev.data_ptr()),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        mag = _hostio.download(mag_dev)
        torch.cuda.current_stream().synchronize()
    return mag.T
