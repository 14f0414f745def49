"""Silence trimming used by ``load_audio`` (reference datasets/lj_speech.py:119 calls
``librosa.effects.trim(wav)`` with its defaults top_db=60, ref=np.max, frame_length=2048,
hop_length=512; wrapper at audio/effects.py:188-215).

SURVEY.md section 8f lists this step as the first "next" row (N1) after the STFT hot path.  The
batched feature path runs it on the device (``sstts_trim_bounds``, used by
``features_batch(..., trim=...)`` / ``DatasetHelper.features_from_wavs``); :func:`trim` below is
the single-clip host utility with the reference's call shape, :func:`trim_batch` the device one.
"""
import numpy as np


def _frame_mean_square(y, frame_length, hop_length):
    y = np.pad(np.asarray(y, dtype=np.float32), int(frame_length // 2), mode='reflect')
    n_frames = 1 + (len(y) - frame_length) // hop_length
    sq = np.concatenate(([0.0], np.cumsum(y.astype(np.float64) ** 2)))
    starts = np.arange(n_frames) * hop_length
    return (sq[starts + frame_length] - sq[starts]) / frame_length


def trim(y, top_db=60, frame_length=2048, hop_length=512):
    """Trim leading and trailing silence: frames whose mean-square energy is more than ``top_db``
    below the loudest frame.  Returns ``(y[start:end], np.array([start, end]))``."""
    y = np.asarray(y)
    if y.size == 0:
        return y, np.asarray([0, 0])
    mse = _frame_mean_square(y, frame_length, hop_length)
    ref = max(1e-10, float(mse.max()))
    db = 10.0 * np.log10(np.maximum(1e-10, mse)) - 10.0 * np.log10(ref)
    nonzero = np.flatnonzero(db > -top_db)
    if nonzero.size > 0:
        start = int(nonzero[0] * hop_length)
        end = min(y.shape[-1], int((nonzero[-1] + 1) * hop_length))
    else:
        start, end = 0, 0
    return y[start:end], np.asarray([start, end])


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
