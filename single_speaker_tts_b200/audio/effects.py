"""Silence trimming used by ``load_audio`` (reference datasets/lj_speech.py:119 calls
``librosa.effects.trim(wav)`` with its defaults top_db=60, ref=np.max, frame_length=2048,
hop_length=512; wrapper at audio/effects.py:188-215).

SURVEY.md section 8f lists this step as the first "next" row (N1) after the STFT hot path; it
runs on the host here: one pass of frame energies per clip, O(N), before the clip is packed for
the device.
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
