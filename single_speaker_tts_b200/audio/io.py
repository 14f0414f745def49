"""``audio/io.py`` of the reference: wav decode / encode on the host.

File decode/encode is host I/O, not part of the device hot path.
The reference goes through ``librosa.load(sr=None)`` / ``librosa.output.write_wav``
(audio/io.py:30,53); here PCM wav files are read and written with ``scipy.io.wavfile`` with the
same conventions (float32 in [-1, 1), mono mix-down, native sampling rate).
"""
import numpy as np
from scipy.io import wavfile


def load_wav(wav_path, offset=0.0, duration=None):
    """reference audio/io.py:5-30 -- returns ``(float32 mono samples, sampling_rate)``."""
    sr, data = wavfile.read(wav_path)
    if data.dtype == np.int16:
        wav = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        wav = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif data.dtype == np.uint8:
        wav = (data.astype(np.float32) - 128.0) / 128.0
    else:
        wav = data.astype(np.float32)
    if wav.ndim > 1:
        wav = wav.mean(axis=1).astype(np.float32)
    start = int(round(offset * sr))
    end = None if duration is None else start + int(round(duration * sr))
    return np.ascontiguousarray(wav[start:end]), sr


def save_wav(wav_path, wav, sampling_rate, norm=False):
    """reference audio/io.py:33-53 -- float wav writer, optional peak normalisation."""
    wav = np.asarray(wav)
    if norm and wav.size and np.max(np.abs(wav)) > 0:
        wav = wav / np.max(np.abs(wav))
    wavfile.write(wav_path, int(sampling_rate), wav.astype(np.float32))
