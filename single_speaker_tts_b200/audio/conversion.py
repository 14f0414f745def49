"""``audio/conversion.py`` of the reference: same names, arguments, dtypes and error behaviour.

These are host-side scalar/elementwise helpers on numpy arrays.  The batched GPU pipelines apply
the same formulas inside their kernels (``csrc/stft_kernels.cuh``: dB conversion and
normalisation in the STFT epilogue; :func:`..audio.synthesis.spectrograms_to_wavs` for the
inverse direction), so bulk data never takes this path.
"""
import numpy as np


def magnitude_to_decibel(mag):
    """``20 * log10(max(1e-5, mag))`` -- reference audio/conversion.py:5-29."""
    return 20.0 * np.log10(np.maximum(1e-5, mag))


def decibel_to_magnitude(mag_db):
    """``10 ** (mag_db / 20)`` -- reference audio/conversion.py:32-53 (AssertionError below
    -100 dB, same message)."""
    if (mag_db < -100.0).any():
        raise AssertionError('"conversion.decibel_to_magnitude" was asked to convert a dB value '
                             'smaller -100 dB.')
    return np.power(10.0, mag_db / 20.0)


def normalize_decibel(db, ref_db, max_db):
    """Map dB to [0, 1] -- reference audio/conversion.py:56-78."""
    return np.clip(1.0 + (db - ref_db) / (abs(ref_db) + abs(max_db)), 0.0, 1.0)


def inv_normalize_decibel(norm_db, ref_db, max_db):
    """Inverse of :func:`normalize_decibel` -- reference audio/conversion.py:81-102."""
    return ((np.clip(norm_db, 0.0, 1.0) - 1.0) * (abs(ref_db) + abs(max_db))) + ref_db


def samples_to_ms(samples, sampling_rate):
    """reference audio/conversion.py:105-119."""
    return (samples / sampling_rate) * 1000


def ms_to_samples(ms, sampling_rate):
    """Truncating conversion -- reference audio/conversion.py:122-136 (50 ms -> 1102, 12.5 ms ->
    275 at 22.05 kHz)."""
    return int((ms / 1000) * sampling_rate)


def get_duration(wav, sr):
    """``librosa.core.get_duration(y=wav, sr=sr)`` -- reference audio/conversion.py:139-153."""
    return float(np.asarray(wav).shape[-1]) / float(sr)
