"""CPU restatement of the reference's own functions on the audio hot path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function names the reference
file:line it follows (paths relative to ``/root/reference``).  The librosa calls the reference
makes are served by ``oracle.librosa_compat``.
"""
import numpy as np

from . import librosa_compat as lc


# ----------------------------------------------------------------------------------------------
# audio/conversion.py
# ----------------------------------------------------------------------------------------------
def magnitude_to_decibel(mag):
    """audio/conversion.py:5-29 -- ``20 log10(max(1e-5, mag))``, dtype of ``mag`` preserved."""
    return 20.0 * np.log10(np.maximum(1e-5, mag))


def decibel_to_magnitude(mag_db):
    """audio/conversion.py:32-53 -- ``10 ** (dB / 20)``; AssertionError below -100 dB."""
    if (mag_db < -100.0).any():
        raise AssertionError('"conversion.decibel_to_magnitude" was asked to convert a dB value '
                             'smaller -100 dB.')
    return np.power(10.0, mag_db / 20.0)


def normalize_decibel(db, ref_db, max_db):
    """audio/conversion.py:56-78 -- ``clip(1 + (db - ref) / (|ref| + |max|), 0, 1)``."""
    return np.clip(1.0 + (db - ref_db) / (abs(ref_db) + abs(max_db)), 0.0, 1.0)


def inv_normalize_decibel(norm_db, ref_db, max_db):
    """audio/conversion.py:81-102 -- inverse of ``normalize_decibel`` on the clipped input."""
    return ((np.clip(norm_db, 0.0, 1.0) - 1.0) * (abs(ref_db) + abs(max_db))) + ref_db


def samples_to_ms(samples, sampling_rate):
    """audio/conversion.py:105-119."""
    return (samples / sampling_rate) * 1000


def ms_to_samples(ms, sampling_rate):
    """audio/conversion.py:122-136 -- truncating ``int((ms / 1000) * sr)``."""
    return int((ms / 1000) * sampling_rate)


# ----------------------------------------------------------------------------------------------
# audio/features.py
# ----------------------------------------------------------------------------------------------
def linear_scale_spectrogram(wav, n_fft, hop_length=None, win_length=None):
    """audio/features.py:116-145 -- ``librosa.stft`` wrapper, (1 + n_fft/2, T) complex64."""
    return lc.stft(wav, n_fft=n_fft, hop_length=hop_length, win_length=win_length)


def mel_scale_spectrogram(wav, n_fft, sampling_rate, n_mels, fmin, fmax, hop_length, win_length,
                          power):
    """audio/features.py:5-86 -- second STFT, ``|.| ** power``, HTK/Slaney filterbank rebuilt per
    call (:75-80), dense ``np.dot`` (:84).  Returns (n_mels, T) float64."""
    mag_phase_spec = lc.stft(wav, n_fft=n_fft, hop_length=hop_length, win_length=win_length)
    mag_spec = np.abs(mag_phase_spec)
    linear_spec = mag_spec ** power
    mel_basis = lc.mel_filterbank(sr=sampling_rate, n_fft=n_fft, n_mels=n_mels, fmin=fmin,
                                  fmax=fmax)
    return np.dot(mel_basis, linear_spec)


# ----------------------------------------------------------------------------------------------
# audio/synthesis.py
# ----------------------------------------------------------------------------------------------
def griffin_lim_v2(spectrogram, win_length, hop_length, n_fft, n_iter, angles=None,
                   batched_fft=False):
    """audio/synthesis.py:43-125.

    ``angles`` (complex, same shape as ``spectrogram``) overrides the initial phase; when None
    it is drawn exactly like the reference (:85) from the global ``np.random`` state, so seeding
    numpy before the call reproduces the reference's stream.  ``batched_fft`` only selects how
    the oracle's istft batches its FFT calls (values unchanged).
    """
    mse = None
    if angles is None:
        angles = np.exp(2j * np.pi * np.random.rand(*spectrogram.shape))              # :85
    for _ in range(n_iter):                                                             # :91
        full = np.abs(spectrogram).astype(np.complex128) * angles                       # :93
        estimated_signal = lc.istft(full, hop_length=hop_length, win_length=win_length,
                                    batched_fft=batched_fft)                            # :96-99
        estimated_stft = lc.stft(estimated_signal, n_fft=n_fft, win_length=win_length,
                                 hop_length=hop_length)                                 # :102-106
        # :109 -- np.angle of complex64 is float32; 1j * float32 stays single precision.
        angles = np.exp(np.complex64(1j) * np.angle(estimated_stft))
        mse = np.square(np.abs(spectrogram) - np.abs(estimated_stft)).mean()            # :112
    full = np.abs(spectrogram).astype(np.complex128) * angles                           # :117
    estimated_signal = lc.istft(full, hop_length=hop_length, win_length=win_length,
                                batched_fft=batched_fft)                                # :120-123
    return estimated_signal, mse


def spectrogram_to_wav(mag, win_length, hop_length, n_fft, n_iter, angles=None,
                       batched_fft=False):
    """audio/synthesis.py:5-40."""
    wav, _ = griffin_lim_v2(mag, win_length=win_length, hop_length=hop_length, n_fft=n_fft,
                            n_iter=n_iter, angles=angles, batched_fft=batched_fft)
    return wav.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# datasets/statistics.py (dB part)
# ----------------------------------------------------------------------------------------------
def decibel_statistics(wav, sampling_rate):
    """datasets/statistics.py:11-66 -- [min lin dB, max lin dB, min mel dB, max mel dB] with the
    hard-coded n_fft=1024 / hop=256 / win=1024 / 80 mels / fmax = sr // 2 analysis."""
    n_fft = 1024
    hop_length = n_fft // 4
    win_length = n_fft
    n_mels = 80
    linear_spec = linear_scale_spectrogram(wav, n_fft=n_fft, hop_length=hop_length,
                                           win_length=win_length)
    mel_spec = mel_scale_spectrogram(wav, n_fft=n_fft, sampling_rate=sampling_rate,
                                     n_mels=n_mels, fmin=0, fmax=sampling_rate // 2,
                                     hop_length=hop_length, win_length=win_length, power=1)
    linear_mag_db = magnitude_to_decibel(np.abs(linear_spec))
    mel_mag_db = magnitude_to_decibel(np.abs(mel_spec))
    return np.array([np.min(linear_mag_db), np.max(linear_mag_db),
                     np.min(mel_mag_db), np.max(mel_mag_db)])


def collect_decibel_statistics_from_wavs(wavs, sampling_rate):
    """datasets/statistics.py:69-98 minus the file decode: float64 running sum of the per-file
    4-vectors in listing order, divided by the file count."""
    stats = np.zeros(4)
    for wav in wavs:
        stats += decibel_statistics(wav, sampling_rate)
    stats /= len(wavs)
    return stats


# ----------------------------------------------------------------------------------------------
# datasets/dataset_helper.py + datasets/lj_speech.py (audio part)
# ----------------------------------------------------------------------------------------------
class LJSpeechConstants:
    """datasets/lj_speech.py:20-29."""
    mel_mag_ref_db = 6.02
    mel_mag_max_db = 99.89
    linear_ref_db = 35.66
    linear_mag_max_db = 100.0


class ModelParams:
    """tacotron/params/model.py:13-48 (hot-path values only)."""
    sampling_rate = 22050
    n_fft = 2048
    win_len = 50.0
    win_hop = 12.5
    n_mels = 80
    mel_fmin = 0
    mel_fmax = 8000
    reduction = 5
    magnitude_power = 1.3
    reconstruction_iterations = 50


def apply_reduction_padding(mel_mag_db, linear_mag_db, reduction_factor):
    """datasets/dataset_helper.py:357-401 -- zero-pad frames to a multiple of r and fold."""
    n_frames = mel_mag_db.shape[0]
    if (n_frames % reduction_factor) != 0:
        n_padding_frames = reduction_factor - (n_frames % reduction_factor)
        mel_mag_db = np.pad(mel_mag_db, [[0, n_padding_frames], [0, 0]], mode="constant")
        linear_mag_db = np.pad(linear_mag_db, [[0, n_padding_frames], [0, 0]], mode="constant")
    mel_mag_db = mel_mag_db.reshape((-1, mel_mag_db.shape[1] * reduction_factor))
    linear_mag_db = linear_mag_db.reshape((-1, linear_mag_db.shape[1] * reduction_factor))
    return mel_mag_db, linear_mag_db


def load_audio_from_wav(wav, sr, constants=LJSpeechConstants, params=ModelParams, trim=True):
    """datasets/lj_speech.py:106-156 after the file decode (:114).  ``trim=False`` skips the
    ``librosa.effects.trim`` step (:119) so the STFT -> dB -> normalise core can be compared on
    its own."""
    win_len = ms_to_samples(params.win_len, params.sampling_rate)
    hop_len = ms_to_samples(params.win_hop, params.sampling_rate)
    if trim:
        wav, _ = lc.trim(wav)
    linear_spec = linear_scale_spectrogram(wav, params.n_fft, hop_len, win_len).T
    mel_spec = mel_scale_spectrogram(wav, params.n_fft, sr, params.n_mels, params.mel_fmin,
                                     params.mel_fmax, hop_len, win_len, 1).T
    linear_mag_db = magnitude_to_decibel(np.abs(linear_spec))
    linear_mag_db = normalize_decibel(linear_mag_db, constants.linear_ref_db,
                                      constants.linear_mag_max_db)
    mel_mag_db = magnitude_to_decibel(np.abs(mel_spec))
    mel_mag_db = normalize_decibel(mel_mag_db, constants.mel_mag_ref_db, constants.mel_mag_max_db)
    if params.reduction > 1:
        mel_mag_db, linear_mag_db = apply_reduction_padding(mel_mag_db, linear_mag_db,
                                                            params.reduction)
    return np.array(mel_mag_db).astype(np.float32), np.array(linear_mag_db).astype(np.float32)


class BlizzardNancyConstants:      # datasets/blizzard_nancy.py:20-29
    mel_mag_ref_db = 9.55
    mel_mag_max_db = 100.0
    linear_ref_db = 36.50
    linear_mag_max_db = 100.0


class CMUConstants:                # datasets/cmu_slt.py:19-28
    mel_mag_ref_db = 9.33
    mel_mag_max_db = 100.0
    linear_ref_db = 36.50
    linear_mag_max_db = 100.0


class PAVOQUEConstants:            # datasets/pavoque.py:20-32
    mel_mag_ref_db = 12.63
    mel_mag_max_db = 100.0
    linear_ref_db = 24
    linear_mag_max_db = 100.0
    raw_silence_db = -15.0


def silence_interval_from_spectrogram(mag_spec_db, threshold_db, ref=np.max):
    """audio/effects.py:218-232."""
    ref_trim_spec_db = ref(mag_spec_db, axis=0)
    non_silent = np.array(ref_trim_spec_db > threshold_db, dtype=np.int32)
    nonzero = np.flatnonzero(non_silent)
    if len(nonzero) == 0:
        return None
    return np.min(nonzero), np.max(nonzero)


def pavoque_load_audio_from_wav(wav, sr, constants=PAVOQUEConstants, params=ModelParams):
    """datasets/pavoque.py:104-160 after the file decode (:112)."""
    win_len = ms_to_samples(params.win_len, params.sampling_rate)
    hop_len = ms_to_samples(params.win_hop, params.sampling_rate)
    linear_spec = linear_scale_spectrogram(wav, params.n_fft, hop_len, win_len).T
    linear_spec = np.array(linear_spec)
    linear_spec[:, 0:8] = 0                                                          # :119
    linear_mag_db = magnitude_to_decibel(np.abs(linear_spec))
    linear_mag_db = normalize_decibel(linear_mag_db, constants.linear_ref_db, constants.linear_mag_max_db)
    trim_start, trim_end = silence_interval_from_spectrogram(linear_mag_db, constants.raw_silence_db, np.max)
    mel_spec = mel_scale_spectrogram(wav, params.n_fft, sr, params.n_mels, params.mel_fmin,
                                     params.mel_fmax, hop_len, win_len, 1).T
    mel_mag_db = magnitude_to_decibel(np.abs(mel_spec))
    mel_mag_db = normalize_decibel(mel_mag_db, constants.mel_mag_ref_db, constants.mel_mag_max_db)
    linear_mag_db = linear_mag_db[trim_start:trim_end, :]
    mel_mag_db = mel_mag_db[trim_start:trim_end, :]
    if params.reduction > 1:
        mel_mag_db, linear_mag_db = apply_reduction_padding(mel_mag_db, linear_mag_db, params.reduction)
    return np.array(mel_mag_db).astype(np.float32), np.array(linear_mag_db).astype(np.float32)


def reconstruction_error(wav, sampling_rate, n_iters, angles=None):
    """Loop body of collect_reconstruction_error, datasets/statistics.py:156-181."""
    win = ms_to_samples(50.0, sampling_rate)
    hop = ms_to_samples(12.5, sampling_rate)
    mag = np.abs(linear_scale_spectrogram(wav, 2048, hop, win))
    _, mse = griffin_lim_v2(mag, win, hop, 2048, n_iters, angles=angles, batched_fft=True)
    return mse


def calculate_mfccs(mel_spec, sampling_rate, n_mfcc):
    """audio/features.py:89-113."""
    return lc.mfcc(mel_spec, n_mfcc=n_mfcc)


def time_stretch(wav, rate, angles=None):
    """audio/effects.py:46-86: STFT (1024 / 256 / 1024) -> phase vocoder -> |.| -> 25 Griffin-Lim
    iterations.  Only the magnitude of the stretched spectrogram is used (:80)."""
    if rate <= 0.0:
        raise ValueError('The fixed rate used to stretch the signal must be greater 0.')
    n_fft = 1024
    win_len = n_fft
    hop_len = win_len // 4
    stft = linear_scale_spectrogram(wav, n_fft, hop_len, win_len)
    mag = np.abs(lc.phase_vocoder(stft, rate))
    return spectrogram_to_wav(mag, win_len, hop_len, n_fft, 25, angles=angles, batched_fft=True)


# ----------------------------------------------------------------------------------------------
# tacotron/inference.py glue before Griffin-Lim
# ----------------------------------------------------------------------------------------------
def inference_postprocess(spectrogram, constants=LJSpeechConstants, params=ModelParams):
    """tacotron/inference.py:94-101,175 -- model output (T, 1025) in [0, 1] -> magnitude ** 1.3,
    oriented (1025, T).  De-normalises with the MEL constants, as the reference does."""
    linear_mag_db = inv_normalize_decibel(spectrogram.T, constants.mel_mag_ref_db,
                                          constants.mel_mag_max_db)
    linear_mag = decibel_to_magnitude(linear_mag_db)
    return np.power(linear_mag, params.magnitude_power)
