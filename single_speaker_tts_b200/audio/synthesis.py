"""``audio/synthesis.py`` of the reference, executed by the sm_100a Griffin-Lim kernels.

``spectrogram_to_wav`` / ``griffin_lim_v2`` keep the reference signatures (audio/synthesis.py:5,43)
including the use of numpy's GLOBAL random state for the initial phase (:85), so
``np.random.seed(s)`` before a call selects the same initial phase as in the reference.
``spectrograms_to_wavs`` is the batched entry point that replaces the 6-worker pool of
``tacotron/inference.py:185-188`` / ``tacotron/serve.py:69-72``.
"""
import numpy as np

from .. import _runtime


def _initial_uniform(shape):
    # reference audio/synthesis.py:85 draws np.random.rand(*spectrogram.shape) from numpy's GLOBAL stream;
    # the exp(2j * pi * u) around it runs on the device (sstts_phase_from_uniform)
    return np.random.rand(*shape)


def _check_reanalysis_possible(spectrogram, n_iter):
    """A single-frame spectrogram gives an EMPTY signal (hop * (T - 1) == 0 samples) whose re-analysis
    fails in the reference: ``librosa.stft`` reflect-pads it and numpy raises ValueError
    (audio/synthesis.py:96-106).  The per-item functions keep that error behaviour."""
    if n_iter > 0 and spectrogram.ndim == 2 and spectrogram.shape[1] == 1:
        raise ValueError("can't extend empty axis 0 using modes other than 'constant' or 'empty'")


def _phase_arguments(shape, angles, device_phase):
    if angles is not None:
        return {'angles': [angles]}
    if device_phase:
        # one 63-bit seed from numpy's global stream (np.random.seed still makes the call reproducible);
        # the phasors themselves come from the counter-based generator inside the first launch
        return {'seed': int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))}
    return {'uniform': [_initial_uniform(shape)]}


def griffin_lim_v2(spectrogram, win_length, hop_length, n_fft, n_iter, angles=None,
                   precision='f32', device_phase=False):
    """Griffin-Lim reconstruction -- reference audio/synthesis.py:43-125.

    Returns ``(audio float32 of length hop*(T-1), mse)``; ``mse`` is the mean squared magnitude
    error of the last iteration (None when ``n_iter`` == 0, like the reference).
    ``angles`` (extension) overrides the random initial phase; ``device_phase=True`` (extension)
    draws it on the device instead of taking ``np.random.rand(*shape)`` from numpy's global stream --
    the 1,025 x T host draws cost several times the whole reconstruction of one utterance and are
    serialised across threads by the RandomState lock (tacotron/serve.py:69-72 calls from six threads).
    """
    spectrogram = np.asarray(spectrogram)
    _check_reanalysis_possible(spectrogram, n_iter)
    phase = _phase_arguments(spectrogram.shape, angles, device_phase)
    wavs, mses = _runtime.griffin_lim_batch([spectrogram], win_length, hop_length, n_fft, n_iter,
                                            precision=precision, return_mse=n_iter > 0, **phase)
    return wavs[0], (mses[0] if mses is not None else None)


def spectrogram_to_wav(mag, win_length, hop_length, n_fft, n_iter, angles=None, precision='f32',
                       device_phase=False):
    """Magnitude spectrogram -> float32 waveform -- reference audio/synthesis.py:5-40 (``angles`` /
    ``device_phase``: see :func:`griffin_lim_v2`)."""
    mag = np.asarray(mag)
    _check_reanalysis_possible(mag, n_iter)
    phase = _phase_arguments(mag.shape, angles, device_phase)
    wavs, _ = _runtime.griffin_lim_batch([mag], win_length, hop_length, n_fft, n_iter,
                                         precision=precision, **phase)
    return wavs[0].astype(np.float32)


def spectrograms_to_wavs(mags, win_length, hop_length, n_fft, n_iter, angles=None, seed=None,
                         precision='f32', return_mse=False, normalize_peak=False):
    """Batched Griffin-Lim over a ragged list of (1 + n_fft/2, T_i) magnitude spectrograms.

    One device call for the whole batch (utterances packed with an offsets table).  ``angles`` is
    an optional list of initial phasors (one per item); otherwise the phase is generated on the
    device from ``seed`` (None: one seed is drawn from numpy's global RNG).
    ``normalize_peak`` divides every waveform by its peak on the device (the ``norm=True`` of the
    reference's ``save_wav``, tacotron/inference.py:199).
    Returns a list of float32 waveforms (and the list of last-iteration mse values if requested).
    """
    wavs, mses = _runtime.griffin_lim_batch(list(mags), win_length, hop_length, n_fft, n_iter,
                                            angles=angles, seed=seed, precision=precision,
                                            return_mse=return_mse, normalize_peak=normalize_peak)
    return (wavs, mses) if return_mse else wavs


def model_outputs_to_wavs(spectrograms, ref_db, max_db, magnitude_power, win_length, hop_length, n_fft,
                          n_iter, seed=None, precision='f32', normalize_peak=False):
    """Model output -> waveforms in one device call (extension; replaces
    tacotron/inference.py:92-101 + :170-188 and tacotron/serve.py:39-72).

    ``spectrograms``: iterable of normalised linear spectrograms exactly as ``session.run`` returns
    them, ``(T, 1 + n_fft/2)`` float32 in [0, 1] (a ``(B, T, bins)`` array works too).  The
    de-normalisation ``inv_normalize_decibel(spec.T, ref_db, max_db)`` -> ``decibel_to_magnitude``
    -> ``** magnitude_power`` runs fused on the device (the reference passes the MEL constants
    here, tacotron/inference.py:96-98), followed by ``n_iter`` Griffin-Lim iterations.  Raises the
    reference's AssertionError if a de-normalised value falls below -100 dB.
    """
    specs = [np.asarray(s) for s in spectrograms]
    wavs, _ = _runtime.griffin_lim_batch(specs, win_length, hop_length, n_fft, n_iter, seed=seed,
                                         precision=precision,
                                         denormalize=(ref_db, max_db, magnitude_power),
                                         normalize_peak=normalize_peak)
    return wavs
