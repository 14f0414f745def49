"""numpy/scipy restatement of the librosa 0.6.x numerics used by the reference hot path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  librosa is a third-party dependency of the
reference (``/root/reference/requirements.txt:2``: ``librosa >= 0.6.1``, effective window
0.6.1 .. 0.7.2), absent from ``/root/reference`` and from this image.  Each function below
restates the published 0.6.x algorithm and names the reference call sites that depend on it.

Precision notes that matter for parity (they are part of the reference's behaviour):
  * ``stft`` multiplies a float64 window with float32 frames (-> float64), transforms in
    float64 (scipy.fftpack) and stores the first ``1 + n_fft//2`` bins as complex64 in a
    Fortran-ordered (frame-major in memory) matrix.
  * ``istft`` Hermitian-extends every column, runs a complex ``ifft`` in the input precision
    (the reference hands it complex128, ``audio/synthesis.py:93``), takes ``.real``, multiplies
    by the float64 window and overlap-adds frame by frame into a **float32** buffer; the
    window-sum-of-squares is accumulated in float32 as well and divided out where it exceeds
    ``np.finfo(float32).tiny``.
"""
import numpy as np
import scipy.fftpack as fftpack

# librosa.util.MAX_MEM_BLOCK (0.6.x): column-block size of the STFT loop.  It only affects
# how many frames are transformed per fftpack call, not the values.
MAX_MEM_BLOCK = 2 ** 8 * 2 ** 10


def hann_window(win_length):
    """``scipy.signal.get_window('hann', win_length, fftbins=True)`` -- periodic Hann, float64.

    Used through ``librosa.filters.get_window`` by stft/istft
    (reference: ``audio/synthesis.py:82,96-106``; ``audio/features.py:62,145``).
    """
    n = np.arange(win_length, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)


def pad_center(data, size):
    """``librosa.util.pad_center`` for 1-D data: zero-pad to ``size`` with ``lpad = (size-n)//2``."""
    n = data.shape[0]
    lpad = int((size - n) // 2)
    if lpad < 0:
        raise ValueError('Target size ({:d}) must be at least input size ({:d})'.format(size, n))
    return np.pad(data, (lpad, int(size - n - lpad)), mode='constant')


def padded_window(win_length, n_fft):
    """Hann(win_length) centred in n_fft samples (float64) -- the window stft/istft apply."""
    return pad_center(hann_window(win_length), n_fft)


def frame(y, frame_length, hop_length):
    """``librosa.util.frame``: (frame_length, n_frames) strided view, n = 1 + (len-frame)//hop."""
    if len(y) < frame_length:
        raise ValueError('Buffer is too short (n={:d}) for frame_length={:d}'.format(len(y), frame_length))
    n_frames = 1 + int((len(y) - frame_length) / hop_length)
    y = np.ascontiguousarray(y)
    return np.lib.stride_tricks.as_strided(y, shape=(frame_length, n_frames),
                                           strides=(y.itemsize, hop_length * y.itemsize))


def stft(y, n_fft=2048, hop_length=None, win_length=None):
    """``librosa.stft(y, n_fft, hop_length, win_length, window='hann', center=True,
    dtype=complex64, pad_mode='reflect')`` (0.6.x).

    Reference call sites: ``audio/synthesis.py:102-106``, ``audio/features.py:62,145``.
    Returns a (1 + n_fft//2, T) complex64 array in Fortran order, ``T = 1 + len(y)//hop``.
    """
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    fft_window = padded_window(win_length, n_fft).reshape((-1, 1))
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError('Invalid shape for monophonic audio: ndim={:d}'.format(y.ndim))
    y = np.pad(y, int(n_fft // 2), mode='reflect')
    y_frames = frame(y, frame_length=n_fft, hop_length=hop_length)
    stft_matrix = np.empty((int(1 + n_fft // 2), y_frames.shape[1]), dtype=np.complex64, order='F')
    n_columns = int(MAX_MEM_BLOCK / (stft_matrix.shape[0] * stft_matrix.itemsize))
    for bl_s in range(0, stft_matrix.shape[1], n_columns):
        bl_t = min(bl_s + n_columns, stft_matrix.shape[1])
        stft_matrix[:, bl_s:bl_t] = fftpack.fft(fft_window * y_frames[:, bl_s:bl_t],
                                                axis=0)[:stft_matrix.shape[0]]
    return stft_matrix


def window_sumsquare(n_frames, hop_length, win_length, n_fft, dtype=np.float32):
    """``librosa.filters.window_sumsquare('hann', ...)``: sum of squared, centred windows.

    The 0.6.x fill loop adds the float64 squared window into a ``dtype`` (float32) buffer
    frame by frame, i.e. every partial sum is rounded to float32.
    """
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = pad_center(hann_window(win_length) ** 2, n_fft)
    for i in range(n_frames):
        sample = i * hop_length
        x[sample:min(n, sample + n_fft)] += win_sq[:max(0, min(n_fft, n - sample))]
    return x


def istft(stft_matrix, hop_length=None, win_length=None, batched_fft=False):
    """``librosa.istft(stft_matrix, hop_length, win_length, window='hann', center=True,
    dtype=float32)`` (0.6.x).  Reference call sites: ``audio/synthesis.py:96-99,120-123``.

    ``batched_fft=True`` transforms all columns with one fftpack call instead of one call per
    frame; the overlap-add below is unchanged (frame order, float32 accumulator), so the result
    is the same up to the FFT library's batching (checked in ``tests/test_oracle.py``).
    Returns float32 of length ``hop * (T - 1)``.
    """
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    ifft_window = padded_window(win_length, n_fft)
    n_frames = stft_matrix.shape[1]
    expected_signal_len = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(expected_signal_len, dtype=np.float32)
    if batched_fft:
        full = np.concatenate((stft_matrix, stft_matrix[-2:0:-1].conj()), 0)
        frames = ifft_window[:, None] * fftpack.ifft(full, axis=0).real
    for i in range(n_frames):
        sample = i * hop_length
        if batched_fft:
            ytmp = frames[:, i]
        else:
            spec = stft_matrix[:, i].flatten()
            spec = np.concatenate((spec, spec[-2:0:-1].conj()), 0)
            ytmp = ifft_window * fftpack.ifft(spec).real
        y[sample:(sample + n_fft)] = y[sample:(sample + n_fft)] + ytmp
    ifft_window_sum = window_sumsquare(n_frames, hop_length, win_length, n_fft, dtype=np.float32)
    approx_nonzero_indices = ifft_window_sum > np.finfo(np.float32).tiny
    y[approx_nonzero_indices] /= ifft_window_sum[approx_nonzero_indices]
    return y[int(n_fft // 2):-int(n_fft // 2)]


def hz_to_mel_htk(frequencies):
    """``librosa.hz_to_mel(f, htk=True)``."""
    return 2595.0 * np.log10(1.0 + np.asanyarray(frequencies, dtype=np.float64) / 700.0)


def mel_to_hz_htk(mels):
    """``librosa.mel_to_hz(m, htk=True)``."""
    return 700.0 * (10.0 ** (np.asanyarray(mels, dtype=np.float64) / 2595.0) - 1.0)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=True, norm=1)`` (0.6.x, float64).

    Reference call site: ``audio/features.py:75-80``.  Returns (n_mels, 1 + n_fft//2).
    """
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)))
    fftfreqs = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    mel_f = mel_to_hz_htk(np.linspace(hz_to_mel_htk(fmin), hz_to_mel_htk(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    """``librosa.power_to_db`` (0.6.x)."""
    magnitude = np.abs(np.asarray(S))
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def rmse(y, frame_length=2048, hop_length=512):
    """``librosa.feature.rmse(y=y, frame_length, hop_length, center=True, pad_mode='reflect')``."""
    y = np.pad(y, int(frame_length // 2), mode='reflect')
    x = frame(y, frame_length=frame_length, hop_length=hop_length)
    return np.sqrt(np.mean(np.abs(x) ** 2, axis=0, keepdims=True))


def trim(y, top_db=60, frame_length=2048, hop_length=512):
    """``librosa.effects.trim(y, top_db=60, ref=np.max, frame_length=2048, hop_length=512)``.

    Reference call site: ``datasets/lj_speech.py:119`` (and siblings).  Returns
    ``(y[start:end], np.array([start, end]))``.
    """
    mse = rmse(y, frame_length=frame_length, hop_length=hop_length) ** 2
    non_silent = power_to_db(mse.squeeze(), ref=np.max, top_db=None) > -top_db
    nonzero = np.flatnonzero(non_silent)
    if nonzero.size > 0:
        start = int(nonzero[0] * hop_length)
        end = min(y.shape[-1], int((nonzero[-1] + 1) * hop_length))
    else:
        start, end = 0, 0
    return y[start:end], np.asarray([start, end])


def phase_vocoder(D, rate, hop_length=None):
    """librosa.core.phase_vocoder (0.6.x) -- used by the reference at audio/effects.py:77.
    D: (1 + n_fft/2, T) complex; returns (1 + n_fft/2, len(arange(0, T, rate))) of D's dtype."""
    n_fft = 2 * (D.shape[0] - 1)
    if hop_length is None:
        hop_length = int(n_fft // 4)
    time_steps = np.arange(0, D.shape[1], rate, dtype=np.float64)
    d_stretch = np.zeros((D.shape[0], len(time_steps)), D.dtype, order='F')
    phi_advance = np.linspace(0, np.pi * hop_length, D.shape[0])
    phase_acc = np.angle(D[:, 0])
    D = np.pad(D, [(0, 0), (0, 2)], mode='constant')
    for (t, step) in enumerate(time_steps):
        columns = D[:, int(step):int(step + 2)]
        alpha = np.mod(step, 1.0)
        mag = ((1.0 - alpha) * np.abs(columns[:, 0]) + alpha * np.abs(columns[:, 1]))
        d_stretch[:, t] = mag * np.exp(1.j * phase_acc)
        dphase = (np.angle(columns[:, 1]) - np.angle(columns[:, 0]) - phi_advance)
        dphase = dphase - 2.0 * np.pi * np.round(dphase / (2.0 * np.pi))
        phase_acc += phi_advance + dphase
    return d_stretch


def dct_filters(n_filters, n_input):
    """librosa.filters.dct (0.6.x): orthonormal DCT-II basis, shape (n_filters, n_input)."""
    basis = np.empty((n_filters, n_input))
    basis[0, :] = 1.0 / np.sqrt(n_input)
    samples = np.arange(1, 2 * n_input, 2) * np.pi / (2.0 * n_input)
    for i in range(1, n_filters):
        basis[i, :] = np.cos(i * samples) * np.sqrt(2.0 / n_input)
    return basis


def mfcc(S, n_mfcc=20):
    """librosa.feature.mfcc(S=S, n_mfcc=n_mfcc) (0.6.x): ``np.dot(filters.dct(n_mfcc, S.shape[0]), S)``;
    0.7 computes the same numbers with scipy.fftpack.dct(type=2, norm='ortho')."""
    return np.dot(dct_filters(n_mfcc, S.shape[0]), S)
