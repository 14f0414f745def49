"""``audio/io.py`` of the reference: wav decode / encode on the host.

File decode/encode is host I/O, not part of the device hot path.
The reference goes through ``librosa.load(sr=None)`` / ``librosa.output.write_wav``
(audio/io.py:30,53); here PCM wav files are read and written with ``scipy.io.wavfile`` with the
same conventions (float32 in [-1, 1), mono mix-down, native sampling rate).
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from scipy.io import wavfile

_IO_THREADS = 8


def load_wav(wav_path, sampling_rate=None, offset=0.0, duration=None):
    """reference audio/io.py:5-30 (same positional order) -- returns ``(float32 mono samples,
    sampling_rate)``.  Every caller of the reference loads at the file's own rate
    (``sampling_rate=None``: datasets/lj_speech.py:114, datasets/statistics.py:90,107,160); a target
    rate equal to the file's is accepted, another one raises ValueError instead of silently returning
    un-resampled audio (librosa would resample with resampy, which is not available here)."""
    sr, data = wavfile.read(wav_path)
    if sampling_rate is not None and int(sampling_rate) != int(sr):
        raise ValueError('load_wav: resampling is not supported (file rate {} Hz, requested {} Hz); '
                         'pass sampling_rate=None'.format(sr, sampling_rate))
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


def load_wav_pcm16(wav_path):
    """``(int16 mono samples, sampling_rate)`` when the file holds 16-bit mono PCM (the LJSpeech format),
    else the float32 result of :func:`load_wav`.  The batched feature calls accept int16 clips and do the
    ``/ 32768`` conversion on the device (``sstts_pcm16_to_float``), which halves the upload."""
    sr, data = wavfile.read(wav_path)
    if data.dtype == np.int16 and data.ndim == 1:
        return np.ascontiguousarray(data), sr
    return load_wav(wav_path)


def save_wav(wav_path, wav, sampling_rate, norm=False):
    """reference audio/io.py:33-53 -- float wav writer, optional peak normalisation."""
    wav = np.asarray(wav)
    if norm and wav.size and np.max(np.abs(wav)) > 0:
        wav = wav / np.max(np.abs(wav))
    wavfile.write(wav_path, int(sampling_rate), wav.astype(np.float32))


def load_wavs(paths, threads=_IO_THREADS, pcm16=False):
    """Decode a list of wav files with a few threads (file reads and the int16 -> float32 conversion
    release the GIL); same result, same order as ``[load_wav(p) for p in paths]``.  ``pcm16=True`` keeps
    16-bit mono files as int16 (:func:`load_wav_pcm16`) when EVERY file of the list is such a file."""
    paths = list(paths)
    loader = load_wav_pcm16 if pcm16 else load_wav
    if len(paths) < 2 or threads < 2:
        out = [loader(p) for p in paths]
    else:
        with ThreadPoolExecutor(max_workers=min(threads, len(paths))) as ex:
            out = list(ex.map(loader, paths))
    if pcm16 and not all(w.dtype == np.int16 for w, _ in out):
        out = [((w.astype(np.float32) / 32768.0) if w.dtype == np.int16 else w, sr) for w, sr in out]
    return out


def prefetch_batches(paths, batch, threads=_IO_THREADS, pcm16=False):
    """Yield ``(paths[s:s + batch], load_wavs(...))`` with the NEXT batch being decoded in the
    background while the caller works on the current one (the device calls block the caller's
    thread only until their final synchronisation)."""
    paths = list(paths)
    starts = list(range(0, len(paths), batch))
    if not starts:
        return
    with ThreadPoolExecutor(max_workers=1) as ex:
        fut = ex.submit(load_wavs, paths[0:batch], threads, pcm16)
        for i, s in enumerate(starts):
            cur = fut.result()
            if i + 1 < len(starts):
                n = starts[i + 1]
                fut = ex.submit(load_wavs, paths[n:n + batch], threads, pcm16)
            yield paths[s:s + batch], cur


def save_npz_many(items, threads=_IO_THREADS):
    """Write ``(path, {key: array})`` pairs with ``np.savez`` from a few threads (uncompressed zip:
    byte-compatible with the reference's files, datasets/dataset_helper.py:355)."""
    items = list(items)

    def write(item):
        path, arrays = item
        np.savez(path, **arrays)

    if len(items) < 2 or threads < 2:
        for it in items:
            write(it)
        return
    with ThreadPoolExecutor(max_workers=min(threads, len(items))) as ex:
        list(ex.map(write, items))
