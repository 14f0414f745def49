"""Audio part of ``datasets/dataset_helper.py`` and the four corpus loaders of the reference
(``datasets/lj_speech.py``, ``blizzard_nancy.py``, ``cmu_slt.py``, ``pavoque.py``).

Only the feature side is restated (``load_audio``, ``apply_reduction_padding``,
``pre_compute_features``); the text side (vocabulary, sentence ids) is host string processing that
stays with the reference's own code.
"""
import os

import numpy as np

from ..audio.conversion import ms_to_samples
from ..audio.features import features_batch
from ..audio.io import load_wav, prefetch_batches, save_npz_many
from ..params import model_params


class DatasetHelper:
    """Feature recipe shared by the corpus loaders.  Subclasses provide the four dB constants."""
    mel_mag_ref_db = None
    mel_mag_max_db = None
    linear_ref_db = None
    linear_mag_max_db = None

    @staticmethod
    def apply_reduction_padding(mel_mag_db, linear_mag_db, reduction_factor):
        """reference datasets/dataset_helper.py:357-401: zero-pad frames to a multiple of r, fold."""
        n_frames = mel_mag_db.shape[0]
        if (n_frames % reduction_factor) != 0:
            n_padding_frames = reduction_factor - (n_frames % reduction_factor)
            mel_mag_db = np.pad(mel_mag_db, [[0, n_padding_frames], [0, 0]], mode="constant")
            linear_mag_db = np.pad(linear_mag_db, [[0, n_padding_frames], [0, 0]], mode="constant")
        mel_mag_db = mel_mag_db.reshape((-1, mel_mag_db.shape[1] * reduction_factor))
        linear_mag_db = linear_mag_db.reshape((-1, linear_mag_db.shape[1] * reduction_factor))
        return mel_mag_db, linear_mag_db

    @classmethod
    def features_from_wavs(cls, wavs, sampling_rate=None, trim_silence=True, precision='f64'):
        """Batched core of ``load_audio`` (reference datasets/lj_speech.py:119-156) for decoded
        clips: trim -> one STFT -> linear + mel dB, normalised with the class constants ->
        reduction padding -> float32.  Returns a list of ``(mel, lin)``."""
        sr = sampling_rate or model_params.sampling_rate
        win_len = ms_to_samples(model_params.win_len, model_params.sampling_rate)
        hop_len = ms_to_samples(model_params.win_hop, model_params.sampling_rate)
        # librosa.effects.trim defaults (datasets/lj_speech.py:119): top_db 60, frames 2048 / 512;
        # runs as a device kernel on the same upload the feature kernel reads
        return features_batch(wavs, model_params.n_fft, hop_len, win_len, sr, model_params.n_mels,
                              model_params.mel_fmin, model_params.mel_fmax, cls.linear_ref_db,
                              cls.linear_mag_max_db, cls.mel_mag_ref_db, cls.mel_mag_max_db,
                              reduction=model_params.reduction, precision=precision,
                              trim=(60.0, 2048, 512) if trim_silence else None)

    @classmethod
    def load_audio(cls, file_path):
        """reference datasets/lj_speech.py:106-156 -- ``file_path`` is ``bytes`` (tf.py_func)."""
        wav, sr = load_wav(file_path.decode())
        return cls.features_from_wavs([wav], sampling_rate=sr)[0]

    @classmethod
    def pre_compute_features(cls, paths, batch_clips=256):
        """reference datasets/dataset_helper.py:326-355 -- ``<name>.npz`` next to every wav with
        keys ``mel_mag_db`` / ``linear_mag_db``; clips go to the device in batches."""
        n_samples = len(paths)
        print('Loaded {} dataset entries.'.format(n_samples))
        # decode of batch k + 1 (threads) overlaps the device work and the .npz writes of batch k
        # 16-bit mono files (LJSpeech) go to the device as int16 and are converted there
        for chunk, loaded in prefetch_batches(paths, batch_clips, pcm16=True):
            wavs, srs = zip(*loaded)
            # the mel filterbank depends on the file's own sampling rate (datasets/lj_speech.py:129-131
            # passes each file's sr): one device batch per distinct rate of the chunk
            feats = [None] * len(wavs)
            for sr in sorted(set(srs)):
                idx = [i for i, r in enumerate(srs) if r == sr]
                for i, f in zip(idx, cls.features_from_wavs([wavs[i] for i in idx], sampling_rate=sr)):
                    feats[i] = f
            items = []
            for wav_path, (mel_mag_db, linear_mag_db) in zip(chunk, feats):
                out_path = '{}.npz'.format(os.path.splitext(wav_path)[0])
                print('Writing: "{}"'.format(out_path))
                items.append((out_path, {'mel_mag_db': mel_mag_db, 'linear_mag_db': linear_mag_db}))
            save_npz_many(items)


class LJSpeechDatasetHelper(DatasetHelper):
    """dB constants of reference datasets/lj_speech.py:20-29."""
    mel_mag_ref_db = 6.02
    mel_mag_max_db = 99.89
    linear_ref_db = 35.66
    linear_mag_max_db = 100.0


class BlizzardNancyDatasetHelper(DatasetHelper):
    """dB constants of reference datasets/blizzard_nancy.py:20-29 (same recipe as LJSpeech, :88-138)."""
    mel_mag_ref_db = 9.55
    mel_mag_max_db = 100.0
    linear_ref_db = 36.50
    linear_mag_max_db = 100.0


class CMUDatasetHelper(DatasetHelper):
    """dB constants of reference datasets/cmu_slt.py:19-28 (same recipe as LJSpeech, :87-137)."""
    mel_mag_ref_db = 9.33
    mel_mag_max_db = 100.0
    linear_ref_db = 36.50
    linear_mag_max_db = 100.0


def silence_interval_from_spectrogram(mag_spec_db, threshold_db, ref=np.max):
    """reference audio/effects.py:218-232, restated as written (``ref`` reduces over axis 0)."""
    ref_trim_spec_db = ref(mag_spec_db, axis=0)
    nonzero = np.flatnonzero(np.array(ref_trim_spec_db > threshold_db, dtype=np.int32))
    if len(nonzero) == 0:
        return None
    return np.min(nonzero), np.max(nonzero)


class PAVOQUEDatasetHelper(DatasetHelper):
    """reference datasets/pavoque.py:20-32,104-160: no waveform trimming; the first 8 linear bins are
    zeroed (``linear_spec[:, 0:8] = 0`` -> the -100 dB floor), and rows are cut with
    ``silence_interval_from_spectrogram`` of the normalised linear spectrogram before the reduction
    padding.  The STFT / dB / mel work is the same device batch as for the other corpora; the
    corpus-specific slicing is host glue on the returned arrays."""
    mel_mag_ref_db = 12.63
    mel_mag_max_db = 100.0
    linear_ref_db = 24
    linear_mag_max_db = 100.0
    raw_silence_db = -15.0

    @classmethod
    def features_from_wavs(cls, wavs, sampling_rate=None, trim_silence=True, precision='f64'):
        sr = sampling_rate or model_params.sampling_rate
        win_len = ms_to_samples(model_params.win_len, model_params.sampling_rate)
        hop_len = ms_to_samples(model_params.win_hop, model_params.sampling_rate)
        n_bins = 1 + model_params.n_fft // 2
        feats = features_batch(wavs, model_params.n_fft, hop_len, win_len, sr, model_params.n_mels,
                               model_params.mel_fmin, model_params.mel_fmax, cls.linear_ref_db,
                               cls.linear_mag_max_db, cls.mel_mag_ref_db, cls.mel_mag_max_db,
                               reduction=1, precision=precision, trim=None)
        # normalised value of a zeroed bin: magnitude_to_decibel(0) = -100 dB (audio/conversion.py:29)
        floor = np.float32(np.clip(1.0 + (np.float32(-100.0) - cls.linear_ref_db) /
                                   (abs(cls.linear_ref_db) + abs(cls.linear_mag_max_db)), 0.0, 1.0))
        out = []
        for mel, lin in feats:
            lin = lin.reshape(-1, n_bins).copy()
            mel = mel.reshape(-1, model_params.n_mels)
            lin[:, 0:8] = floor
            interval = silence_interval_from_spectrogram(lin, cls.raw_silence_db, np.max)
            if interval is None:
                raise TypeError('cannot unpack non-iterable NoneType object')   # as the reference (:131)
            trim_start, trim_end = interval
            lin, mel = lin[trim_start:trim_end, :], mel[trim_start:trim_end, :]
            if model_params.reduction > 1:
                mel, lin = DatasetHelper.apply_reduction_padding(mel, lin, model_params.reduction)
            out.append((np.array(mel).astype(np.float32), np.array(lin).astype(np.float32)))
        return out
