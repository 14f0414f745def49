"""Host-side logic and the C-ABI boundary, without a GPU: the library loads and exports every
symbol include/sstts.h declares, argument errors surface as Python exceptions, there is no CPU
fallback, and the numpy-level helpers equal the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import librosa_compat as lc
from oracle import reference_audio as ra
from single_speaker_tts_b200 import _lib, _runtime, distributed
from single_speaker_tts_b200._lib import SsttsError
from single_speaker_tts_b200.audio import conversion, effects
from single_speaker_tts_b200.datasets.dataset_helper import DatasetHelper, LJSpeechDatasetHelper
from single_speaker_tts_b200.datasets.statistics import reduce_decibel_statistics
from single_speaker_tts_b200.synthetic import ClipPlan, make_clips, speech_like_clip

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'sstts.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sstts_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_lib.SIGNATURES) == names            # the binding covers exactly the header
    assert lib.sstts_version() == 100


def test_plan_argument_errors_are_reported():
    lib = _lib.load()
    cfg = _runtime._make_config(4096, 1102, 275, 'f32')
    h = ctypes.c_void_p()
    off = np.array([0, 10], dtype=np.int64)
    rc = lib.sstts_gl_plan_create(ctypes.byref(cfg), 1, off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                  ctypes.byref(h))
    assert rc == -1 and b'n_fft' in lib.sstts_last_error()
    with pytest.raises(_lib.SsttsError):
        _lib.check(rc)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback():
    from single_speaker_tts_b200.audio import features, synthesis
    mag = np.ones((1025, 4), np.float32)
    with pytest.raises(_lib.SsttsError):
        synthesis.spectrogram_to_wav(mag, 1102, 275, 2048, 1)
    with pytest.raises(_lib.SsttsError):
        features.linear_scale_spectrogram(np.zeros(1000, np.float32), 2048, 275, 1102)


def test_conversion_module_equals_oracle():
    rng = np.random.default_rng(0)
    mag = rng.random((1025, 7)).astype(np.float32) * 3
    db = conversion.magnitude_to_decibel(mag)
    assert db.dtype == np.float32 and np.array_equal(db, ra.magnitude_to_decibel(mag))
    n = conversion.normalize_decibel(db, 35.66, 100.0)
    assert np.array_equal(n, ra.normalize_decibel(db, 35.66, 100.0))
    assert np.array_equal(conversion.inv_normalize_decibel(n, 6.02, 99.89), ra.inv_normalize_decibel(n, 6.02, 99.89))
    assert np.array_equal(conversion.decibel_to_magnitude(db), ra.decibel_to_magnitude(db))
    with pytest.raises(AssertionError, match='smaller -100 dB'):
        conversion.decibel_to_magnitude(np.array([-101.0]))
    assert conversion.ms_to_samples(50.0, 22050) == 1102 and conversion.ms_to_samples(12.5, 22050) == 275
    assert conversion.samples_to_ms(22050, 22050) == 1000 and conversion.get_duration(np.zeros(44100), 22050) == 2.0


def test_trim_has_no_host_implementation():
    """audio.effects.trim is the device kernel with the reference's call shape: without a CUDA device it
    raises like every other compute entry (its values are checked in tests/test_gpu_parity.py and, for the
    kernel source, in tests/test_emulator.py::test_trim_bounds_match_librosa_trim)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA device present')
    with pytest.raises(SsttsError):
        effects.trim(np.ones(5000, np.float32))
    y, idx = effects.trim(np.zeros(0, np.float32))          # librosa returns the empty clip unchanged
    assert y.shape == (0,) and tuple(idx) == (0, 0)


def test_geometry_is_validated_up_front_with_value_errors():
    """Unsupported STFT geometries raise ValueError naming the supported set before any device work
    (they used to surface late as SsttsError from plan creation / 'does not fit')."""
    v = _runtime.validate_geometry
    v(2048, 1102, 275, griffin_lim=True); v(1024, 1024, 256, griffin_lim=True); v(512, 512, 128); v(2048, 2048, 2048)
    for bad in ((4096, 4096, 1024), (2048, 1101, 275), (2048, 2050, 275), (2048, 1102, 0), (2048, 1102, 4096)):
        with pytest.raises(ValueError, match='unsupported|must be'):
            v(*bad)
    for bad in ((2048, 1102, 137), (2048, 1102, 1200)):          # win / hop > 5, hop > win
        with pytest.raises(ValueError, match='Griffin-Lim'):
            v(*bad, griffin_lim=True)
    with pytest.raises(ValueError, match='unsupported n_fft'):    # raised before the CUDA check
        _runtime.griffin_lim_batch([np.ones((2049, 4), np.float32)], 4096, 1024, 4096, 1)
    with pytest.raises(ValueError, match='unsupported n_fft'):
        _runtime.stft_features_batch([np.ones(5000, np.float32)], 4096, 1024, 4096)


def test_load_wav_keeps_the_reference_signature(tmp_path):
    """audio/io.py:5 -- load_wav(wav_path, sampling_rate=None, offset=0.0, duration=None): the second
    positional argument is the target rate, not an offset."""
    from scipy.io import wavfile
    from single_speaker_tts_b200.audio.io import load_wav
    x = (np.arange(22050) % 100 - 50).astype(np.int16) * 200
    path = str(tmp_path / 'a.wav')
    wavfile.write(path, 22050, x)
    w, sr = load_wav(path)
    assert sr == 22050 and w.dtype == np.float32 and len(w) == 22050
    w2, _ = load_wav(path, 22050)                      # same rate: accepted, full clip (not offset = 22050 s)
    assert np.array_equal(w, w2)
    w3, _ = load_wav(path, None, 0.5, 0.25)
    assert np.array_equal(w3, w[11025:11025 + 5512])
    with pytest.raises(ValueError, match='resampling'):
        load_wav(path, 16000)


def test_reduction_padding_equals_oracle():
    mel = np.random.default_rng(2).random((13, 80)).astype(np.float32)
    lin = np.random.default_rng(3).random((13, 1025)).astype(np.float32)
    a = DatasetHelper.apply_reduction_padding(mel, lin, 5)
    b = ra.apply_reduction_padding(mel, lin, 5)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert (LJSpeechDatasetHelper.mel_mag_ref_db, LJSpeechDatasetHelper.mel_mag_max_db,
            LJSpeechDatasetHelper.linear_ref_db, LJSpeechDatasetHelper.linear_mag_max_db) == (6.02, 99.89, 35.66, 100.0)


def test_shard_by_cost_is_balanced_partition():
    rng = np.random.default_rng(4)
    costs = rng.integers(81, 803, 256)
    for world in (1, 2, 4, 8):
        shards = distributed.shard_by_cost(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(256))
        loads = [costs[s].sum() for s in shards]
        assert max(loads) - min(loads) <= costs.max()


def test_reduce_statistics_is_mean_in_listing_order():
    rows = np.random.default_rng(5).normal(size=(37, 4)) * 30
    ref = np.zeros(4)
    for r in rows:
        ref += r
    ref /= len(rows)
    assert np.array_equal(reduce_decibel_statistics(rows), ref)
    mean, mn, mx, table = distributed.reduce_corpus_statistics(rows, np.arange(37), 37)
    assert np.array_equal(mean, ref) and np.array_equal(mn, rows.min(0)) and np.array_equal(mx, rows.max(0))


def test_synthetic_clips_are_seeded_and_shaped():
    a = make_clips(4, seed=1)
    b = make_clips(4, seed=1)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert all(x.dtype == np.float32 and 22050 <= len(x) <= 220500 and abs(np.abs(x).max() - 0.5) < 1e-6 for x in a)
    c = make_clips(6, seed=2, kind='ljspeech', pool=2)
    assert all(int(1.11 * 22050) <= len(x) <= int(10.1 * 22050) + 1 for x in c)


def test_clip_plan_is_identical_on_every_rank():
    a = ClipPlan(50, seed=3, kind='ljspeech', pool=2)
    b = ClipPlan(50, seed=3, kind='ljspeech', pool=2)
    assert np.array_equal(a.lengths, b.lengths) and np.array_equal(a.start, b.start)
    for i in (0, 17, 49):
        assert np.array_equal(a.clip(i), b.clip(i)) and len(a.clip(i)) == a.lengths[i]
    assert np.array_equal(a.frames(275), 1 + a.lengths // 275)
    shards = distributed.shard_by_cost(a.frames(275), 4)
    assert sorted(i for s in shards for i in s) == list(range(50))


def test_bench_reference_arm_line_shape():
    """bench.py --impl reference prints one JSON line with the contract keys (tiny run)."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, RANK='0', WORLD_SIZE='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0', '--ref-items', '4'], capture_output=True, text=True, env=env, timeout=600)
    line = [l for l in out.stdout.splitlines() if l.startswith('{')][-1]
    d = json.loads(line)
    assert d['impl'] == 'reference' and d['unit'] == 'audio-s/s' and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['e2e']['h2d_bytes_per_step'] == 0
    import bench
    assert d['config'] == bench.workload_config()          # both arms state the same workload
    assert d['metric'] == 'griffin_lim_50it_audio_sec_per_sec' and d['higher_is_better'] is True


def test_gl_sub_batch_split_covers_every_utterance_once():
    """Host-side pipeline plan of griffin_lim_batch: contiguous ranges, geometric growth, remainder merged."""
    from single_speaker_tts_b200._runtime import _split_by_frames
    rng = np.random.default_rng(0)
    for _ in range(20):
        frames = rng.integers(1, 900, size=int(rng.integers(1, 300))).tolist()
        for first, growth in ((6000, 3), (50, 1), (30, 2), (10 ** 9, 3)):
            ranges = _split_by_frames(frames, first, growth)
            assert ranges[0][0] == 0 and ranges[-1][1] == len(frames)
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert all(i1 > i0 for i0, i1 in ranges)
    sizes = [400 * (i1 - i0) for i0, i1 in _split_by_frames([400] * 300, 6000, 3)]
    assert sizes[0] == 6000 and sizes[1] == 18000 and sizes[2] == 54000 and sum(sizes) == 120000


def test_threaded_packing_keeps_order_and_values():
    """_hostio._pack_and_copy (staging for ragged uploads): worker groups + chunked copies must
    reproduce np.concatenate for any mix of sizes, dtypes needing conversion and tiny chunks."""
    import torch
    from single_speaker_tts_b200 import _hostio
    rng = np.random.default_rng(0)
    arrays = [rng.standard_normal(int(n)).astype(np.float64 if i % 3 == 0 else np.float32)
              for i, n in enumerate(rng.integers(1, 50000, size=120))]
    lens = [len(a) for a in arrays]
    total = sum(lens)
    stage, out = torch.empty(total, dtype=torch.float32), torch.empty(total, dtype=torch.float32)
    st = np.concatenate([[0], np.cumsum(lens)])
    views = [stage.numpy()[st[k]:st[k + 1]] for k in range(len(arrays))]
    old = _hostio._CHUNK_BYTES
    try:
        for chunk in (1 << 12, 1 << 18, 1 << 30):
            _hostio._CHUNK_BYTES = chunk
            out.zero_()
            _hostio._pack_and_copy(views, arrays, lens, stage, out, 4)
            assert np.array_equal(out.numpy(), np.concatenate(arrays).astype(np.float32))
    finally:
        _hostio._CHUNK_BYTES = old


def test_threaded_wav_decode_prefetch_and_npz_writer(tmp_path):
    """audio/io.py helpers behind pre_compute_features / collect_decibel_statistics: same order and
    values as the serial loop; .npz files readable with the reference's keys (tacotron/train.py:135-137)."""
    from scipy.io import wavfile
    from single_speaker_tts_b200.audio import io as aio
    rng = np.random.default_rng(3)
    paths = []
    for i in range(7):
        p = str(tmp_path / ('c%d.wav' % i))
        wavfile.write(p, 22050, (rng.standard_normal(500 + 37 * i) * 3000).astype(np.int16))
        paths.append(p)
    serial = [aio.load_wav(p) for p in paths]
    threaded = aio.load_wavs(paths, threads=4)
    assert all(np.array_equal(a[0], b[0]) and a[1] == b[1] for a, b in zip(serial, threaded))
    seen = []
    for chunk, loaded in aio.prefetch_batches(paths, 3, threads=2):
        assert len(chunk) == len(loaded)
        seen.extend(zip(chunk, loaded))
    assert [p for p, _ in seen] == paths
    assert all(np.array_equal(w[0], s[0]) for (_, w), s in zip(seen, serial))
    items = [(str(tmp_path / ('f%d.npz' % i)), {'mel_mag_db': np.full((3, 400), i, np.float32),
                                                 'linear_mag_db': np.full((3, 5125), -i, np.float32)}) for i in range(5)]
    aio.save_npz_many(items, threads=3)
    for i, (p, _) in enumerate(items):
        z = np.load(p)
        assert z['mel_mag_db'].shape == (3, 400) and float(z['linear_mag_db'][0, 0]) == -i
