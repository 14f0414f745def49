"""Host-side batch engines over libsstts: packing of ragged batches, device buffers, streams.

PyTorch is used for plumbing only (pinned/device allocations, streams, NCCL); the arithmetic
is done by the CUDA kernels behind the C ABI (``include/sstts.h``).  Nothing here falls back to
the CPU: without the built library or without a CUDA device the calls raise.
"""
import collections
import ctypes
import os
import threading
import time

import numpy as np
import torch

from . import _hostio, _lib
from ._lib import SsttsError

_PRECISIONS = {'f32': _lib.SSTTS_F32, 'f64': _lib.SSTTS_F64,
               'float32': _lib.SSTTS_F32, 'float64': _lib.SSTTS_F64}


def require_cuda(device=None):
    """Return the torch CUDA device to run on; raise if there is none (no CPU fallback)."""
    if not torch.cuda.is_available():
        raise SsttsError('single_speaker_tts_b200 needs a CUDA device (B200, sm_100a); '
                         'there is no CPU fallback.')
    if device is None:
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device(device)


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _make_config(n_fft, win_length, hop_length, precision, sampling_rate=0, n_mels=0, fmin=0.0,
                 fmax=None):
    cfg = _lib.StftConfig()
    cfg.n_fft = int(n_fft)
    cfg.win_length = int(win_length)
    cfg.hop_length = int(hop_length)
    cfg.sampling_rate = int(sampling_rate or 0)
    cfg.n_mels = int(n_mels or 0)
    cfg.mel_fmin = float(fmin or 0.0)
    cfg.mel_fmax = float(fmax) if fmax else 0.0
    cfg.precision = _PRECISIONS[precision]
    return cfg


SUPPORTED_N_FFT = (512, 1024, 2048)


def validate_geometry(n_fft, win_length, hop_length, griffin_lim=False):
    """The kernels cover the geometries the reference uses (n_fft 2048 / 1102 / 275 for the model,
    1024 / 1024 / 256 for the statistics and ``time_stretch``) and their neighbourhood, not everything
    librosa accepts.  Checked here, up front, with a ValueError that names the supported set -- instead
    of a late ``SsttsError`` from plan creation or a launch that does not fit in shared memory:

      * ``n_fft`` in {512, 1024, 2048};  2 <= ``win_length`` <= ``n_fft`` with ``n_fft - win_length`` even
        (librosa centres the window with ``(n_fft - win_length) // 2`` zeros on the left);
      * feature path: 1 <= ``hop_length`` <= ``n_fft``;
      * Griffin-Lim: 1 <= ``hop_length`` <= ``win_length`` and ``ceil(win_length / hop_length)`` <= 5
        (at most five frames overlap one sample -- the gather of the overlap-add is unrolled for that).
    """
    n_fft, win_length, hop_length = int(n_fft), int(win_length), int(hop_length)
    if n_fft not in SUPPORTED_N_FFT:
        raise ValueError('unsupported n_fft={} (supported: 512, 1024, 2048)'.format(n_fft))
    if not 2 <= win_length <= n_fft:
        raise ValueError('unsupported win_length={} (need 2 <= win_length <= n_fft={})'.format(win_length, n_fft))
    if (n_fft - win_length) % 2:
        raise ValueError('unsupported win_length={}: n_fft - win_length must be even'.format(win_length))
    if hop_length < 1:
        raise ValueError('hop_length must be >= 1, got {}'.format(hop_length))
    if griffin_lim:
        if hop_length > win_length or -(-win_length // hop_length) > 5:
            raise ValueError('unsupported hop_length={} for Griffin-Lim with win_length={}: need '
                             'win_length / 5 <= hop_length <= win_length'.format(hop_length, win_length))
    elif hop_length > n_fft:
        raise ValueError('unsupported hop_length={} (need hop_length <= n_fft={})'.format(hop_length, n_fft))


class _Plan:
    """Owns one native plan handle; destroyed with the object."""

    def __init__(self, handle, destroy):
        self.handle = handle
        self._destroy = destroy

    def __del__(self):
        try:
            if self.handle:
                self._destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _PlanCache:
    """Small thread-safe LRU of native plans keyed by geometry + batch shape + device.

    Destroying a plan frees device memory (``cudaFree`` waits for the device), so evicted plans are
    parked and only released by :meth:`reap`, which the batch entry points call after their final
    synchronisation -- an eviction in the middle of a sub-batch pipeline would stall it."""

    def __init__(self, capacity=2048):
        self._cap = capacity
        self._d = collections.OrderedDict()
        self._dead = []
        self._lock = threading.Lock()

    def get(self, key, factory):
        with self._lock:
            plan = self._d.get(key)
            if plan is not None:
                self._d.move_to_end(key)
                return plan
        plan = factory()
        with self._lock:
            self._d[key] = plan
            while len(self._d) > self._cap:
                self._dead.append(self._d.popitem(last=False)[1])
        return plan

    def reap(self):
        with self._lock:
            dead, self._dead = self._dead, []
        del dead


_gl_plans = _PlanCache()
_feat_plans = _PlanCache()


def _offsets(counts):
    off = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(np.asarray(counts, dtype=np.int64), out=off[1:])
    return off


# ----------------------------------------------------------------------------------------------
# Griffin-Lim
# ----------------------------------------------------------------------------------------------
# Large batches are split into equal sub-batches whose packing + upload overlaps the previous
# sub-batch's iterations (copy streams + events); results do not depend on the split.  Packing into
# pinned memory (~25-35 GB/s with the staging threads) is faster than the device iterates, so only the
# first sub-batch's upload is exposed; sub-batches alternate between two compute streams so that the
# tail of one launch (the last, partly filled wave of tiles) is filled by the next sub-batch's CTAs.
# tools/e2e_probe.py measures the trade-off (more sub-batches: shorter exposed upload, more launches).
_GL_CHUNK_FRAMES = 10000
_GL_CHUNK_GROWTH = 1
_GL_CHUNK_HEAD = None
_GL_CHUNK_CAP = None
_trace = None     # tools/e2e_trace.py sets this to a list: (what, sub-batch, t0, t1) host timestamps of a call
_aux_streams = threading.local()


def _aux_stream(dev, name):
    d = getattr(_aux_streams, 'd', None)
    if d is None:
        d = _aux_streams.d = {}
    if (dev.index, name) not in d:
        d[(dev.index, name)] = torch.cuda.Stream(device=dev)
    return d[(dev.index, name)]


def _uploader():
    """Single helper thread of the calling thread (kept for the thread's lifetime, so its pinned staging
    buffers persist): packs and uploads the sub-batches of a pipelined call two ahead of the launches."""
    ex = getattr(_aux_streams, 'uploader', None)
    if ex is None:
        from concurrent.futures import ThreadPoolExecutor
        ex = _aux_streams.uploader = ThreadPoolExecutor(max_workers=1, thread_name_prefix='sstts-upload')
    return ex


def _split_by_frames(frames, first, growth=1, head=None, cap=None):
    """Contiguous index ranges of roughly first, first * growth, first * growth ** 2, ... frames (at most
    ``cap`` each); a short remainder is merged into the last range.  ``head``: size of the very first range
    only (a small first sub-batch gets the device started while the rest is still being packed)."""
    ranges, i0, acc, limit = [], 0, 0, (head or first)
    for i, t in enumerate(frames):
        acc += t
        if acc >= limit:
            ranges.append((i0, i + 1))
            i0, acc, limit = i + 1, 0, (first if (head and len(ranges) == 1) else limit * growth)
            if cap:
                limit = min(limit, cap)
    if i0 < len(frames):
        if ranges and acc * 4 < limit // max(growth, 1):
            ranges[-1] = (ranges[-1][0], len(frames))
        else:
            ranges.append((i0, len(frames)))
    return ranges


def griffin_lim_batch(mags, win_length, hop_length, n_fft, n_iter, angles=None, seed=None,
                      precision='f32', return_mse=False, device=None, denormalize=None, normalize_peak=False,
                      uniform=None):
    """Griffin-Lim for a ragged batch (reference: audio/synthesis.py:43-125, one call per item).

    mags   : list of (1 + n_fft/2, T_i) magnitude spectrograms (any float dtype / layout).
    angles : optional list of complex (1 + n_fft/2, T_i) initial unit phasors (the reference's
             ``np.exp(2j*pi*np.random.rand(...))``); if None they are generated on the device from
             ``seed`` (seed=None draws one 63-bit seed from numpy's global RNG, so
             ``np.random.seed`` still makes a run reproducible).
    uniform : optional list of float64 ``(1 + n_fft/2, T_i)`` arrays of U[0, 1) numbers, the reference's
             ``np.random.rand(*spectrogram.shape)`` (audio/synthesis.py:85): uploaded as drawn and turned
             into the initial phasors ``exp(2j * pi * u)`` on the device (the per-item functions use this
             to keep numpy's global random stream without paying for the complex exponential and the
             transposing copy on the host).
    denormalize : None, or ``(ref_db, max_db, power)``: ``mags`` then holds the model's normalised
             outputs in its own orientation ``(T_i, 1 + n_fft/2)`` and the glue of
             tacotron/inference.py:94-101,175 (inv_normalize_decibel -> decibel_to_magnitude ->
             ** power) runs fused on the device before the first iteration.
    normalize_peak : divide every waveform by its peak on the device -- the ``norm=True`` of the
             reference's ``save_wav`` (audio/io.py:33-53), which tacotron/inference.py:199 applies.
    Returns (list of float32 waveforms of length hop*(T_i-1), list of mse floats or None).

    Large batches are processed as a pipeline of sub-batches: while sub-batch k iterates on the
    compute stream, the host packs sub-batch k + 1 into pinned memory and its H2D copy runs on a
    copy stream.
    """
    validate_geometry(n_fft, win_length, hop_length, griffin_lim=True)
    lib = _lib.load()
    dev = require_cuda(device)
    n_bins = 1 + n_fft // 2
    n = len(mags)
    if n == 0:
        return [], ([] if return_mse else None)
    if denormalize is not None:
        mags = [np.asarray(m).T for m in mags]      # (bins, T) views of the (T, bins) model outputs
    frames = []
    for m in mags:
        if m.ndim != 2 or m.shape[0] != n_bins:
            raise ValueError('spectrogram must have shape ({}, T), got {}'.format(n_bins, m.shape))
        if m.shape[1] < 1:
            raise ValueError('spectrogram needs at least one frame')
        frames.append(m.shape[1])
    if return_mse and n_iter < 1:
        raise ValueError('mse needs n_iter >= 1')
    if angles is not None or uniform is not None:
        given = angles if angles is not None else uniform
        if len(given) != n:
            raise ValueError('need one initial phase array per spectrogram')
        for a, m in zip(given, mags):
            if a.shape != m.shape:
                raise ValueError('initial phase shape {} != spectrogram shape {}'.format(a.shape, m.shape))
    elif seed is None:
        seed = int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))
    cfg = _make_config(n_fft, win_length, hop_length, precision)
    i64p = ctypes.POINTER(ctypes.c_int64)
    ranges = _split_by_frames(frames, _GL_CHUNK_FRAMES, _GL_CHUNK_GROWTH, head=_GL_CHUNK_HEAD, cap=_GL_CHUNK_CAP)
    frame_base = _offsets(frames)

    def get_plan(i0, i1):
        fo = _offsets(frames[i0:i1])
        # the library picks the n_fft 1024 kernel variant at plan creation (SSTTS_GL_NATIVE1024, an A/B switch)
        key = (dev.index, n_fft, win_length, hop_length, cfg.precision, fo.tobytes(),
               os.environ.get('SSTTS_GL_NATIVE1024') if n_fft == 1024 else None)

        def factory():
            h = ctypes.c_void_p()
            with torch.cuda.device(dev):
                _lib.check(lib.sstts_gl_plan_create(ctypes.byref(cfg), i1 - i0, fo.ctypes.data_as(i64p),
                                                    ctypes.byref(h)))
            return _Plan(h, lib.sstts_gl_plan_destroy)

        return _gl_plans.get(key, factory), fo

    # All work runs on streams private to the calling thread (inputs and outputs are host arrays, so
    # nothing has to be ordered against the caller's current stream): concurrent callers -- the six
    # synthesis threads of tacotron/serve.py:69-72 -- then overlap on the device instead of queueing
    # behind each other on the default stream.
    with torch.cuda.device(dev), torch.cuda.stream(_aux_stream(dev, 'gl0')):
        main = torch.cuda.current_stream()
        piped = len(ranges) > 1
        copy = _aux_stream(dev, 'h2d') if piped else main     # uploads
        back = _aux_stream(dev, 'd2h') if piped else main     # result downloads (other DMA direction)
        compute = [main, _aux_stream(dev, 'gl1')] if piped else [main]

        keep = []

        def pack(k):
            """Helper thread, host work only: |S| of sub-batch k into a pinned staging buffer (None when the
            arrays are page-locked already and can be DMA-ed from where they are)."""
            i0, i1 = ranges[k]
            t_up = time.perf_counter() if _trace is not None else 0.0
            packed = _hostio.pack_rows([m.T for m in mags[i0:i1]], n_bins, torch.float32, slot='mag%d' % (k % 3))
            if _trace is not None:
                _trace.append(('pack', k, t_up, time.perf_counter()))
            return packed

        def upload(k, packed=None):
            """H2D of sub-batch k on the copy stream (packing it first unless a helper thread did); returns
            device tensors and an event.  Always runs on the launching thread."""
            i0, i1 = ranges[k]
            t_up = time.perf_counter() if _trace is not None else 0.0
            with torch.cuda.device(dev), torch.cuda.stream(copy):
                if packed is not None:
                    mag_dev = _hostio.issue_rows(packed, dev)
                else:
                    mag_dev = _hostio.upload_rows([m.T for m in mags[i0:i1]], n_bins, torch.float32, dev,
                                                  slot='mag%d' % (k % 3))
                ph_dev = None
                if angles is not None:
                    ph_dev = torch.view_as_real(_hostio.upload_rows(
                        [np.asarray(a).T for a in angles[i0:i1]], n_bins, torch.complex64, dev,
                        slot='phase%d' % (k % 3)))
                elif uniform is not None:
                    u_dev = _hostio.upload_flat([np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
                                                 for u in uniform[i0:i1]], torch.float64, dev,
                                                slot='uni%d' % (k % 3))
                    ph_dev = torch.empty((int(sum(frames[i0:i1])), n_bins, 2), dtype=torch.float32, device=dev)
                    o = 0
                    for t in frames[i0:i1]:
                        _lib.check(lib.sstts_phase_from_uniform(
                            ctypes.c_void_p(u_dev.data_ptr() + 8 * o * n_bins), t, n_bins,
                            ctypes.c_void_p(ph_dev.data_ptr() + 8 * o * n_bins), _stream_ptr()))
                        o += t
                    keep.append(u_dev)
                ev = torch.cuda.Event()
                ev.record(copy)
            if _trace is not None:
                _trace.append(('upload', k, t_up, time.perf_counter()))
            return mag_dev, ph_dev, ev

        flag_dev = torch.zeros(1, dtype=torch.int32, device=dev) if denormalize is not None else None
        if piped and flag_dev is not None:
            ready = torch.cuda.Event()
            ready.record(main)
            compute[1].wait_event(ready)
        # Device tensors cross streams here (uploaded on one, consumed on another, downloaded on a third).
        # Instead of record_stream -- which makes the caching allocator defer their reuse behind events
        # and fall back to cudaMalloc in the next call -- every tensor is kept alive in `keep` until all
        # streams have been synchronised at the end of the call: then freeing is hazard-free and the next
        # call finds exactly the blocks it needs in the allocator's per-stream caches.
        keep.append(flag_dev)
        outs = []
        # Pipelined calls: a helper thread packs (pageable -> pinned) sub-batches k + 1 and k + 2 while this thread
        # issues the H2D copy and the launches of sub-batch k, so the host-side packing never sits between two
        # launch sequences.  The helper issues no copy and no launch: copies issued from a second thread while this
        # one launches kernels cost 8 % end to end (36.5 -> 33.7 ms per 256-utterance call, tools/e2e_sweep.py).
        if piped:
            ex = _uploader()
            futs = collections.deque(ex.submit(pack, k) for k in range(min(2, len(ranges))))
        else:
            nxt = upload(0)
        for k, (i0, i1) in enumerate(ranges):
            t_w = time.perf_counter() if _trace is not None else 0.0
            if piped:
                packed = futs.popleft().result()
                if k + 2 < len(ranges):
                    futs.append(ex.submit(pack, k + 2))
                nxt = upload(k, packed)
            if _trace is not None:
                _trace.append(('wait', k, t_w, time.perf_counter()))
            mag_dev, phase_dev, ev = nxt
            plan, fo = get_plan(i0, i1)
            tf = int(fo[-1])
            comp = compute[k % len(compute)]
            with torch.cuda.stream(comp):
                comp.wait_event(ev)
                if _trace is not None:
                    g0 = torch.cuda.Event(enable_timing=True)
                    g0.record(comp)
                if denormalize is not None:
                    ref_db, max_db, power = [float(v) for v in denormalize]
                    _lib.check(lib.sstts_denormalize_magnitude(_ptr(mag_dev), tf * n_bins, ref_db, max_db, power,
                                                               _ptr(mag_dev), _ptr(flag_dev), _stream_ptr()))
                ts = int(lib.sstts_gl_total_samples(plan.handle))
                so = np.ctypeslib.as_array(lib.sstts_gl_sample_offsets(plan.handle), shape=(i1 - i0 + 1,)).copy()
                ws = torch.empty(int(lib.sstts_gl_workspace_bytes(plan.handle)), dtype=torch.uint8, device=dev)
                wav_dev = torch.empty(max(ts, 1), dtype=torch.float32, device=dev)
                mse_dev = torch.zeros(tf, dtype=torch.float64, device=dev) if return_mse else None
                if phase_dev is not None:
                    _lib.check(lib.sstts_griffin_lim(plan.handle, _ptr(mag_dev), _ptr(phase_dev), int(n_iter),
                                                     _ptr(ws), _ptr(wav_dev), _ptr(mse_dev), _stream_ptr()))
                else:
                    # the first launch draws the phase itself: element index = global frame row * bins + bin,
                    # so the result does not depend on how the batch was split
                    _lib.check(lib.sstts_griffin_lim_seeded(plan.handle, _ptr(mag_dev),
                                                            ctypes.c_uint64(int(seed) & (2 ** 64 - 1)),
                                                            int(frame_base[i0]) * n_bins, int(n_iter), _ptr(ws),
                                                            _ptr(wav_dev), _ptr(mse_dev), _stream_ptr()))
                if normalize_peak:
                    _lib.check(lib.sstts_peak_normalize(plan.handle, _ptr(wav_dev), _stream_ptr()))
                if _trace is not None:
                    g1 = torch.cuda.Event(enable_timing=True)
                    g1.record(comp)
                    _trace.append(('gpu', k, g0, g1))
                # results go back on their own stream so that the next iterations start at once
                if piped:
                    done = torch.cuda.Event()
                    done.record(comp)
                    back.wait_event(done)
                keep.append((mag_dev, phase_dev, ws, wav_dev, mse_dev, plan))   # the plan outlives its launches
            with torch.cuda.stream(back):
                outs.append((_hostio.download(wav_dev), _hostio.download(mse_dev) if return_mse else None, so, fo))
            if _trace is not None:
                _trace.append(('enqueue', k, t_w, time.perf_counter()))
        if piped:
            compute[1].synchronize()
        flag = _hostio.download(flag_dev) if flag_dev is not None else None
        main.synchronize()
        if piped:
            back.synchronize()
            copy.synchronize()
        del keep, nxt, mag_dev, phase_dev, ws, wav_dev, mse_dev, plan
        _gl_plans.reap()
    if flag is not None and int(flag[0]) != 0:
        # same error as the reference's decibel_to_magnitude (audio/conversion.py:47-49)
        raise AssertionError('"conversion.decibel_to_magnitude" was asked to convert a dB value '
                             'smaller -100 dB.')
    wavs, mses = [], ([] if return_mse else None)
    for (wav_np, mf, so, fo), (i0, i1) in zip(outs, ranges):
        for j in range(i1 - i0):
            wavs.append(wav_np[so[j]:so[j + 1]])
            if return_mse:
                t = frames[i0 + j]
                mses.append(None if t < 2 else float(mf[fo[j]:fo[j + 1]].sum() / (n_bins * t)))
    return wavs, mses


# ----------------------------------------------------------------------------------------------
# STFT features
# ----------------------------------------------------------------------------------------------
class FeatureBatch:
    """Result of :func:`stft_features_batch` (host arrays, split per clip on access)."""

    def __init__(self, n_clips, frames, row_off, reduction):
        self.n_clips = n_clips
        self.frames = frames
        self.row_off = row_off
        self.reduction = reduction
        self.spec = None      # (rows, bins) complex64
        self.lin_db = None    # (rows, bins) float32
        self.mel_db = None    # (rows, n_mels) float32
        self.mel_raw = None   # (rows, n_mels) float64
        self.minmax = None    # (n_clips, 4) float64
        self.mel_basis = None
        self.trim_bounds = None  # (n_clips, 2) int64 (start, end) when trimming was requested
        self.done = None         # pipelined call: CUDA event after the last download of this batch
        self._keep = None        # device tensors of a pipelined call, released after its final sync

    def rows(self, arr, i, padded=False):
        a, b = int(self.row_off[i]), int(self.row_off[i + 1])
        if not padded:
            b = a + self.frames[i]
        return arr[a:b]


class DeviceClips:
    """Decoded clips resident on the device: one packed float32 buffer and the per-clip lengths.
    Produced by :func:`upload_clips`; :func:`stft_features_batch` accepts it in place of the host list,
    so a corpus pass uploads every clip once and runs the statistics and the pre-calculation kernels on
    the same buffer (datasets/statistics.py:89 and datasets/lj_speech.py:114 each decode the file again
    in the reference)."""

    def __init__(self, wav_dev, lens, keep=None, ready=None):
        self.wav_dev = wav_dev
        self.lens = [int(v) for v in lens]
        self.keep = keep          # e.g. the int16 upload the float buffer was converted from
        self.ready = ready        # CUDA event: the upload / conversion has finished

    def __len__(self):
        return len(self.lens)


def upload_clips(wavs, device=None, stream=None, slot=0):
    """Pack + upload a list of 1-D clips (float32, or all int16 PCM) once; returns :class:`DeviceClips`.
    ``stream``: side stream to run the copy (and the PCM conversion) on; the returned object carries the
    event consumers have to wait for."""
    lib = _lib.load()
    dev = require_cuda(device)
    wavs = list(wavs)
    pcm16 = all(getattr(w, 'dtype', None) == np.int16 for w in wavs)
    if not pcm16 and any(getattr(w, 'dtype', None) == np.int16 for w in wavs):
        wavs = [(w.astype(np.float32) / 32768.0) if w.dtype == np.int16 else w for w in wavs]
    lens = [int(w.shape[0]) for w in wavs]
    with torch.cuda.device(dev):
        st = stream if stream is not None else torch.cuda.current_stream()
        with torch.cuda.stream(st):
            up = _hostio.upload_flat(wavs, torch.int16 if pcm16 else torch.float32, dev, slot='clips%d' % (slot & 1))
            keep = None
            if pcm16:
                keep = up
                up = torch.empty(keep.shape, dtype=torch.float32, device=dev)
                _lib.check(lib.sstts_pcm16_to_float(_ptr(keep), int(sum(lens)), _ptr(up), _stream_ptr()))
            ev = torch.cuda.Event()
            ev.record(st)
    return DeviceClips(up, lens, keep=keep, ready=ev)


def stft_features_batch(wavs, n_fft, hop_length, win_length, sampling_rate=None, n_mels=0, fmin=0.0,
                        fmax=None, reduction=1, want_spec=False, want_lin=False, want_mel=False,
                        want_mel_raw=False, want_minmax=False, normalize=None, power=1.0,
                        precision='f64', device=None, keep_on_device=False, trim=None, force_generic=False,
                        _streams=None, _slot=0):
    """Batched STFT -> |.| -> linear / mel -> dB -> (0,1) pipeline on the GPU.

    normalize: None (raw dB) or (lin_ref_db, lin_max_db, mel_ref_db, mel_max_db) as in
    audio/conversion.py:56-78.  Outputs are frame-major (rows, bins); with ``reduction`` r > 1 every
    clip's rows are zero-padded to a multiple of r (datasets/dataset_helper.py:357-401).
    trim: None, or ``(top_db, frame_length, hop_length)``: silence-trim every clip on the device
    first (librosa.effects.trim as called at datasets/lj_speech.py:119); the features are computed
    on the trimmed part of the same upload and ``result.trim_bounds`` holds (start, end) per clip.
    force_generic: run the kernel's generic mode even where the fused dB-feature mode applies (lin +
    mel dB only, n_fft 2048, power 1) -- for validation; the two agree to float32 rounding.
    _streams / _slot: used by :func:`stft_features_parts` -- (upload, download) side streams; the call
    then returns without synchronising and the caller synchronises both streams.
    """
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    validate_geometry(n_fft, win_length, hop_length)
    lib = _lib.load()
    dev = require_cuda(device)
    n_bins = 1 + n_fft // 2
    n = len(wavs)
    if n == 0:
        raise ValueError('empty batch')
    lens = []
    resident = wavs if isinstance(wavs, DeviceClips) else None
    # clips may be handed over as raw 16-bit PCM (all of them): uploaded as 2-byte samples and converted
    # on the device exactly like load_wav does on the host (int16 / 32768)
    pcm16 = resident is None and all(getattr(w, 'dtype', None) == np.int16 for w in wavs)
    if resident is not None:
        lens = list(resident.lens)
        if min(lens) < 1:
            raise ValueError('clip is empty')
    else:
        if not pcm16 and any(getattr(w, 'dtype', None) == np.int16 for w in wavs):
            wavs = [(w.astype(np.float32) / 32768.0) if w.dtype == np.int16 else w for w in wavs]   # mixed list
        for w in wavs:
            if w.ndim != 1:
                raise ValueError('Invalid shape for monophonic audio: ndim={:d}'.format(w.ndim))
            if w.shape[0] < 1:
                raise ValueError('clip is empty')
            lens.append(w.shape[0])
    sample_off = _offsets(lens)
    need_mel = want_mel or want_mel_raw or want_minmax
    cfg = _make_config(n_fft, win_length, hop_length, precision, sampling_rate if need_mel else 0,
                       n_mels if need_mel else 0, fmin, fmax)
    i64p = ctypes.POINTER(ctypes.c_int64)
    clip_start = np.ascontiguousarray(sample_off[:-1])
    clip_len = np.asarray(lens, dtype=np.int64)
    trim_bounds = None
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()
        if resident is not None:
            wav_dev = resident.wav_dev
            if resident.ready is not None:
                main.wait_event(resident.ready)
        elif _streams is not None:
            with torch.cuda.stream(_streams[0]):
                wav_dev = _hostio.upload_flat(wavs, torch.int16 if pcm16 else torch.float32, dev,
                                              slot='wav%d' % (_slot & 1))
                up = torch.cuda.Event()
                up.record(_streams[0])
            main.wait_event(up)
        else:
            wav_dev = _hostio.upload_flat(wavs, torch.int16 if pcm16 else torch.float32, dev, slot='wav')
        pcm_dev = None
        if pcm16:
            pcm_dev = wav_dev                # kept alive until the conversion has run
            wav_dev = torch.empty(pcm_dev.shape, dtype=torch.float32, device=dev)
            _lib.check(lib.sstts_pcm16_to_float(_ptr(pcm_dev), int(sample_off[-1]), _ptr(wav_dev), _stream_ptr()))
        if trim is not None:
            top_db, t_frame, t_hop = trim
            start_dev = torch.from_numpy(clip_start).to(dev)
            len_dev = torch.from_numpy(clip_len).to(dev)
            bounds_dev = torch.empty((n, 2), dtype=torch.int64, device=dev)
            _lib.check(lib.sstts_trim_bounds(_ptr(wav_dev), n, _ptr(start_dev), _ptr(len_dev), float(top_db),
                                             int(t_frame), int(t_hop), _ptr(bounds_dev), _stream_ptr()))
            trim_bounds = bounds_dev.cpu().numpy()
            clip_start = clip_start + trim_bounds[:, 0]
            clip_len = trim_bounds[:, 1] - trim_bounds[:, 0]
            if (clip_len < 1).any():
                raise ValueError('clip {} is empty after silence trimming'.format(int(np.argmax(clip_len < 1))))
    key = (dev.index, n_fft, win_length, hop_length, cfg.precision, cfg.sampling_rate, cfg.n_mels,
           cfg.mel_fmin, cfg.mel_fmax, int(reduction), clip_start.tobytes(), clip_len.tobytes())

    def factory():
        h = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.sstts_feat_plan_create_ranges(ctypes.byref(cfg), n, clip_start.ctypes.data_as(i64p),
                                                         clip_len.ctypes.data_as(i64p), int(reduction),
                                                         ctypes.byref(h)))
        return _Plan(h, lib.sstts_feat_plan_destroy)

    plan = _feat_plans.get(key, factory)
    rows = int(lib.sstts_feat_total_rows(plan.handle))
    frame_off = np.ctypeslib.as_array(lib.sstts_feat_frame_offsets(plan.handle), shape=(n + 1,)).copy()
    row_off = np.ctypeslib.as_array(lib.sstts_feat_row_offsets(plan.handle), shape=(n + 1,)).copy()
    res = FeatureBatch(n, [int(frame_off[i + 1] - frame_off[i]) for i in range(n)], row_off, reduction)
    res.trim_bounds = trim_bounds
    if want_mel_raw and cfg.n_mels > 0:          # the filterbank itself, for callers of mel_scale_spectrogram
        res.mel_basis = np.ctypeslib.as_array(lib.sstts_feat_mel_basis(plan.handle),
                                              shape=(cfg.n_mels, n_bins)).copy()

    with torch.cuda.device(dev):
        out = _lib.FeatOutputs()
        spec_dev = torch.empty((rows, n_bins, 2), dtype=torch.float32, device=dev) if want_spec else None
        lin_dev = torch.empty((rows, n_bins), dtype=torch.float32, device=dev) if want_lin else None
        mel_dev = torch.empty((rows, cfg.n_mels), dtype=torch.float32, device=dev) if want_mel else None
        raw_dev = torch.empty((rows, cfg.n_mels), dtype=torch.float64, device=dev) if want_mel_raw else None
        mm_dev = torch.empty((n, 4), dtype=torch.float64, device=dev) if want_minmax else None
        out.spec_dev = spec_dev.data_ptr() if want_spec else None
        out.lin_db_dev = lin_dev.data_ptr() if want_lin else None
        out.mel_db_dev = mel_dev.data_ptr() if want_mel else None
        out.mel_raw_dev = raw_dev.data_ptr() if want_mel_raw else None
        out.minmax_dev = mm_dev.data_ptr() if want_minmax else None
        out.normalize = 1 if normalize is not None else 0
        if normalize is not None:
            out.lin_ref_db, out.lin_max_db, out.mel_ref_db, out.mel_max_db = [float(v) for v in normalize]
        out.mel_power = float(power)
        out.force_generic = 1 if force_generic else 0
        _lib.check(lib.sstts_stft_features(plan.handle, _ptr(wav_dev), ctypes.byref(out), _stream_ptr()))
        if keep_on_device:
            res.spec = torch.view_as_complex(spec_dev) if want_spec else None
            res.lin_db, res.mel_db, res.mel_raw, res.minmax = lin_dev, mel_dev, raw_dev, mm_dev
            res._keep = (wav_dev, pcm_dev, plan)     # the launch may still be running: the plan (its tile
            return res                               # tables) and the input live as long as the result

        back = main
        if _streams is not None:
            back = _streams[1]
            done = torch.cuda.Event()
            done.record(main)
            back.wait_event(done)
            # kept alive until stft_features_parts has synchronised the streams (no record_stream: see
            # griffin_lim_batch)
            res._keep = (wav_dev, pcm_dev, spec_dev, lin_dev, mel_dev, raw_dev, mm_dev, plan)
        with torch.cuda.stream(back):
            res.spec = _hostio.download(spec_dev).view(np.complex64).reshape(rows, n_bins) if want_spec else None
            res.lin_db = _hostio.download(lin_dev) if want_lin else None
            res.mel_db = _hostio.download(mel_dev) if want_mel else None
            res.mel_raw = _hostio.download(raw_dev) if want_mel_raw else None
            res.minmax = _hostio.download(mm_dev) if want_minmax else None
            if _streams is not None:
                res.done = torch.cuda.Event()
                res.done.record(back)
        if _streams is None:
            main.synchronize()
            del plan
            _feat_plans.reap()
    return res


# Sub-batches of about this many samples: while one is being transformed, the next one is packed
# and uploaded and the previous one's features travel back (H2D and D2H use different DMA engines).
_FEAT_CHUNK_SAMPLES = 6 << 20
_FEAT_CHUNK_HEAD = None          # size of the first sub-batch only (tools/feat_chunk_sweep.py: no effect, 11.5 ms either way)


def stft_features_parts(wavs, *args, **kwargs):
    """:func:`stft_features_batch` for large batches as a pipeline of sub-batches: returns a list of
    ``(i0, i1, FeatureBatch)`` covering ``wavs[i0:i1]``.  Per-clip results are identical to one big
    call (clips are independent); the output (4.4 KB per frame) dominates the PCIe traffic, so the
    gain is the overlap of uploads, kernels and downloads."""
    wavs = list(wavs)
    dev = require_cuda(kwargs.get('device'))
    ranges = _split_by_frames([int(w.shape[0]) for w in wavs], _FEAT_CHUNK_SAMPLES, head=_FEAT_CHUNK_HEAD)
    if len(ranges) < 2 or kwargs.get('keep_on_device'):
        return [(0, len(wavs), stft_features_batch(wavs, *args, **kwargs))]
    with torch.cuda.device(dev):
        streams = (_aux_stream(dev, 'h2d'), _aux_stream(dev, 'd2h'))
        parts = []
        for k, (i0, i1) in enumerate(ranges):
            parts.append((i0, i1, stft_features_batch(wavs[i0:i1], *args, _streams=streams, _slot=k, **kwargs)))
        torch.cuda.current_stream().synchronize()
        streams[1].synchronize()
        streams[0].synchronize()
        for _, _, part in parts:
            part._keep = None
        _feat_plans.reap()
    return parts
