"""Host <-> device staging for ragged batches: pooled pinned buffers, threaded packing, chunked
asynchronous copies.

The reference hands numpy arrays in and expects numpy arrays back, so the end-to-end cost of a
call is dominated by host-side copies once the kernels are fast: 460 MB of |S| go in and 124 MB
of waveform come out per 256-utterance Griffin-Lim batch.  This module keeps that path short:

* pinned staging buffers are pooled per thread and reused (``cudaHostAlloc`` of hundreds of MB
  costs more than the Griffin-Lim kernels);
* packing into the staging buffer is done by a few worker threads (``np.copyto`` releases the
  GIL) in chunks -- one task per worker and chunk, each a contiguous run of arrays -- and every
  chunk's H2D copy is issued as soon as it is packed, so packing and DMA overlap;
* inputs that already live in pinned memory (``pinned_empty``) in the device layout skip the
  staging copy altogether and are DMA-ed from where they are;
* results are copied straight into pinned blocks of torch's caching host allocator and handed to
  the caller as numpy views of them (no second host copy).
"""
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

_CHUNK_BYTES = 32 << 20
_N_WORKERS = 8
_tls = threading.local()
_executor = None
_executor_lock = threading.Lock()


def set_io_threads(n):
    """Number of host threads that pack ragged batches into the pinned staging buffers (default 8).  With one
    process per GPU on a shared host use about ``cores // world_size``: the packing is memory-bound and
    more threads than that only fight over the same memory channels."""
    global _N_WORKERS, _executor
    n = max(1, int(n))
    with _executor_lock:
        if n != _N_WORKERS:
            _N_WORKERS = n
            if _executor is not None:
                _executor.shutdown(wait=True)
                _executor = None


def _pool():
    global _executor
    if _executor is None:
        with _executor_lock:
            if _executor is None:
                _executor = ThreadPoolExecutor(max_workers=_N_WORKERS, thread_name_prefix='sstts-io')
    return _executor


class _PinnedBuffer:
    """Grow-only pinned byte buffer with the CUDA event of its last asynchronous use."""

    def __init__(self):
        self.buf = None
        self.event = None

    def get(self, nbytes):
        if self.event is not None:
            self.event.synchronize()        # previous async copies out of / into this buffer
            self.event = None
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        return self.buf

    def mark(self):
        self.event = torch.cuda.Event()
        self.event.record()


def _thread_buffers():
    if not hasattr(_tls, 'up'):
        _tls.up = {}
        _tls.down = [_PinnedBuffer(), _PinnedBuffer()]
    return _tls


def pinned_empty(shape, dtype=np.float32):
    """C-ordered numpy array in page-locked host memory (owned by the returned array).  Inputs that
    live in such arrays (in the layout / dtype the device wants) are uploaded straight from where
    they are -- no staging copy, which otherwise costs about as much host memory bandwidth as the
    upload itself; take ``.T`` of a ``(T, bins)`` array for the reference's ``(bins, T)`` views."""
    t = torch.empty(tuple(int(v) for v in np.atleast_1d(shape)), dtype=torch.from_numpy(np.empty(0, dtype)).dtype,
                    pin_memory=True)
    return t.numpy()


def _direct_sources(srcs, dtype):
    """torch views of ``srcs`` when every array can be DMA-ed as it is (pinned, C-contiguous, device
    dtype), else None."""
    out = []
    for a in srcs:
        if not (isinstance(a, np.ndarray) and a.flags.c_contiguous and a.flags.writeable):
            return None
        try:
            t = torch.from_numpy(a)
        except (TypeError, ValueError):
            return None
        if t.dtype != dtype or not t.is_pinned():
            return None
        out.append(t)
    return out


def _copy_group(pairs):
    for dst, src in pairs:
        np.copyto(dst, src, 'unsafe')


def _pack_and_copy(views, srcs, sizes, stage, out, itemsize_row):
    """Pack ``srcs[k]`` into ``views[k]`` (slices of the pinned ``stage``) with the worker threads and
    issue the H2D copy of every ~_CHUNK_BYTES piece as soon as it is packed.  Every worker gets a
    contiguous run of arrays of about equal bytes -- one task per worker and chunk, not one per
    array (a task hand-over costs as much as copying ~100 KB)."""
    starts = np.concatenate([[0], np.cumsum(sizes)])
    per_chunk = max(1, _CHUNK_BYTES // itemsize_row)
    n = len(srcs)
    if n <= 2 and int(starts[-1]) * itemsize_row <= (16 << 20):
        # the single-utterance path (tacotron/serve.py:39-86): a hand-over to the worker threads costs more
        # than copying a few megabytes in place
        _copy_group(list(zip(views, srcs)))
        out[:starts[-1]].copy_(stage[:starts[-1]], non_blocking=True)
        return
    ex = _pool()
    i = 0
    while i < n:
        j = i
        while j < n and (starts[j + 1] - starts[i] <= per_chunk or j == i):
            j += 1
        total = starts[j] - starts[i]
        target = max(1, -(-int(total) // _N_WORKERS))
        futs, g0, acc = [], i, 0
        for k in range(i, j):
            acc += sizes[k]
            if acc >= target or k == j - 1:
                futs.append(ex.submit(_copy_group, [(views[q], srcs[q]) for q in range(g0, k + 1)]))
                g0, acc = k + 1, 0
        for f in futs:
            f.result()
        out[starts[i]:starts[j]].copy_(stage[starts[i]:starts[j]], non_blocking=True)
        i = j


def upload_rows(blocks, width, dtype, device, slot='a'):
    """Stack 2-D host blocks (rows_i, width) -- any layout / float dtype -- into one device tensor
    (sum rows, width) of ``dtype``.  Packing and H2D are pipelined chunk by chunk."""
    rows = [int(b.shape[0]) for b in blocks]
    total = sum(rows)
    out = torch.empty((total, width), dtype=dtype, device=device)
    if total == 0:
        return out
    direct = _direct_sources(blocks, dtype)
    if direct is not None:
        o = 0
        for t, r in zip(direct, rows):
            out[o:o + r].copy_(t, non_blocking=True)
            o += r
        return out
    itemsize = out.element_size()
    tl = _thread_buffers()
    pb = tl.up.setdefault(slot, _PinnedBuffer())
    stage = pb.get(total * width * itemsize)[:total * width * itemsize].view(dtype).view(total, width)
    stage_np = stage.numpy()
    starts = np.concatenate([[0], np.cumsum(rows)])
    views = [stage_np[starts[k]:starts[k + 1]] for k in range(len(blocks))]
    _pack_and_copy(views, blocks, rows, stage, out, width * itemsize)
    pb.mark()
    return out


def pack_rows(blocks, width, dtype, slot='a'):
    """Host half of :func:`upload_rows` for pipelined calls: pack the blocks into this thread's pinned staging
    buffer ``slot`` with the worker threads and return ``(stage, buffer)`` -- NO CUDA copy is issued.  The
    thread that launches the kernels then calls :func:`issue_rows`.  (Issuing the H2D copies from the packing
    thread while another thread launches kernels costs 8 % of the end-to-end throughput: the two threads
    contend inside the driver, tools/e2e_sweep.py.)  Returns None when the blocks can be DMA-ed from where
    they are (page-locked, device layout): the caller then uses :func:`upload_rows` directly."""
    if _direct_sources(blocks, dtype) is not None:
        return None
    rows = [int(b.shape[0]) for b in blocks]
    total = sum(rows)
    itemsize = torch.empty(0, dtype=dtype).element_size()
    tl = _thread_buffers()
    pb = tl.up.setdefault(slot, _PinnedBuffer())
    stage = pb.get(total * width * itemsize)[:total * width * itemsize].view(dtype).view(total, width)
    if total == 0:
        return stage, pb
    stage_np = stage.numpy()
    starts = np.concatenate([[0], np.cumsum(rows)])
    views = [stage_np[starts[k]:starts[k + 1]] for k in range(len(blocks))]
    ex = _pool()
    target = max(1, -(-int(total) // _N_WORKERS))
    futs, g0, acc = [], 0, 0
    for k in range(len(blocks)):
        acc += rows[k]
        if acc >= target or k == len(blocks) - 1:
            futs.append(ex.submit(_copy_group, [(views[q], blocks[q]) for q in range(g0, k + 1)]))
            g0, acc = k + 1, 0
    for f in futs:
        f.result()
    return stage, pb


def issue_rows(packed, device):
    """Device half: H2D of a :func:`pack_rows` result on the current stream; returns the device tensor."""
    stage, pb = packed
    out = torch.empty(stage.shape, dtype=stage.dtype, device=device)
    if stage.numel():
        out.copy_(stage, non_blocking=True)
    pb.mark()
    return out


def upload_flat(arrays, dtype, device, slot='w'):
    """Concatenate 1-D host arrays into one device tensor (pipelined like :func:`upload_rows`)."""
    lens = [int(a.shape[0]) for a in arrays]
    total = sum(lens)
    out = torch.empty(max(total, 1), dtype=dtype, device=device)
    if total == 0:
        return out
    direct = _direct_sources(arrays, dtype)
    if direct is not None:
        o = 0
        for t, r in zip(direct, lens):
            out[o:o + r].copy_(t, non_blocking=True)
            o += r
        return out
    itemsize = out.element_size()
    tl = _thread_buffers()
    pb = tl.up.setdefault(slot, _PinnedBuffer())
    stage = pb.get(total * itemsize)[:total * itemsize].view(dtype)
    stage_np = stage.numpy()
    starts = np.concatenate([[0], np.cumsum(lens)])
    views = [stage_np[starts[k]:starts[k + 1]] for k in range(len(arrays))]
    _pack_and_copy(views, arrays, lens, stage, out, itemsize)
    pb.mark()
    return out


def download(t):
    """Device tensor -> numpy array of the same shape / dtype, backed by pinned host memory from
    torch's caching host allocator (cheap to re-allocate; the array keeps the block alive and it
    returns to the cache when the caller drops the array).  The copy is asynchronous: the caller
    synchronises the stream before touching the data."""
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    return host.numpy()
