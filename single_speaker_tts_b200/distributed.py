"""Multi-GPU execution: one process per GPU, work sharded by utterance / clip.

Griffin-Lim and feature extraction have no dependency between utterances, so ranks only need a
balanced partition (by frame count) and no data-path collective.  The one exchange step of the
path is the corpus dB statistics (reference datasets/statistics.py:69-98): every rank reduces
its clips to per-clip 4-vectors on its GPU, the vectors are combined with one all-reduce
(SUM over a zero-initialised (n_clips, 4) table -- each row has exactly one non-zero
contributor, so the sum is an exact gather), and every rank then forms the float64 mean in
listing order exactly like the reference.  The true global extrema are reduced with
MIN / MAX all-reduces alongside.  ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the
CPU tests) is used for the plumbing.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_by_cost(costs, world_size):
    """Greedy longest-first partition of item indices into ``world_size`` shards balancing the
    summed cost (frames).  Deterministic; every shard keeps ascending index order."""
    costs = np.asarray(costs, dtype=np.int64)
    order = np.argsort(-costs, kind='stable')
    loads = np.zeros(world_size, dtype=np.int64)
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(loads))
        shards[r].append(int(i))
        loads[r] += costs[i]
    return [sorted(s) for s in shards]


def _comm_device():
    if dist.is_initialized() and dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def reduce_corpus_statistics(local_rows, local_indices, n_total, group=None):
    """Combine per-clip statistics across ranks.

    local_rows (n_local, 4) float64 belong to global clip indices ``local_indices``.
    Returns ``(mean4, min4, max4, table)``: the reference's corpus value (float64 mean in listing
    order, datasets/statistics.py:86-96), the global per-column extrema, and the full
    (n_total, 4) table -- identical on every rank.
    """
    dev = _comm_device()
    table = torch.zeros((n_total, 4), dtype=torch.float64, device=dev)
    rows = torch.as_tensor(np.asarray(local_rows, dtype=np.float64).reshape(-1, 4), device=dev)
    if rows.shape[0] > 0:
        table[torch.as_tensor(np.asarray(local_indices, dtype=np.int64), device=dev)] = rows
        mn = rows.min(dim=0).values
        mx = rows.max(dim=0).values
    else:
        mn = torch.full((4,), float('inf'), dtype=torch.float64, device=dev)
        mx = torch.full((4,), float('-inf'), dtype=torch.float64, device=dev)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(table, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    table_np = table.cpu().numpy()
    stats = np.zeros(4)
    for row in table_np:           # listing order, float64, like the reference's loop
        stats += row
    stats /= n_total
    return stats, mn.cpu().numpy(), mx.cpu().numpy(), table_np


def corpus_decibel_statistics(local_wavs, local_indices, n_total, sampling_rate, group=None,
                              per_clip_fn=None, batch_clips=512):
    """Sharded ``collect_decibel_statistics``: this rank's decoded clips -> corpus statistics.

    ``per_clip_fn(wavs, sampling_rate) -> (n, 4)`` defaults to the CUDA kernel path
    (:func:`..datasets.statistics.decibel_statistics_batch`); the CPU tests inject the oracle.
    """
    if per_clip_fn is None:
        from .datasets.statistics import decibel_statistics_batch as per_clip_fn
    rows = []
    for s in range(0, len(local_wavs), batch_clips):
        rows.append(np.asarray(per_clip_fn(local_wavs[s:s + batch_clips], sampling_rate)))
    local_rows = np.concatenate(rows, axis=0) if rows else np.zeros((0, 4))
    return reduce_corpus_statistics(local_rows, local_indices, n_total, group=group)


def corpus_pass(local_wavs, local_indices, n_total, sampling_rate, n_fft, hop_length, win_length, n_mels, fmin,
                fmax, reduction=5, group=None, chunk_samples=6 << 20, sink=None, precision='f64'):
    """BASELINE configs[3] on this rank's shard: ``tacotron/dataset_statistics.py`` followed by
    ``tacotron/dataset_precalc_features.py`` with the constants the first one printed.

      1. every clip is packed and uploaded ONCE (chunks of ``chunk_samples`` samples; the upload of chunk
         k + 1 overlaps the statistics kernel of chunk k) and stays on the device;
      2. per-clip ``[min lin, max lin, min mel, max mel]`` dB (datasets/statistics.py:11-66) -> the one
         exchange step of the path: :func:`reduce_corpus_statistics` (all-reduce) -> the float64 mean in
         listing order that the reference prints as ``linear_mag_max_db, linear_ref_db, mel_mag_max_db,
         mel_mag_ref_db`` (tacotron/dataset_statistics.py:35-39);
      3. the feature kernel (datasets/lj_speech.py:124-156) runs on the resident clips with those
         constants; results travel back on a download stream into page-locked blocks while the next
         chunk is transformed, and are handed to ``sink(global_indices, FeatureBatch)`` (numpy views of
         the blocks, no second host copy) two chunks behind the device.

    Returns ``(mean4, n_rows)``; ``n_rows`` counts the feature rows (frames padded to the reduction factor)
    this rank produced."""
    from . import _runtime
    dev = _runtime.require_cuda()
    local_wavs = list(local_wavs)
    ranges = _runtime._split_by_frames([int(w.shape[0]) for w in local_wavs], chunk_samples) if local_wavs else []
    resident, stats = [], []
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()
        copy, back = _runtime._aux_stream(dev, 'h2d'), _runtime._aux_stream(dev, 'd2h')
        nxt = _runtime.upload_clips(local_wavs[ranges[0][0]:ranges[0][1]], dev, copy, 0) if ranges else None
        for k in range(len(ranges)):
            clips = nxt
            if k + 1 < len(ranges):
                nxt = _runtime.upload_clips(local_wavs[ranges[k + 1][0]:ranges[k + 1][1]], dev, copy, k + 1)
            res = _runtime.stft_features_batch(clips, 1024, 256, 1024, sampling_rate=sampling_rate, n_mels=80,
                                               fmin=0, fmax=sampling_rate // 2, want_minmax=True,
                                               precision=precision, keep_on_device=True)
            resident.append(clips)
            stats.append(res)
        main.synchronize()
        local_rows = (torch.cat([r.minmax for r in stats]).cpu().numpy() if stats else np.zeros((0, 4)))
        del stats
    mean4, _, _, _ = reduce_corpus_statistics(local_rows, local_indices, n_total, group=group)
    lin_max, lin_ref, mel_max, mel_ref = mean4          # tacotron/dataset_statistics.py:35-39
    n_rows = 0
    with torch.cuda.device(dev):
        pending = []

        def retire(limit):
            nonlocal n_rows
            while len(pending) > limit:
                (i0, i1), part = pending.pop(0)
                part.done.synchronize()
                n_rows += int(part.row_off[-1])
                if sink is not None:
                    sink(local_indices[i0:i1], part)
                part._keep = None

        for k, clips in enumerate(resident):
            part = _runtime.stft_features_batch(clips, n_fft, hop_length, win_length, sampling_rate=sampling_rate,
                                                n_mels=n_mels, fmin=fmin, fmax=fmax, reduction=reduction,
                                                want_lin=True, want_mel=True,
                                                normalize=(lin_ref, lin_max, mel_ref, mel_max),
                                                precision=precision, _streams=(copy, back), _slot=k)
            pending.append((ranges[k], part))
            retire(2)
        retire(0)
        main.synchronize()
        back.synchronize()
        del resident
        _runtime._feat_plans.reap()
    return mean4, n_rows
