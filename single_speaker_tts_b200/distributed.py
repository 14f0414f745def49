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
