"""world_size-2 gloo test of the one exchange step on the path: the corpus dB statistics
(reference datasets/statistics.py:69-98) reduced across ranks.  The per-clip numbers come from
the oracle here (no GPU); on the B200 box the same reduction runs over NCCL with the CUDA kernel
producing the rows (tests/test_gpu_parity.py, bench.py --workload corpus)."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from oracle import reference_audio as ra
    from single_speaker_tts_b200 import distributed
    from single_speaker_tts_b200.synthetic import make_clips
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    clips = [c[:6000 + 500 * i] for i, c in enumerate(make_clips(7, seed=21, pool=2))]
    frames = [1 + len(c) // 256 for c in clips]
    shards = distributed.shard_by_cost(frames, world)
    mine = shards[rank]

    def per_clip(wavs, sr):
        return np.stack([ra.decibel_statistics(w, sr) for w in wavs])

    mean, mn, mx, table = distributed.corpus_decibel_statistics([clips[i] for i in mine], mine, len(clips),
                                                                22050, per_clip_fn=per_clip, batch_clips=2)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), mean=mean, mn=mn, mx=mx, table=table)
    dist.destroy_process_group()


def test_corpus_statistics_world2(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import reference_audio as ra
    from single_speaker_tts_b200.synthetic import make_clips
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    clips = [c[:6000 + 500 * i] for i, c in enumerate(make_clips(7, seed=21, pool=2))]
    ref = ra.collect_decibel_statistics_from_wavs(clips, 22050)
    rows = np.stack([ra.decibel_statistics(c, 22050) for c in clips])
    r0 = np.load(tmp_path / 'rank0.npz')
    r1 = np.load(tmp_path / 'rank1.npz')
    for r in (r0, r1):
        assert np.array_equal(r['mean'], ref)            # bit-exact: same float64 sum order
        assert np.array_equal(r['table'], rows)
        assert np.array_equal(r['mn'], rows.min(0)) and np.array_equal(r['mx'], rows.max(0))
