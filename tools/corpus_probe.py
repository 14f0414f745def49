"""Phase timing of distributed.corpus_pass on one GPU (host wall clock, device synchronised at phase ends).

    python tools/corpus_probe.py [n_clips]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from single_speaker_tts_b200 import _hostio, _runtime, distributed   # noqa: E402
from single_speaker_tts_b200.synthetic import ClipPlan                # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 13100
plan = ClipPlan(n, seed=3, kind='ljspeech', pool=32)
idx = list(range(n))
wavs = plan.clips(idx)
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
total_samples = sum(len(w) for w in wavs)
print('clips', n, 'samples', total_samples, 'audio_s', total_samples / 22050)
import single_speaker_tts_b200 as pkg
IO = [int(v) for v in os.environ.get('SSTTS_PROBE_IO_THREADS', '8,8,8').split(',')]     # one pass per entry
for rep in range(len(IO)):
    pkg.set_io_threads(IO[rep])
    t0 = time.perf_counter()
    mean4, rows = distributed.corpus_pass(wavs, idx, n, 22050, 2048, 275, 1102, 80, 0, 8000, reduction=5)
    torch.cuda.synchronize()
    print('corpus_pass %.1f ms  rows %d  io threads %d' % ((time.perf_counter() - t0) * 1000, rows, IO[rep]))

# pieces
copy = _runtime._aux_stream(dev, 'h2d')
ranges = _runtime._split_by_frames([len(w) for w in wavs], 6 << 20)
for rep in range(2):
    t0 = time.perf_counter()
    res = []
    for k, (i0, i1) in enumerate(ranges):
        res.append(_runtime.upload_clips(wavs[i0:i1], dev, copy, k))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print('upload only: %.1f ms (%.1f GB/s)' % ((t1 - t0) * 1000, total_samples * 4 / (t1 - t0) / 1e9))
    t0 = time.perf_counter()
    st = [_runtime.stft_features_batch(c, 1024, 256, 1024, sampling_rate=22050, n_mels=80, fmin=0, fmax=11025, want_minmax=True,
                                       keep_on_device=True) for c in res]
    torch.cuda.synchronize()
    print('statistics kernels on resident clips: %.1f ms' % ((time.perf_counter() - t0) * 1000))
    t0 = time.perf_counter()
    outs = [_runtime.stft_features_batch(c, 2048, 275, 1102, sampling_rate=22050, n_mels=80, fmin=0, fmax=8000, reduction=5,
                                         want_lin=True, want_mel=True, normalize=(35.66, 100.0, 6.02, 99.89), keep_on_device=True)
            for c in res[:8]]
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print('feature kernels, 8 of %d chunks, results kept on device: %.1f ms' % (len(res), (t1 - t0) * 1000))
    nbytes = sum(o.lin_db.numel() * 4 + o.mel_db.numel() * 4 for o in outs)
    t0 = time.perf_counter()
    hosts = []
    for o in outs:
        hosts.append(_hostio.download(o.lin_db)); hosts.append(_hostio.download(o.mel_db))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print('download of those: %.1f ms (%.1f GB/s)' % ((t1 - t0) * 1000, nbytes / (t1 - t0) / 1e9))
    del res, st, outs, hosts
