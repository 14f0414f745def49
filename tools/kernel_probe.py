"""Device-resident timings of the feature / statistics kernel instances on the 256-clip benchmark set
(CUDA events, 20 launches each after 3 warm-ups) -- the quick A/B loop for kernel work.

    [SSTTS_LIB=path/to/lib.so] python tools/kernel_probe.py
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from single_speaker_tts_b200 import _lib, _runtime            # noqa: E402
from single_speaker_tts_b200.synthetic import make_clips      # noqa: E402

SR = 22050
lib = _lib.load()
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
clips = make_clips(256, seed=1, pool=16)
soff = np.concatenate([[0], np.cumsum([len(c) for c in clips])]).astype(np.int64)
wav = torch.from_numpy(np.concatenate(clips)).to(dev)
audio_s = float(soff[-1]) / SR
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
i64p = ctypes.POINTER(ctypes.c_int64)


def time_kernel(n_fft, win, hop, prec, fmax, outputs, reduction=1, normalize=False, force_generic=False):
    cfg = _runtime._make_config(n_fft, win, hop, prec, SR, 80, 0, fmax)
    plan = ctypes.c_void_p()
    _lib.check(lib.sstts_feat_plan_create(ctypes.byref(cfg), 256, soff.ctypes.data_as(i64p), reduction, ctypes.byref(plan)))
    rows = int(lib.sstts_feat_total_rows(plan))
    nb = n_fft // 2 + 1
    out = _lib.FeatOutputs()
    keep = []
    if 'lin' in outputs:
        t = torch.empty((rows, nb), dtype=torch.float32, device=dev); keep.append(t); out.lin_db_dev = t.data_ptr()
    if 'mel' in outputs:
        t = torch.empty((rows, 80), dtype=torch.float32, device=dev); keep.append(t); out.mel_db_dev = t.data_ptr()
    if 'spec' in outputs:
        t = torch.empty((rows, nb, 2), dtype=torch.float32, device=dev); keep.append(t); out.spec_dev = t.data_ptr()
    if 'raw' in outputs:
        t = torch.empty((rows, 80), dtype=torch.float64, device=dev); keep.append(t); out.mel_raw_dev = t.data_ptr()
    if 'minmax' in outputs:
        t = torch.empty((256, 4), dtype=torch.float64, device=dev); keep.append(t); out.minmax_dev = t.data_ptr()
    out.normalize = 1 if normalize else 0
    out.lin_ref_db, out.lin_max_db, out.mel_ref_db, out.mel_max_db = 35.66, 100.0, 6.02, 99.89
    out.mel_power = 1.0
    out.force_generic = 1 if force_generic else 0

    def step():
        _lib.check(lib.sstts_stft_features(plan, ctypes.c_void_p(wav.data_ptr()), ctypes.byref(out), stream))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    lib.sstts_feat_plan_destroy(plan)
    return e0.elapsed_time(e1) / 20


res = {}
for prec in ('f64', 'f32'):
    res['stats_minmax_' + prec] = time_kernel(1024, 1024, 256, prec, SR // 2, ('minmax',))
    res['stft1024_spec_' + prec] = time_kernel(1024, 1024, 256, prec, SR // 2, ('spec',))
    res['feat2048_fused_' + prec] = time_kernel(2048, 1102, 275, prec, 8000, ('lin', 'mel'), 5, True)
    res['feat2048_generic_' + prec] = time_kernel(2048, 1102, 275, prec, 8000, ('lin', 'mel'), 5, True, True)
    res['stft2048_spec_' + prec] = time_kernel(2048, 1102, 275, prec, 8000, ('spec',))
    res['mel2048_raw_' + prec] = time_kernel(2048, 1102, 275, prec, 8000, ('raw',))
print(json.dumps({k: round(v, 4) for k, v in res.items()}))
