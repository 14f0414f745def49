"""Device-resident timing of Griffin-Lim at n_fft 1024 / hop 256 / win 1024 (the geometry of
audio/effects.py:71-86 and datasets/statistics.py:31-34) on the 256-clip benchmark set, 25 iterations,
through the C ABI: ms per call and per iteration launch, audio-seconds per second.

    python tools/gl1024_probe.py            # native 512-point complex transform (two frames per warp)
    SSTTS_GL_NATIVE1024=0 python tools/gl1024_probe.py   # embedded in the 2048-point transform (the older path)
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

SR, NFFT, WIN, HOP, ITERS = 22050, 1024, 1024, 256, 25
lib = _lib.load()
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
clips = make_clips(256, seed=1, pool=16)
fb = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, want_spec=True, precision='f64', keep_on_device=True)
mag = fb.spec.abs().contiguous()
frames = [1 + len(c) // HOP for c in clips]
foff = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
audio_s = sum(HOP * (t - 1) for t in frames) / SR
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
res = {'native': os.environ.get('SSTTS_GL_NATIVE1024', '1') != '0', 'frames': int(foff[-1])}
for prec in ('f32', 'f64'):
    cfg = _runtime._make_config(NFFT, WIN, HOP, prec)
    plan = ctypes.c_void_p()
    _lib.check(lib.sstts_gl_plan_create(ctypes.byref(cfg), 256, foff.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                        ctypes.byref(plan)))
    ws = torch.empty(int(lib.sstts_gl_workspace_bytes(plan)), dtype=torch.uint8, device=dev)
    wav = torch.empty(int(lib.sstts_gl_total_samples(plan)), dtype=torch.float32, device=dev)

    def step(n_iter):
        _lib.check(lib.sstts_griffin_lim_seeded(plan, ctypes.c_void_p(mag.data_ptr()), ctypes.c_uint64(7), 0, n_iter,
                                                ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(wav.data_ptr()), None, stream))

    def timed(n_iter, reps):
        for _ in range(3):
            step(n_iter)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            step(n_iter)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    reps = 10 if prec == 'f32' else 4
    full, base = timed(ITERS, reps), timed(0, reps)
    res[prec] = {'ms_per_call': round(full, 3), 'ms_per_iteration': round((full - base) / ITERS, 4),
                 'audio_s_per_s': round(audio_s / (full / 1000.0), 1), 'checksum': float(wav.double().abs().mean().item())}
    lib.sstts_gl_plan_destroy(plan)
print(json.dumps(res))
