"""Does host-side packing (pageable -> pinned memcpy on several threads) or a concurrent H2D stream slow the
device-resident Griffin-Lim kernels?   python tools/host_interference_probe.py"""
import ctypes
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from single_speaker_tts_b200 import _lib, _runtime            # noqa: E402
from single_speaker_tts_b200.synthetic import make_clips      # noqa: E402

WIN, HOP, NFFT = 1102, 275, 2048
lib = _lib.load()
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
clips = make_clips(256, seed=1, pool=16)
fb = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, want_spec=True, precision='f32', keep_on_device=True)
mag_dev = fb.spec.abs().contiguous()
frames = fb.frames
foff = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
cfg = _runtime._make_config(NFFT, WIN, HOP, 'f32')
plan = ctypes.c_void_p()
_lib.check(lib.sstts_gl_plan_create(ctypes.byref(cfg), 256, foff.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.byref(plan)))
ws = torch.empty(int(lib.sstts_gl_workspace_bytes(plan)), dtype=torch.uint8, device=dev)
wav = torch.empty(int(lib.sstts_gl_total_samples(plan)), dtype=torch.float32, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def step():
    _lib.check(lib.sstts_griffin_lim_seeded(plan, ctypes.c_void_p(mag_dev.data_ptr()), ctypes.c_uint64(1), 0, 50,
                                            ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(wav.data_ptr()), None, st))


def timed(n=5):
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print('idle host: %.2f ms per 50-iteration call' % timed())
src = np.random.rand(64 << 20).astype(np.float32)            # 256 MB pageable
dst = torch.empty(64 << 20, dtype=torch.float32, pin_memory=True).numpy()
stop = False


def churn(i):
    n = len(src) // 8
    while not stop:
        np.copyto(dst[i * n:(i + 1) * n], src[i * n:(i + 1) * n])


ths = [threading.Thread(target=churn, args=(i,)) for i in range(8)]
for t in ths:
    t.start()
print('8 host threads copying pageable -> pinned: %.2f ms' % timed())
stop = True
for t in ths:
    t.join()
side = torch.cuda.Stream()
big = torch.empty(128 << 20, dtype=torch.float32, pin_memory=True)
dbig = torch.empty(128 << 20, dtype=torch.float32, device=dev)
stop = False


def h2d():
    while not stop:
        with torch.cuda.stream(side):
            dbig.copy_(big, non_blocking=True)
        side.synchronize()


t = threading.Thread(target=h2d); t.start()
print('continuous H2D on a side stream: %.2f ms' % timed())
stop = True; t.join()
stop = False


def d2h():
    while not stop:
        with torch.cuda.stream(side):
            big.copy_(dbig, non_blocking=True)
        side.synchronize()


t = threading.Thread(target=d2h); t.start()
print('continuous D2H on a side stream: %.2f ms' % timed())
stop = True; t.join()
print('idle host again: %.2f ms' % timed())
