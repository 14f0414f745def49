"""Single-utterance latency breakdown (tacotron/serve.py:39-86 shape: one 1000-frame utterance, 50 it).

    python tools/latency_probe.py [T] [n_iter]

Prints one JSON object: the device-resident C-ABI call (CUDA events: kernels + launch gaps), the per-item
drop-in call (numpy global RNG phase), the seeded batched call with one item, and the host-side pieces."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from single_speaker_tts_b200 import _lib, _runtime  # noqa: E402
from single_speaker_tts_b200.audio import synthesis  # noqa: E402

WIN, HOP, NFFT = 1102, 275, 2048
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
N_ITER = int(sys.argv[2]) if len(sys.argv) > 2 else 50


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(round(q * (len(v) - 1))))]


def main():
    lib = _lib.load()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    rng = np.random.default_rng(7)
    mag = (rng.random((T, 1025), dtype=np.float32) ** 4 * 20.0).T
    out = {'T': T, 'n_iter': N_ITER, 'audio_s': HOP * (T - 1) / 22050.0}

    cfg = _runtime._make_config(NFFT, WIN, HOP, 'f32')
    plan = ctypes.c_void_p()
    fo = np.array([0, T], dtype=np.int64)
    _lib.check(lib.sstts_gl_plan_create(ctypes.byref(cfg), 1, fo.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                        ctypes.byref(plan)))
    mag_dev = torch.from_numpy(np.ascontiguousarray(mag.T)).to(dev)
    ws = torch.empty(int(lib.sstts_gl_workspace_bytes(plan)), dtype=torch.uint8, device=dev)
    wav = torch.empty(HOP * (T - 1), dtype=torch.float32, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def dev_call(n_iter=N_ITER):
        _lib.check(lib.sstts_griffin_lim_seeded(plan, ctypes.c_void_p(mag_dev.data_ptr()), ctypes.c_uint64(1), 0, n_iter,
                                                ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(wav.data_ptr()), None, st))

    for _ in range(5):
        dev_call()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
    ev[0].record()
    for i in range(40):
        dev_call()
        ev[i + 1].record()
    torch.cuda.synchronize()
    d = [ev[i].elapsed_time(ev[i + 1]) for i in range(40)]
    out['device_call_ms'] = {'p50': pct(d, 0.5), 'min': min(d), 'p99': pct(d, 0.99)}
    t0 = time.perf_counter()
    for _ in range(40):
        dev_call()
    out['host_enqueue_ms_per_call'] = (time.perf_counter() - t0) / 40 * 1000
    torch.cuda.synchronize()

    def timeit(fn, n=30, warm=5):
        for _ in range(warm):
            fn()
        v = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            v.append((time.perf_counter() - t0) * 1000)
        return {'p50': pct(v, 0.5), 'min': min(v), 'p99': pct(v, 0.99)}

    out['dropin_spectrogram_to_wav_ms'] = timeit(lambda: synthesis.spectrogram_to_wav(mag, WIN, HOP, NFFT, N_ITER))
    out['seeded_single_item_ms'] = timeit(lambda: synthesis.spectrograms_to_wavs([mag], WIN, HOP, NFFT, N_ITER, seed=3))
    out['zero_iter_single_item_ms'] = timeit(lambda: synthesis.spectrograms_to_wavs([mag], WIN, HOP, NFFT, 0, seed=3))
    out['host_np_random_rand_ms'] = timeit(lambda: np.random.rand(1025, T), n=10, warm=2)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
