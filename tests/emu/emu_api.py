"""ctypes wrapper of tests/emu/libsstts_emu.so -- the real kernel sources compiled for the CPU
SIMT emulator (TEST INFRASTRUCTURE; see cpu_simt.h).  Mirrors the packing the product's
``_runtime`` does, so emulator tests read like the GPU parity tests."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_fp = ctypes.POINTER(ctypes.c_float)
_dp = ctypes.POINTER(ctypes.c_double)
_lp = ctypes.POINTER(ctypes.c_longlong)
_lib = None


def load():
    global _lib
    if _lib is None:
        sys.path.insert(0, ROOT)
        import __graft_entry__ as ge
        path = ge.build_emulator()
        lib = ctypes.CDLL(path)
        lib.emu_fft1024.argtypes = [_dp, _dp, ctypes.c_int, ctypes.c_int]
        lib.emu_griffin_lim.argtypes = [ctypes.c_int] * 4 + [_lp, _fp, _fp, ctypes.c_int, _fp, _dp,
                                                           ctypes.c_int, ctypes.c_int]
        lib.emu_stft_features.argtypes = ([ctypes.c_int] * 6 + [ctypes.c_double] * 2 +
                                          [ctypes.c_int, _lp, ctypes.c_int, _fp, _fp, _fp, _fp, _dp,
                                           _dp, ctypes.c_int] + [ctypes.c_double] * 5 + [ctypes.c_int] * 2)
        lib.emu_mel_basis.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double] * 2 + [_dp]
        lib.emu_trim_bounds.argtypes = [_fp, ctypes.c_int, _lp, _lp, ctypes.c_double, ctypes.c_int, ctypes.c_int, _lp]
        _lib = lib
    return sys.modules[__name__]


def fft1024(z, inverse=False, prec=1):
    inp = np.empty(2048)
    inp[0::2], inp[1::2] = z.real, z.imag
    out = np.empty(2048)
    assert _lib.emu_fft1024(inp.ctypes.data_as(_dp), out.ctypes.data_as(_dp), int(inverse), prec) == 0
    return out[0::2] + 1j * out[1::2]


def griffin_lim(mags, angles, n_iter, prec=0, win=1102, hop=275, want_mse=False, grid_cap=3, n_fft=2048):
    Ts = [m.shape[1] for m in mags]
    fo = np.concatenate([[0], np.cumsum(Ts)]).astype(np.int64)
    mag = np.ascontiguousarray(np.concatenate([np.asarray(m).T for m in mags], 0), dtype=np.float32)
    ph = None if angles is None else np.ascontiguousarray(np.concatenate([a.T for a in angles], 0).astype(np.complex64))
    so = np.concatenate([[0], np.cumsum([hop * (t - 1) for t in Ts])]).astype(np.int64)
    out = np.full(max(1, so[-1]), np.nan, dtype=np.float32)
    mse = np.zeros(fo[-1]) if want_mse else None
    rc = _lib.emu_griffin_lim(win, hop, prec, len(mags), fo.ctypes.data_as(_lp), mag.ctypes.data_as(_fp),
                              None if ph is None else ph.view(np.float32).ctypes.data_as(_fp), n_iter, out.ctypes.data_as(_fp),
                              mse.ctypes.data_as(_dp) if want_mse else None, grid_cap, n_fft)
    assert rc == 0
    wavs = [out[so[i]:so[i + 1]] for i in range(len(mags))]
    if want_mse:
        return wavs, [mse[fo[i]:fo[i + 1]].sum() / ((1 + n_fft // 2) * Ts[i]) for i in range(len(mags))]
    return wavs


def stft_features(wavs, prec=1, r=1, n_fft=2048, win=1102, hop=275, sr=22050, n_mels=80, fmin=0.,
                  fmax=8000., normalize=None, power=1.0, grid_cap=3, fast=False):
    """fast=True requests only lin + mel dB (n_fft 1024: + the per-clip extrema), which selects the kernel's
    fused float32-epilogue mode."""
    nb = n_fft // 2 + 1
    so = np.concatenate([[0], np.cumsum([len(w) for w in wavs])]).astype(np.int64)
    wav = np.concatenate(wavs).astype(np.float32)
    Ts = [1 + len(w) // hop for w in wavs]
    ro = np.concatenate([[0], np.cumsum([-(-t // r) * r for t in Ts])])
    R = int(ro[-1])
    spec = np.full((R, nb), np.nan, np.complex64)
    lin = np.full((R, nb), np.nan, np.float32)
    mel = np.full((R, n_mels), np.nan, np.float32)
    melraw = np.full((R, n_mels), np.nan)
    mm = np.zeros((len(wavs), 4))
    consts = normalize if normalize is not None else (0., 0., 0., 0.)
    rc = _lib.emu_stft_features(n_fft, win, hop, prec, sr, n_mels, fmin, fmax, len(wavs),
                                so.ctypes.data_as(_lp), r, wav.ctypes.data_as(_fp),
                                None if fast else spec.view(np.float32).ctypes.data_as(_fp), lin.ctypes.data_as(_fp),
                                mel.ctypes.data_as(_fp), None if fast else melraw.ctypes.data_as(_dp),
                                None if (fast and n_fft != 1024) else mm.ctypes.data_as(_dp),
                                int(normalize is not None), *consts, power, grid_cap, int(fast))
    assert rc == 0
    return [dict(spec=spec[ro[i]:ro[i + 1]], lin=lin[ro[i]:ro[i + 1]], mel=mel[ro[i]:ro[i + 1]],
                 melraw=melraw[ro[i]:ro[i + 1]], minmax=mm[i], T=Ts[i]) for i in range(len(wavs))]


def mel_basis(sr, n_fft, n_mels, fmin, fmax):
    mb = np.zeros((n_mels, n_fft // 2 + 1))
    _lib.emu_mel_basis(sr, n_fft, n_mels, float(fmin), float(fmax), mb.ctypes.data_as(_dp))
    return mb


def trim_bounds(wavs, top_db=60.0, frame_length=2048, hop_length=512):
    lens = np.asarray([len(w) for w in wavs], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    wav = np.concatenate(wavs).astype(np.float32)
    bounds = np.zeros((len(wavs), 2), dtype=np.int64)
    _lib.emu_trim_bounds(wav.ctypes.data_as(_fp), len(wavs), starts.ctypes.data_as(_lp), lens.ctypes.data_as(_lp),
                         float(top_db), frame_length, hop_length, bounds.ctypes.data_as(_lp))
    return bounds
