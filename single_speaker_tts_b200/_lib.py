"""ctypes binding of ``csrc/libsstts.so`` (C ABI declared in ``include/sstts.h``).

There is NO CPU fallback: if the shared library has not been built, or no CUDA device is
visible, every compute entry point of this package raises.  Build the library with
``python __graft_entry__.py build`` (one ``nvcc -shared`` command for sm_100a, see ``build()`` there).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# SSTTS_LIB selects another build of the same library (kernel A/B experiments under tools/).
LIB_PATH = os.environ.get('SSTTS_LIB') or os.path.join(_HERE, 'csrc', 'libsstts.so')

SSTTS_F32 = 0
SSTTS_F64 = 1

_c_i64_p = ctypes.POINTER(ctypes.c_int64)
_c_f32_p = ctypes.POINTER(ctypes.c_float)
_c_f64_p = ctypes.POINTER(ctypes.c_double)


class StftConfig(ctypes.Structure):
    """``sstts_stft_config`` (include/sstts.h)."""
    _fields_ = [('n_fft', ctypes.c_int), ('win_length', ctypes.c_int), ('hop_length', ctypes.c_int),
                ('sampling_rate', ctypes.c_int), ('n_mels', ctypes.c_int),
                ('mel_fmin', ctypes.c_double), ('mel_fmax', ctypes.c_double),
                ('precision', ctypes.c_int)]


class FeatOutputs(ctypes.Structure):
    """``sstts_feat_outputs`` (include/sstts.h)."""
    _fields_ = [('spec_dev', ctypes.c_void_p), ('lin_db_dev', ctypes.c_void_p),
                ('mel_db_dev', ctypes.c_void_p), ('mel_raw_dev', ctypes.c_void_p),
                ('minmax_dev', ctypes.c_void_p), ('normalize', ctypes.c_int),
                ('lin_ref_db', ctypes.c_double), ('lin_max_db', ctypes.c_double),
                ('mel_ref_db', ctypes.c_double), ('mel_max_db', ctypes.c_double),
                ('mel_power', ctypes.c_double), ('force_generic', ctypes.c_int)]


# name -> (restype, argtypes); also the list the symbol-export test checks against the header.
SIGNATURES = {
    'sstts_version': (ctypes.c_int, []),
    'sstts_last_error': (ctypes.c_char_p, []),
    'sstts_device_count': (ctypes.c_int, []),
    'sstts_gl_plan_create': (ctypes.c_int, [ctypes.POINTER(StftConfig), ctypes.c_int, _c_i64_p,
                                            ctypes.POINTER(ctypes.c_void_p)]),
    'sstts_gl_plan_destroy': (None, [ctypes.c_void_p]),
    'sstts_gl_workspace_bytes': (ctypes.c_size_t, [ctypes.c_void_p]),
    'sstts_gl_total_frames': (ctypes.c_int64, [ctypes.c_void_p]),
    'sstts_gl_total_samples': (ctypes.c_int64, [ctypes.c_void_p]),
    'sstts_gl_sample_offsets': (_c_i64_p, [ctypes.c_void_p]),
    'sstts_griffin_lim': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    'sstts_griffin_lim_seeded': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64,
                                                ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p]),
    'sstts_peak_normalize': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'sstts_phase_from_uniform': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                                ctypes.c_void_p]),
    'sstts_random_phase': (ctypes.c_int, [ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p,
                                          ctypes.c_void_p]),
    'sstts_random_phase_at': (ctypes.c_int, [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    'sstts_pcm16_to_float': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    'sstts_dct_project': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    'sstts_stretch_frames': (ctypes.c_int64, [ctypes.c_int64, ctypes.c_double]),
    'sstts_stretch_magnitude': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
                                               ctypes.c_void_p, ctypes.c_void_p]),
    'sstts_denormalize_magnitude': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                                   ctypes.c_double, ctypes.c_double, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_void_p]),
    'sstts_feat_plan_create': (ctypes.c_int, [ctypes.POINTER(StftConfig), ctypes.c_int, _c_i64_p,
                                              ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    'sstts_feat_plan_create_ranges': (ctypes.c_int, [ctypes.POINTER(StftConfig), ctypes.c_int, _c_i64_p, _c_i64_p,
                                                     ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    'sstts_feat_plan_destroy': (None, [ctypes.c_void_p]),
    'sstts_feat_total_frames': (ctypes.c_int64, [ctypes.c_void_p]),
    'sstts_feat_total_rows': (ctypes.c_int64, [ctypes.c_void_p]),
    'sstts_feat_frame_offsets': (_c_i64_p, [ctypes.c_void_p]),
    'sstts_feat_row_offsets': (_c_i64_p, [ctypes.c_void_p]),
    'sstts_feat_mel_basis': (_c_f64_p, [ctypes.c_void_p]),
    'sstts_stft_features': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.POINTER(FeatOutputs), ctypes.c_void_p]),
    'sstts_trim_bounds': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p]),
}

_lib = None
_lock = threading.Lock()


class SsttsError(RuntimeError):
    """A libsstts call failed (message from ``sstts_last_error``)."""


def load():
    """Load libsstts.so once per process; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise SsttsError(
                'libsstts.so is missing at {} -- build it with `python __graft_entry__.py build`. '
                'single_speaker_tts_b200 has no CPU fallback.'.format(LIB_PATH))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc):
    """Raise on a negative sstts status; AssertionError for SSTTS_ERR_ASSERT."""
    if rc is not None and rc < 0:
        msg = load().sstts_last_error().decode('utf-8', 'replace')
        if rc == -4:
            raise AssertionError(msg)
        raise SsttsError('libsstts error {}: {}'.format(rc, msg))
    return rc
