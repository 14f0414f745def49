"""sstts-b200: the audio hot path of yweweler/single-speaker-tts on B200."""


def pinned_empty(shape, dtype='float32'):
    """Page-locked, C-ordered numpy array: batches whose arrays live in such memory are uploaded
    without the staging copy (see ``_hostio.pinned_empty``)."""
    import numpy as np
    from . import _hostio
    return _hostio.pinned_empty(shape, np.dtype(dtype))


def set_io_threads(n):
    """Host threads used to pack ragged batches into pinned staging buffers (see ``_hostio.set_io_threads``)."""
    from . import _hostio
    _hostio.set_io_threads(n)
