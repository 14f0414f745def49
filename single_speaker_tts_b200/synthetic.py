"""Seeded synthetic, speech-shaped audio for benchmarks and parity tests.

There is no network and no corpus on the GPU box, so every workload of ``BASELINE.json`` runs
on synthetic clips shaped like LJSpeech (22.05 kHz mono float32): a harmonic stack on a
random-walk f0 (90-250 Hz) under a 3-6 Hz syllabic envelope, pink-ish noise 40 dB below the
voiced level, 50-300 ms of -70 dB "silence" on both ends, peak 0.5.
"""
import numpy as np

SAMPLING_RATE = 22050


def speech_like_clip(n_samples, rng, sr=SAMPLING_RATE):
    """One clip of ``n_samples`` float32 samples drawn from ``rng`` (np.random.Generator)."""
    n = int(n_samples)
    t = np.arange(n, dtype=np.float64) / sr
    # f0 random walk, smoothed to ~20 Hz control rate.
    n_ctl = max(4, n // (sr // 20) + 2)
    walk = np.cumsum(rng.normal(0.0, 6.0, n_ctl))
    f0_ctl = np.clip(160.0 + walk, 90.0, 250.0)
    f0 = np.interp(np.arange(n), np.linspace(0, n - 1, n_ctl), f0_ctl)
    phase = 2.0 * np.pi * np.cumsum(f0) / sr
    voiced = np.zeros(n, dtype=np.float64)
    n_harm = 24
    for h in range(1, n_harm + 1):
        amp = 1.0 / h ** 1.2
        voiced += amp * np.sin(h * phase + rng.uniform(0, 2 * np.pi))
    syll = rng.uniform(3.0, 6.0)
    env = 0.55 + 0.45 * np.sin(2.0 * np.pi * syll * t + rng.uniform(0, 2 * np.pi))
    voiced *= env
    voiced /= max(1e-9, np.max(np.abs(voiced)))
    # coloured noise 40 dB below the voiced level (one-pole low-pass of white noise).
    white = rng.normal(0.0, 1.0, n)
    spec = np.fft.rfft(white)
    f = np.arange(spec.shape[0], dtype=np.float64)
    spec /= np.sqrt(1.0 + f / max(1.0, spec.shape[0] / 64.0))
    noise = np.fft.irfft(spec, n)
    noise *= 0.01 / max(1e-9, np.sqrt(np.mean(noise ** 2)))
    x = voiced + noise
    # -70 dB leading / trailing silence.
    lead = min(n // 4, int(rng.uniform(0.05, 0.30) * sr))
    tail = min(n // 4, int(rng.uniform(0.05, 0.30) * sr))
    gate = np.ones(n, dtype=np.float64)
    floor = 10.0 ** (-70.0 / 20.0)
    gate[:lead] = floor
    if tail > 0:
        gate[n - tail:] = floor
    x *= gate
    x *= 0.5 / max(1e-9, np.max(np.abs(x)))
    return x.astype(np.float32)


def ragged_durations(n_clips, rng, kind='uniform', sr=SAMPLING_RATE):
    """Clip lengths in samples.  ``uniform``: U(1, 10) s (BASELINE configs 1, 2, 4);
    ``ljspeech``: clip(N(6.57, 2.19), 1.11, 10.10) s (config 3)."""
    if kind == 'uniform':
        dur = rng.uniform(1.0, 10.0, n_clips)
    elif kind == 'ljspeech':
        dur = np.clip(rng.normal(6.57, 2.19, n_clips), 1.11, 10.10)
    else:
        raise ValueError('unknown duration model {!r}'.format(kind))
    return (dur * sr).astype(np.int64)


def make_clips(n_clips, seed, kind='uniform', pool=None, sr=SAMPLING_RATE):
    """``n_clips`` ragged clips.  With ``pool=k`` only ``k`` distinct base clips of 10.1 s are
    synthesised and every clip is a seeded random crop of one of them (keeps host generation
    time bounded for the 13,100- and 4,096-clip workloads)."""
    rng = np.random.default_rng(seed)
    lengths = ragged_durations(n_clips, rng, kind=kind, sr=sr)
    if pool is None:
        return [speech_like_clip(n, rng, sr=sr) for n in lengths]
    base_len = int(10.2 * sr)
    bases = [speech_like_clip(base_len, rng, sr=sr) for _ in range(int(pool))]
    clips = []
    for n in lengths:
        b = bases[int(rng.integers(0, len(bases)))]
        start = int(rng.integers(0, base_len - n + 1))
        clips.append(np.ascontiguousarray(b[start:start + n]))
    return clips


class ClipPlan:
    """Deterministic description of a large synthetic corpus: ``n_clips`` ragged clips, each a
    seeded crop of one of ``pool`` base clips.  Every rank builds the same plan (cheap: no audio
    is synthesised per clip) and materialises only the clips of its own shard."""

    def __init__(self, n_clips, seed, kind='ljspeech', pool=32, sr=SAMPLING_RATE):
        rng = np.random.default_rng(seed)
        self.sr = sr
        self.lengths = ragged_durations(n_clips, rng, kind=kind, sr=sr)
        base_len = int(10.2 * sr)
        self.bases = [speech_like_clip(base_len, rng, sr=sr) for _ in range(int(pool))]
        self.base_id = rng.integers(0, len(self.bases), n_clips)
        self.start = (rng.random(n_clips) * (base_len - self.lengths + 1)).astype(np.int64)

    def frames(self, hop):
        return 1 + self.lengths // hop

    def clip(self, i):
        b = self.bases[int(self.base_id[i])]
        s = int(self.start[i])
        return b[s:s + int(self.lengths[i])]

    def clips(self, indices):
        return [self.clip(i) for i in indices]
