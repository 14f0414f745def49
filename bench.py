#!/usr/bin/env python
"""Benchmark of the audio hot path (BASELINE.json): batched 50-iteration Griffin-Lim and the
STFT feature pipeline, in audio-seconds per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" is one pass of the hot path over one batch.  The primary metric is BASELINE configs[2]
-- Griffin-Lim, 50 iterations, 256 ragged synthetic utterances (1-10 s, 22.05 kHz) per GPU;
`value` is device-resident throughput through the C ABI, `e2e` goes through the public Python API
with ordinary (pageable) host numpy buffers, H2D and D2H inside the timed region.  Weak scaling:
every rank owns its own 256 utterances, no data-path collective (utterances are independent).

The same JSON line carries, under their own keys,
  * "features":   BASELINE configs[1] (the same 256 clips) -- device-resident float64 (default) and
                  float32 transforms with their rooflines, and end to end through features_batch;
  * "statistics": the n_fft 1024 / hop 256 dB-statistics kernel of datasets/statistics.py:31-34;
  * "griffin_lim_n_fft_1024": 25 iterations at n_fft 1024 / hop 256 (audio/effects.py:71-86), device-resident;
  * "latency":    one 1000-frame utterance through spectrogram_to_wav (tacotron/serve.py:39-86),
                  p50 / p99, alone and with 6 concurrent callers (:69-72);
  * "corpus":     BASELINE configs[3] -- 13,100 LJSpeech-length clips sharded by clip over the ranks,
                  statistics + NCCL all-reduce + feature pre-calculation, STRONG scaling;
  * "gl4096":     BASELINE configs[4] -- 4,096 utterances, 100 iterations, sharded, STRONG scaling;
  * "cpu_baseline": the reference's CPU path (numpy restatement) timed on this box's cores, N = 1 only.

--impl reference times the reference's CPU implementation of the same path (the numpy oracle
restating librosa 0.6 + audio/synthesis.py; the reference itself cannot be installed: librosa and
TensorFlow 1.8 are unavailable offline) on all host cores on a bounded sample drawn from the same
256-utterance set.
"""
import argparse
import json
import os
import statistics as pystats
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIN, HOP, NFFT, SR = 1102, 275, 2048, 22050
N_BINS = 1025
N_UTTS = 256
GL_ITERS = 50
# algorithmic bytes per frame and Griffin-Lim iteration
# (phase read + |S| read + phase write + waveform write + waveform read).
GL_BYTES_PER_FRAME_ITER = 22700
GL_BYTES_FINAL_PER_FRAME = 13400
FEAT_BYTES_PER_FRAME = 4420   # (1025 + 80) float32 written per frame; + 4 bytes per input sample
CONSTS = (35.66, 100.0, 6.02, 99.89)   # datasets/lj_speech.py:20-29 (lin ref, lin max, mel ref, mel max)

GL_KERNEL = 'gl_step_kernel<float, StaticGeom<1102, 275, 2048>, 8, 0, 0, 0>'
FEAT_KERNELS = {'f64': 'stft_feature_kernel<double, StaticGeom<1102, 275, 2048>, 4, 1>',
                'f32': 'stft_feature_kernel<float, StaticGeom<1102, 275, 2048>, 8, 1>'}


def workload_config():
    """`config` of the JSON line -- identical for both arms (the reference arm runs a bounded sample
    of the very same utterance set, described in its `cpu_baseline.sample`)."""
    return {'workload': 'BASELINE configs[2]: Griffin-Lim 50 it over 256 ragged synthetic utterances '
                        '(1-10 s, 22.05 kHz) per GPU', 'n_fft': NFFT, 'win': WIN, 'hop': HOP,
            'n_iter': GL_ITERS, 'n_utterances_per_gpu': N_UTTS, 'clip_seed': 1,
            'l2': 'inputs_exceed_l2 (|S| 4.1 KB/frame + waveform state >> 126 MB)',
            'sharding': 'by utterance, no collective; every rank runs the same 256-clip set'}


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


_NCU_FILES = ('r2_traffic.json', 'r1_traffic.json')


def measured_ncu(kernel, key):
    """One metric of `kernel` from the committed ncu --set full capture (profiles/r2_traffic.json, else the
    round-1 file), or None."""
    for name in _NCU_FILES:
        try:
            with open(os.path.join(ROOT, 'profiles', name)) as f:
                return json.load(f)[kernel][key]
        except (OSError, KeyError, ValueError):
            continue
    return None


def ncu_block(kernel):
    """The per-kernel pipe utilisation north_star asks for, from the committed ncu capture."""
    keys = ('fp32_pipe_active_pct', 'fp64_pipe_active_pct', 'issue_slots_active_pct', 'lsu_pipe_active_pct',
            'dram_throughput_pct')
    out = {k + '_ncu': measured_ncu(kernel, k) for k in keys}
    return {k: v for k, v in out.items() if v is not None}


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': pystats.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on host cores
# ------------------------------------------------------------------------------------------------
def _single_thread_env():
    """One thread per worker process: the BLAS behind the reference's dense mel np.dot would otherwise
    start one thread per core in every forked worker (environment variables come too late after fork)."""
    sys.path.insert(0, ROOT)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        os.environ['OMP_NUM_THREADS'] = os.environ['MKL_NUM_THREADS'] = os.environ['OPENBLAS_NUM_THREADS'] = '1'


def _oracle_gl_worker(args):
    """One utterance of the reference's per-item path: |STFT| -> 50 Griffin-Lim iterations
    (audio/synthesis.py:5-125 on librosa 0.6 numerics)."""
    _single_thread_env()
    from oracle import reference_audio as ra
    mag, seed, n_iter = args
    angles = np.exp(2j * np.pi * np.random.RandomState(seed).rand(*mag.shape))
    wav = ra.spectrogram_to_wav(mag, WIN, HOP, NFFT, n_iter, angles=angles)
    return len(wav)


def _oracle_feature_worker(clip):
    """The reference's load_audio recipe after the file decode (datasets/lj_speech.py:119-156):
    trim, TWO STFTs, dense 80 x 1025 mel np.dot, dB, normalise, reduction padding."""
    _single_thread_env()
    from oracle import reference_audio as ra
    mel, lin = ra.load_audio_from_wav(clip, SR, trim=True)
    return len(clip)


def workload_sample(n_items, order='first'):
    """`n_items` utterances of the benchmark's own 256-clip set (same seed, same ragged 1-10 s length
    distribution): the first ones ('first'), or the ones at evenly spaced length ranks ('spread')."""
    from single_speaker_tts_b200.synthetic import make_clips
    clips = make_clips(N_UTTS, seed=1, pool=16)
    if order == 'spread':
        rank = np.argsort([len(c) for c in clips], kind='stable')
        idx = [int(rank[int(round(q))]) for q in np.linspace(0, N_UTTS - 1, n_items)]
    else:
        idx = list(range(n_items))
    return [clips[i] for i in idx], idx


def oracle_mags(clips):
    from oracle import librosa_compat as lc
    return [np.abs(lc.stft(c, NFFT, HOP, WIN)) for c in clips]


def time_oracle_gl(pool, mags, seed0, n_iter=GL_ITERS):
    """Pool.map over utterances, longest first, one task at a time per worker (dynamic balance)."""
    order = np.argsort([-m.shape[1] for m in mags], kind='stable')
    t0 = time.perf_counter()
    lens = pool.map(_oracle_gl_worker, [(mags[i], seed0 + int(i), n_iter) for i in order], chunksize=1)
    dt = time.perf_counter() - t0
    return sum(lens) / SR, dt


def run_reference(args):
    """--impl reference: the oracle (kind "port": the reference's numpy/librosa code restated; the
    reference itself is not installable offline) on all host cores.  Each step is a bounded sample of
    the workload: 2 x cores utterances at evenly spaced length ranks of the same 256-utterance set."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    n_items = min(N_UTTS, args.ref_items or 2 * cores)
    clips, idx = workload_sample(n_items, order='spread')
    mags = oracle_mags(clips)
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        for _ in range(min(args.warmup, 1)):       # one warm-up pass is enough for a CPU pool (forked, numpy warm)
            time_oracle_gl(pool, mags, 1000)
        audio, total = 0.0, 0.0
        for s in range(args.steps):
            a, dt = time_oracle_gl(pool, mags, 1000 + 7 * s)
            audio += a
            total += dt
    value = audio / total
    sample = ('{} of the 256 utterances per step (evenly spaced length ranks, {:.1f} s of audio), one process per '
              'core, {} iterations; same utterance set, lengths and iteration count as the GPU arm').format(
                  n_items, audio / max(1, args.steps), GL_ITERS)
    line = {
        'impl': 'reference', 'metric': 'griffin_lim_50it_audio_sec_per_sec', 'value': value,
        'unit': 'audio-s/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1000.0 * total / max(1, args.steps), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(),
        'cpu_baseline': {'value': value, 'unit': 'audio-s/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'audio-s/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'arithmetic': 'float64 FFTs, float32 overlap-add, complex64 spectra (librosa 0.6 / audio/synthesis.py)',
    }
    print(json.dumps(line), flush=True)


def cpu_baselines():
    """The reference CPU path on this box's cores (reported baseline, SURVEY.md section 8d i-iii); about
    20-30 s of CPU work in total."""
    import multiprocessing as mp
    from oracle import librosa_compat as lc
    from oracle import reference_audio as ra
    from single_speaker_tts_b200.synthetic import speech_like_clip
    cores = len(os.sched_getaffinity(0))
    ctx = mp.get_context('fork')
    out = {}
    # (ii) the inference pool: 6 workers (tacotron/params/inference.py:34) over 16 utterances of the set
    use = min(6, cores)
    clips, _ = workload_sample(16, order='spread')
    mags = oracle_mags(clips)
    with ctx.Pool(use) as pool:
        a, dt = time_oracle_gl(pool, mags, 1000)
    out.update({'value': a / dt, 'unit': 'audio-s/s', 'cores': use, 'kind': 'port',
                'sample': '16 of the 256 utterances (evenly spaced length ranks, {:.1f} s of audio), {} iterations, '
                          'pool of {} processes as in tacotron/inference.py:185 (host has {} cores)'.format(
                              a, GL_ITERS, use, cores)})
    # (i) BASELINE configs[0]: one 5 s clip, single process -- literally the reference's per-utterance path
    clip = speech_like_clip(5 * SR, np.random.default_rng(0))
    mag = np.abs(lc.stft(clip, NFFT, HOP, WIN))
    ang = np.exp(2j * np.pi * np.random.RandomState(0).rand(*mag.shape))
    t0 = time.perf_counter()
    wav = ra.spectrogram_to_wav(mag, WIN, HOP, NFFT, GL_ITERS, angles=ang)
    dt = time.perf_counter() - t0
    out['config0_single_process'] = {'value': len(wav) / SR / dt, 'unit': 'audio-s/s', 'cores': 1, 'seconds': dt,
                                     'sample': 'BASELINE configs[0]: one 5 s clip (T = 401), 50 iterations'}
    # (iii) feature recipe with the reference's two STFTs + dense mel dot, 64 clips, 1 process and all cores
    fclips, _ = workload_sample(64, order='spread')
    audio = sum(len(c) for c in fclips) / SR
    t0 = time.perf_counter()
    for c in fclips[::4]:
        _oracle_feature_worker(c)
    dt1 = time.perf_counter() - t0
    audio1 = sum(len(c) for c in fclips[::4]) / SR
    with ctx.Pool(cores) as pool:
        pool.map(_oracle_feature_worker, fclips[:cores], chunksize=1)       # warm the workers
        t0 = time.perf_counter()
        pool.map(_oracle_feature_worker, fclips, chunksize=1)
        dtn = time.perf_counter() - t0
    out['features'] = {'value': audio / dtn, 'unit': 'audio-s/s', 'cores': cores, 'kind': 'port',
                       'single_process_single_thread_value': audio1 / dt1,
                       'sample': '64 of the 256 clips ({:.0f} s of audio) through the load_audio recipe after decode '
                                 '(datasets/lj_speech.py:119-156: trim, two STFTs, dense mel np.dot); single process '
                                 'on every 4th of them'.format(audio)}
    return out


def cpu_baselines_in_subprocess(timeout_s=420):
    """The CPU baselines fork worker pools; forking THIS process -- CUDA initialised, helper threads of the
    staging pipeline alive -- can leave a child waiting for a lock that a thread which does not exist in the
    child held at fork time.  They therefore run in a fresh interpreter (bench.py --cpu-baseline-only), with
    a time limit so that the benchmark line is printed in any case."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), '--cpu-baseline-only'],
                             capture_output=True, text=True, timeout=timeout_s)
        lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
        if out.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {'error': 'cpu baseline subprocess failed (rc %d): %s' % (out.returncode, out.stderr[-300:])}
    except subprocess.TimeoutExpired:
        return {'error': 'cpu baseline subprocess exceeded %d s' % timeout_s}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Harness:
    """Process-group plumbing and CUDA-event timing shared by all sub-benchmarks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device('cuda', self.local_rank)
        if self.world > 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=self.dev)
        # one process per GPU on a shared host: the ranks split the cores for the (memory-bound) host-side packing
        import single_speaker_tts_b200 as pkg
        cores = len(os.sched_getaffinity(0))
        self.io_threads = max(2, min(8, cores // self.world))
        pkg.set_io_threads(self.io_threads)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def timed(self, fn, steps, warmup, per_call=False):
        """`warmup` untimed calls, then `steps` calls between barriers; CUDA-event time on the current
        stream, max over ranks.  per_call: also this rank's per-call times (an event after every call;
        the end-to-end calls synchronise internally, so the list shows host-side hiccups)."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            fn()
            if per_call or i == steps - 1:
                ev[i + 1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[steps])
        calls = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)] if per_call else None
        self.barrier()
        ms = self.max_over_ranks(ms)
        return (ms, calls) if per_call else ms


def host_pack_bandwidth(H):
    """Aggregate rate at which all ranks together can pack pageable arrays into pinned staging buffers (the
    host-side step of every end-to-end call with ordinary numpy inputs), measured with the ranks running
    simultaneously: a 256 MB copy per rank on its io threads.  End-to-end numbers at N GPUs cannot exceed
    this rate divided by the staged bytes per audio-second."""
    from concurrent.futures import ThreadPoolExecutor
    torch = H.torch
    n = 64 << 20
    src = np.ones(n, dtype=np.float32)
    dst = torch.empty(n, dtype=torch.float32, pin_memory=True).numpy()
    parts = H.io_threads
    step = n // parts

    def run(ex):
        list(ex.map(lambda i: np.copyto(dst[i * step:(i + 1) * step], src[i * step:(i + 1) * step]), range(parts)))

    with ThreadPoolExecutor(max_workers=parts) as ex:
        run(ex)
        H.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            run(ex)
        dt = (time.perf_counter() - t0) / 3
    dt = H.max_over_ranks(dt)
    return H.world * n * 4 / dt / 1e9


def host_link_bandwidth(H):
    """Aggregate device->host and host->device rates with every rank copying at once (256 MB per rank and
    direction between a device buffer and page-locked memory).  The feature / corpus paths return 354 KB of
    float32 features per audio-second, so their end-to-end rate at N GPUs cannot exceed d2h / 354 KB whatever
    N is -- on the boxes of this pool the aggregate D2H rate stops growing after two GPUs."""
    torch = H.torch
    n = 64 << 20
    host = torch.empty(n, dtype=torch.float32, pin_memory=True)
    dev = torch.empty(n, dtype=torch.float32, device=H.dev)
    out = {}
    for name, (dst, src) in (('d2h_gbs_all_ranks', (host, dev)), ('h2d_gbs_all_ranks', (dev, host))):
        dst.copy_(src, non_blocking=True)
        ms = H.timed(lambda: dst.copy_(src, non_blocking=True), 3, 1)
        out[name] = H.world * 3 * n * 4 / (ms / 1000.0) / 1e9
    return out


def bench_primary(H, args):
    """configs[2] (Griffin-Lim, the JSON line's metric) and configs[1] (features) on 256 clips per rank."""
    import ctypes
    torch = H.torch
    import single_speaker_tts_b200 as pkg
    from single_speaker_tts_b200 import _lib, _runtime
    from single_speaker_tts_b200.audio import features as feat_api
    from single_speaker_tts_b200.audio import synthesis
    from single_speaker_tts_b200.synthetic import make_clips
    lib = _lib.load()
    dev = H.dev
    peak_gbs, peak_src = load_peaks()

    # weak scaling: every rank owns the same 256-clip workload (same seed), so per-GPU work is fixed exactly
    clips = make_clips(N_UTTS, seed=1, pool=16)
    audio_in_s = sum(len(c) for c in clips) / SR
    frames = [1 + len(c) // HOP for c in clips]
    total_frames = sum(frames)
    audio_out_s = sum(HOP * (t - 1) for t in frames) / SR

    # magnitudes |STFT(clip)| produced on the device by our own feature kernel (float64 transform)
    fb = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, want_spec=True, precision='f64', keep_on_device=True)
    mag_dev = fb.spec.abs().contiguous()                      # (sum T, 1025) float32, frame-major
    del fb
    # host copies of the inputs: ordinary pageable numpy memory (what a TF session.run hands over; staged
    # through pooled pinned buffers by worker threads) -- the stated e2e -- and page-locked arrays
    # (pkg.pinned_empty: uploaded with no staging copy), reported beside it
    mag_pageable = mag_dev.cpu().numpy()
    mag_host = pkg.pinned_empty(tuple(mag_dev.shape))
    mag_host[:] = mag_pageable
    foff = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
    mags_pinned = [mag_host[foff[i]:foff[i + 1]].T for i in range(N_UTTS)]   # (1025, T) views, like spec.T
    mags_pageable = [mag_pageable[foff[i]:foff[i + 1]].T for i in range(N_UTTS)]
    clip_pin = pkg.pinned_empty((sum(len(c) for c in clips),))
    clip_pin[:] = np.concatenate(clips)
    soff = np.concatenate([[0], np.cumsum([len(c) for c in clips])]).astype(np.int64)
    clips_pinned = [clip_pin[soff[i]:soff[i + 1]] for i in range(N_UTTS)]

    # ---- device-resident Griffin-Lim through the C ABI ----
    cfg = _runtime._make_config(NFFT, WIN, HOP, 'f32')
    plan = ctypes.c_void_p()
    _lib.check(lib.sstts_gl_plan_create(ctypes.byref(cfg), N_UTTS,
                                        foff.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.byref(plan)))
    ws = torch.empty(int(lib.sstts_gl_workspace_bytes(plan)), dtype=torch.uint8, device=dev)
    n_samples = int(lib.sstts_gl_total_samples(plan))
    wav_dev = torch.empty(n_samples, dtype=torch.float32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def gl_step(n_iter=GL_ITERS):
        # seeded random initial phase drawn inside the synthesis launch (what spectrograms_to_wavs does)
        _lib.check(lib.sstts_griffin_lim_seeded(plan, ctypes.c_void_p(mag_dev.data_ptr()), ctypes.c_uint64(1234), 0,
                                                n_iter, ctypes.c_void_p(ws.data_ptr()),
                                                ctypes.c_void_p(wav_dev.data_ptr()), None, stream))

    sampler = ClockSampler(H.local_rank)
    if H.rank == 0:
        sampler.start()
    gl_ms = H.timed(gl_step, args.steps, args.warmup)
    clocks = sampler.stop() if H.rank == 0 else None
    total_audio = H.sum_over_ranks(audio_out_s)
    gl_value = total_audio * args.steps / (gl_ms / 1000.0)

    # dominant kernel (gl_step_kernel, 50 launches per step): average launch duration from the
    # CUDA-event time of the full step minus the same call with n_iter = 0 (synth + finalize only)
    base_ms = H.timed(lambda: gl_step(0), args.steps, 1)
    iter_ms = max(1e-9, (gl_ms - base_ms) / (args.steps * GL_ITERS))
    gl_alg_bytes = GL_BYTES_PER_FRAME_ITER * total_frames
    gl_achieved = gl_alg_bytes / (iter_ms / 1000.0) / 1e9

    # ---- e2e Griffin-Lim through the public API (host numpy in, host numpy out) ----
    def gl_e2e():
        synthesis.spectrograms_to_wavs(mags_pageable, WIN, HOP, NFFT, GL_ITERS, seed=1234)

    def gl_e2e_pinned():
        synthesis.spectrograms_to_wavs(mags_pinned, WIN, HOP, NFFT, GL_ITERS, seed=1234)

    def alloc_counters():
        st = torch.cuda.memory_stats(dev)
        out = {'device_alloc': st.get('num_device_alloc', 0), 'device_free': st.get('num_device_free', 0)}
        try:
            hs = torch.cuda.host_memory_stats()
            out['host_alloc'] = hs.get('num_host_alloc', 0)
            out['host_free'] = hs.get('num_host_free', 0)
        except Exception:
            pass
        return out

    e2e_steps = max(1, min(args.steps, 5))
    e2e_warm = max(3, args.warmup)
    # steady state first: the sub-batch plans, the pinned staging buffers and the cached CUDA graphs of the
    # sub-batches (captured at the fourth call with the same buffers) all exist before anything is timed
    for _ in range(8):
        gl_e2e()
    c0 = alloc_counters()
    gl_e2e_ms, gl_e2e_calls = H.timed(gl_e2e, e2e_steps, e2e_warm, per_call=True)
    c1 = alloc_counters()
    e2e_allocs = {k: c1[k] - c0[k] for k in c1}      # cudaMalloc / cudaHostAlloc calls inside the e2e loop
    gl_e2e_value = total_audio * e2e_steps / (gl_e2e_ms / 1000.0)
    for _ in range(6):
        gl_e2e_pinned()
    gl_e2e_pin_ms = H.timed(gl_e2e_pinned, e2e_steps, e2e_warm)
    h2d = total_frames * N_BINS * 4
    d2h = n_samples * 4
    lib.sstts_gl_plan_destroy(plan)

    # ---- feature pipeline (configs[1]) ----
    feat = {}
    wav_in = torch.from_numpy(np.concatenate(clips)).to(dev)
    for prec in ('f64', 'f32'):
        fcfg = _runtime._make_config(NFFT, WIN, HOP, prec, SR, 80, 0, 8000)
        fplan = ctypes.c_void_p()
        _lib.check(lib.sstts_feat_plan_create(ctypes.byref(fcfg), N_UTTS,
                                              soff.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 5,
                                              ctypes.byref(fplan)))
        rows = int(lib.sstts_feat_total_rows(fplan))
        lin = torch.empty((rows, N_BINS), dtype=torch.float32, device=dev)
        mel = torch.empty((rows, 80), dtype=torch.float32, device=dev)
        out = _lib.FeatOutputs()
        out.lin_db_dev, out.mel_db_dev = lin.data_ptr(), mel.data_ptr()
        out.normalize = 1
        out.lin_ref_db, out.lin_max_db, out.mel_ref_db, out.mel_max_db = CONSTS
        out.mel_power = 1.0

        def feat_step():
            _lib.check(lib.sstts_stft_features(fplan, ctypes.c_void_p(wav_in.data_ptr()), ctypes.byref(out), stream))

        f_ms = H.timed(feat_step, args.steps * 5, args.warmup)
        per_launch_ms = f_ms / (args.steps * 5)
        alg = 4 * int(soff[-1]) + FEAT_BYTES_PER_FRAME * total_frames
        achieved = alg / (per_launch_ms / 1000.0) / 1e9
        roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peak_gbs, 'unit': 'GB/s', 'frac': achieved / peak_gbs,
                'traffic': measured_ncu(FEAT_KERNELS[prec], 'dram_bytes_per_launch'), 'kernel': 'stft_feature_kernel',
                'algorithmic_bytes_per_launch': alg, 'ms_per_launch': per_launch_ms}
        roof.update(ncu_block(FEAT_KERNELS[prec]))
        feat[prec] = {'value': H.sum_over_ranks(audio_in_s) / (per_launch_ms / 1000.0), 'unit': 'audio-s/s',
                      'ms_per_step': per_launch_ms, 'roofline': roof}
        lib.sstts_feat_plan_destroy(fplan)
        del lin, mel

    # ---- statistics kernel (datasets/statistics.py:31-34: n_fft 1024, hop 256, win 1024, fmax sr // 2) ----
    scfg = _runtime._make_config(1024, 1024, 256, 'f64', SR, 80, 0, SR // 2)
    splan = ctypes.c_void_p()
    _lib.check(lib.sstts_feat_plan_create(ctypes.byref(scfg), N_UTTS,
                                          soff.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 1, ctypes.byref(splan)))
    mm = torch.empty((N_UTTS, 4), dtype=torch.float64, device=dev)
    sout = _lib.FeatOutputs()
    sout.minmax_dev = mm.data_ptr()
    sout.mel_power = 1.0

    def stats_step():
        _lib.check(lib.sstts_stft_features(splan, ctypes.c_void_p(wav_in.data_ptr()), ctypes.byref(sout), stream))

    s_ms = H.timed(stats_step, args.steps * 5, args.warmup) / (args.steps * 5)
    s_alg = 4 * int(soff[-1]) + 16 * N_UTTS
    s_frames = sum(1 + len(c) // 256 for c in clips)
    stats_kernel = measured_ncu('statistics_kernel_name', 'name') or 'stft_feature_kernel (statistics geometry)'
    statistics = {'metric': 'statistics_audio_sec_per_sec', 'value': H.sum_over_ranks(audio_in_s) / (s_ms / 1000.0),
                  'unit': 'audio-s/s', 'ms_per_step': s_ms, 'frames_per_gpu': s_frames, 'dtype': 'f64',
                  'workload': 'per-clip [min lin, max lin, min mel, max mel] dB of the 256 clips, n_fft 1024 / hop 256',
                  'roofline': dict({'bound': 'hbm', 'achieved': s_alg / (s_ms / 1000.0) / 1e9, 'peak': peak_gbs,
                                    'unit': 'GB/s', 'frac': s_alg / (s_ms / 1000.0) / 1e9 / peak_gbs,
                                    'algorithmic_bytes_per_launch': s_alg, 'kernel': stats_kernel,
                                    'note': 'transform-bound by construction (4 bytes read per sample, 16 bytes '
                                            'written per clip): the FP64 pipe figure is the relevant one'},
                                   **ncu_block('statistics'))}
    lib.sstts_feat_plan_destroy(splan)
    del wav_in, mm

    # ---- Griffin-Lim at n_fft 1024 / hop 256 / win 1024, 25 iterations (audio/effects.py:71-86): the native
    # half-warp transform, device-resident through the C ABI like `value` ----
    fb1 = _runtime.stft_features_batch(clips, 1024, 256, 1024, want_spec=True, precision='f64', keep_on_device=True)
    mag1 = fb1.spec.abs().contiguous()
    del fb1
    frames1 = [1 + len(c) // 256 for c in clips]
    foff1 = np.concatenate([[0], np.cumsum(frames1)]).astype(np.int64)
    cfg1 = _runtime._make_config(1024, 1024, 256, 'f32')
    plan1 = ctypes.c_void_p()
    _lib.check(lib.sstts_gl_plan_create(ctypes.byref(cfg1), N_UTTS,
                                        foff1.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.byref(plan1)))
    ws1 = torch.empty(int(lib.sstts_gl_workspace_bytes(plan1)), dtype=torch.uint8, device=dev)
    wav1 = torch.empty(int(lib.sstts_gl_total_samples(plan1)), dtype=torch.float32, device=dev)

    def gl1_step(n_iter=25):
        _lib.check(lib.sstts_griffin_lim_seeded(plan1, ctypes.c_void_p(mag1.data_ptr()), ctypes.c_uint64(1234), 0,
                                                n_iter, ctypes.c_void_p(ws1.data_ptr()),
                                                ctypes.c_void_p(wav1.data_ptr()), None, stream))

    g1_ms = H.timed(gl1_step, args.steps, args.warmup) / args.steps
    g1_base = H.timed(lambda: gl1_step(0), args.steps, 1) / args.steps
    audio1_s = sum(256 * (t - 1) for t in frames1) / SR
    gl_1024 = {'metric': 'griffin_lim_25it_n_fft_1024_audio_sec_per_sec',
               'value': H.sum_over_ranks(audio1_s) / (g1_ms / 1000.0), 'unit': 'audio-s/s', 'ms_per_step': g1_ms,
               'ms_per_iteration_launch': (g1_ms - g1_base) / 25, 'frames_per_gpu': int(foff1[-1]), 'dtype': 'f32',
               'workload': 'the 256 clips at n_fft 1024 / hop 256 / win 1024, 25 iterations (time_stretch geometry), '
                           'device-resident; kernel: gl_step_kernel<float, NativeGeom1024<1024, 256>, ...> '
                           '(two frames per warp)'}
    # same byte model as the headline kernel: spectrum out and back (513 x 8 B x 2), |S| row, waveform in / out
    g1_alg = (513 * 8 * 2 + 513 * 4 + 256 * 4 * 2) * int(foff1[-1])
    g1_iter_ms = max(1e-9, (g1_ms - g1_base) / 25)
    g1_kernel = 'gl_step_kernel<float, NativeGeom1024<1024, 256>, 8, 0, 0, 0>'
    gl_1024['roofline'] = dict({'bound': 'hbm', 'achieved': g1_alg / (g1_iter_ms / 1000.0) / 1e9, 'peak': peak_gbs,
                                'unit': 'GB/s', 'frac': g1_alg / (g1_iter_ms / 1000.0) / 1e9 / peak_gbs,
                                'traffic': measured_ncu(g1_kernel, 'dram_bytes_per_launch'), 'kernel': g1_kernel,
                                'algorithmic_bytes_per_launch': g1_alg, 'ms_per_launch': g1_iter_ms}, **ncu_block(g1_kernel))
    lib.sstts_gl_plan_destroy(plan1)
    del mag1, ws1, wav1

    def feat_e2e():
        feat_api.features_batch(clips, NFFT, HOP, WIN, SR, 80, 0, 8000, *CONSTS, reduction=5)

    def feat_e2e_pinned():
        feat_api.features_batch(clips_pinned, NFFT, HOP, WIN, SR, 80, 0, 8000, *CONSTS, reduction=5)

    f_e2e_ms, f_e2e_calls = H.timed(feat_e2e, e2e_steps, e2e_warm, per_call=True)
    f_e2e_pin_ms = H.timed(feat_e2e_pinned, e2e_steps, e2e_warm)
    rows5 = sum(-(-t // 5) * 5 for t in frames)
    audio_in_total = H.sum_over_ranks(audio_in_s)
    feat_e2e_line = {'value': audio_in_total * e2e_steps / (f_e2e_ms / 1000.0), 'unit': 'audio-s/s',
                     'h2d_bytes_per_step': int(sum(len(c) for c in clips)) * 4,
                     'd2h_bytes_per_step': rows5 * (N_BINS + 80) * 4, 'inputs': 'pageable host numpy arrays',
                     'ms_per_step': f_e2e_ms / e2e_steps, 'per_call_ms_rank0': f_e2e_calls,
                     'pinned_inputs_value': audio_in_total * e2e_steps / (f_e2e_pin_ms / 1000.0)}

    pack_gbs = host_pack_bandwidth(H)
    link_gbs = host_link_bandwidth(H)
    roofline = {'bound': 'hbm', 'achieved': gl_achieved, 'peak': peak_gbs, 'unit': 'GB/s',
                'frac': gl_achieved / peak_gbs, 'traffic': measured_ncu(GL_KERNEL, 'dram_bytes_per_launch'),
                'kernel': 'gl_step_kernel', 'peak_source': peak_src, 'ms_per_launch': iter_ms,
                'algorithmic_bytes_per_launch': gl_alg_bytes,
                'how': '(CUDA-event time of the 50-iteration call - same call with 0 iterations) / 50'}
    roofline.update(ncu_block(GL_KERNEL))
    cfg_line = workload_config()
    cfg_line.update({'frames_per_gpu': total_frames, 'audio_seconds_per_gpu': audio_out_s})
    return {
        'metric': 'griffin_lim_50it_audio_sec_per_sec', 'value': gl_value, 'unit': 'audio-s/s',
        'n_gpus': H.world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': gl_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': cfg_line,
        'e2e': {'value': gl_e2e_value, 'unit': 'audio-s/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': gl_e2e_ms / e2e_steps,
                'inputs': 'pageable host numpy arrays (staged through pooled pinned buffers), outputs numpy',
                'per_call_ms_rank0': gl_e2e_calls, 'allocator_calls_in_loop': e2e_allocs,
                'pinned_inputs_value': total_audio * e2e_steps / (gl_e2e_pin_ms / 1000.0)},
        'gpu_launches': args.steps * (GL_ITERS + 2),
        'roofline': roofline,
        'clocks': clocks,
        'host': {'cores': len(os.sched_getaffinity(0)), 'io_threads_per_rank': H.io_threads,
                 'pack_gbs_all_ranks': pack_gbs, **link_gbs,
                 'note': 'pageable-input e2e stages 4.1 KB per frame through pinned memory: at most pack_gbs_all_ranks / '
                         '(329.5 KB per audio-second) audio-s/s on this host, whatever the number of GPUs; the feature / '
                         'corpus paths download 354 KB per audio-second: at most d2h_gbs_all_ranks / 354 KB'},
        'features': {'metric': 'feature_audio_sec_per_sec',
                     'workload': 'BASELINE configs[1]: STFT -> linear + 80-mel dB-normalised features, '
                                 '256 ragged clips per GPU',
                     'value': feat['f64']['value'], 'unit': 'audio-s/s',
                     'f64': feat['f64'], 'f32_fast': feat['f32'], 'e2e': feat_e2e_line,
                     'dtype': 'f64 (default; f32_fast is the opt-in float32 transform)'},
        'statistics': statistics,
        'griffin_lim_n_fft_1024': gl_1024,
    }


def bench_latency(H, args):
    """Single-utterance latency of the serve loop (tacotron/serve.py:39-86): spectrogram_to_wav on one
    1000-frame utterance (decoder.maximum_iterations, tacotron/params/model.py:108), host numpy in / out,
    alone and from 6 concurrent threads (:69-72).  Rank 0 only."""
    from concurrent.futures import ThreadPoolExecutor
    from single_speaker_tts_b200.audio import synthesis
    rng = np.random.default_rng(7)
    T = 1000
    mag = (rng.random((T, N_BINS), dtype=np.float32) ** 4 * 20.0).T          # (1025, T) view of a (T, 1025) array
    audio_s = HOP * (T - 1) / SR

    def dropin():
        # the reference's call: initial phase from numpy's global RandomState (audio/synthesis.py:85), whose
        # 1,025,000 draws cost ~4.4 ms on the host and are serialised by the RandomState lock across threads
        t0 = time.perf_counter()
        synthesis.spectrogram_to_wav(mag, WIN, HOP, NFFT, GL_ITERS)
        return (time.perf_counter() - t0) * 1000.0

    def seeded():
        # same call with the phase drawn on the device (one seed taken from numpy's global stream)
        t0 = time.perf_counter()
        synthesis.spectrogram_to_wav(mag, WIN, HOP, NFFT, GL_ITERS, device_phase=True)
        return (time.perf_counter() - t0) * 1000.0

    def pct(v, q):
        return v[min(len(v) - 1, int(round(q * (len(v) - 1))))]

    out = {'workload': 'spectrogram_to_wav, one 1000-frame utterance (12.46 s of audio), 50 iterations, host numpy '
                       'in / out (tacotron/serve.py:39-86); six_callers = ThreadPool(6) as in :69-72',
           'audio_seconds': audio_s}
    for name, one in (('dropin_numpy_rng_phase', dropin), ('device_phase', seeded)):
        for _ in range(5):
            one()
        n = 40
        alone = sorted(one() for _ in range(n))
        with ThreadPoolExecutor(max_workers=6) as ex:
            list(ex.map(lambda _: one(), range(12)))
            t0 = time.perf_counter()
            conc = sorted(ex.map(lambda _: one(), range(60)))
            wall = time.perf_counter() - t0
        out[name] = {'single_caller_ms': {'p50': pct(alone, 0.5), 'p99': pct(alone, 0.99), 'min': alone[0], 'n': n},
                     'six_callers_ms': {'p50': pct(conc, 0.5), 'p99': pct(conc, 0.99), 'n': 60,
                                        'throughput_audio_s_per_s': 60 * audio_s / wall}}
    return out


def bench_corpus(H, args, n_total=13100):
    """configs[3]: corpus pass over 13,100 synthetic LJSpeech-length clips sharded by clip (balanced by
    frames): per-clip dB statistics (n_fft 1024 / hop 256) + the all-reduce -> corpus constants, then the
    feature pre-calculation (2048 / 275 / 1102, reduction 5) with those constants.  Host numpy in, host
    numpy out; STRONG scaling (fixed total work); the collective is inside the timed region."""
    torch = H.torch
    from single_speaker_tts_b200 import distributed
    from single_speaker_tts_b200.synthetic import ClipPlan
    plan = ClipPlan(n_total, seed=3, kind='ljspeech', pool=32)
    shards = distributed.shard_by_cost(plan.frames(HOP), H.world)
    mine = shards[H.rank]
    wavs = plan.clips(mine)
    audio_total = float(plan.lengths.sum()) / SR
    frames_mine = int(plan.frames(HOP)[mine].sum())
    state = {}

    def one_pass():
        mean4, n_rows = distributed.corpus_pass(wavs, mine, n_total, SR, NFFT, HOP, WIN, 80, 0, 8000, reduction=5,
                                                chunk_samples=args.corpus_chunk)
        state['mean4'], state['rows'] = mean4, n_rows

    steps = max(1, min(args.steps, 2))
    ms = H.timed(one_pass, steps, 1)
    mean4 = state['mean4']
    out_bytes = state['rows'] * (N_BINS + 80) * 4
    return {'metric': 'corpus_pass_audio_sec_per_sec', 'value': audio_total * steps / (ms / 1000.0),
            'unit': 'audio-s/s', 'scaling': 'strong', 'steps': steps, 'ms_per_step': ms / steps, 'dtype': 'f64',
            'workload': 'BASELINE configs[3]: statistics (1024/256) + all-reduce + feature precalc (2048/275/1102, '
                        'r=5) over %d LJSpeech-length clips, end to end (pageable numpy in, numpy out)' % n_total,
            'sharding': 'by clip, balanced by frames; one all-reduce of the (n_clips, 4) table inside the timed region',
            'frames_rank0': frames_mine, 'd2h_bytes_rank0': out_bytes,
            'h2d_bytes_rank0': int(sum(len(w) for w in wavs)) * 4,
            'corpus_statistics': {'linear_mag_max_db': mean4[0], 'linear_ref_db': mean4[1],
                                  'mel_mag_max_db': mean4[2], 'mel_mag_ref_db': mean4[3]}}


def bench_gl_sharded(H, args, n_total=4096, n_iter=100):
    """configs[4]: Griffin-Lim, 100 iterations, 4,096 ragged utterances sharded by utterance (balanced by
    frames, no collective), end to end through spectrograms_to_wavs; STRONG scaling."""
    from single_speaker_tts_b200 import _runtime, distributed
    from single_speaker_tts_b200.audio import synthesis
    from single_speaker_tts_b200.synthetic import ClipPlan
    plan = ClipPlan(n_total, seed=4, kind='uniform', pool=32)
    frames = plan.frames(HOP)
    mine = distributed.shard_by_cost(frames, H.world)[H.rank]
    audio_total = float((HOP * (frames - 1)).sum()) / SR
    mags = []
    for s in range(0, len(mine), 256):       # |STFT| of this rank's clips via the feature kernel
        idx = mine[s:s + 256]
        fb = _runtime.stft_features_batch(plan.clips(idx), NFFT, HOP, WIN, want_spec=True, precision='f64')
        for i in range(len(idx)):
            mags.append(np.abs(fb.rows(fb.spec, i)).T)
        del fb

    def one_pass():
        for s in range(0, len(mags), 512):
            synthesis.spectrograms_to_wavs(mags[s:s + 512], WIN, HOP, NFFT, n_iter, seed=1234 + s)

    steps = max(1, min(args.steps, 2))
    ms = H.timed(one_pass, steps, 1)
    return {'metric': 'griffin_lim_100it_audio_sec_per_sec', 'value': audio_total * steps / (ms / 1000.0),
            'unit': 'audio-s/s', 'scaling': 'strong', 'steps': steps, 'ms_per_step': ms / steps, 'dtype': 'f32',
            'workload': 'BASELINE configs[4]: Griffin-Lim 100 it over %d ragged utterances, end to end '
                        '(pageable numpy in / out)' % n_total,
            'sharding': 'by utterance, balanced by frames; no collective',
            'frames_rank0': int(frames[mine].sum())}


def run_ours(args):
    H = Harness()
    line = None
    # The sharded (strong-scaling) workloads run first: behind the 256-utterance workloads in the same process the
    # corpus pass was 15-25 % slower at 2 and 4 GPUs (644 vs 492 ms per pass at 2 GPUs; allocator caches ruled out;
    # see DESIGN.md 7 for the unverified explanation: pinned buffers allocated before vs after the first collective),
    # while this order leaves the other numbers unchanged.  BENCH_SHARDED_LAST=1: the old order.
    sharded_first = args.workload == 'all' and not args.no_sharded and os.environ.get('BENCH_SHARDED_LAST') != '1'
    if sharded_first:
        corpus = bench_corpus(H, args, args.clips or 13100)
        gl4096 = bench_gl_sharded(H, args, args.clips or 4096)
    if args.workload in ('all', 'gl256'):
        line = bench_primary(H, args)
    if args.workload == 'all':
        if H.rank == 0 and not args.no_latency:
            line['latency'] = bench_latency(H, args)
        H.barrier()
        if not args.no_sharded:
            if not sharded_first:
                corpus = bench_corpus(H, args, args.clips or 13100)
                gl4096 = bench_gl_sharded(H, args, args.clips or 4096)
            line['corpus'], line['gl4096'] = corpus, gl4096
        if H.rank == 0 and H.world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baselines_in_subprocess()
        else:
            line['cpu_baseline'] = None
    elif args.workload == 'corpus':
        line = bench_corpus(H, args, args.clips or 13100)
        line.update({'n_gpus': H.world, 'higher_is_better': True, 'data': 'synthetic'})
    elif args.workload == 'gl4096':
        line = bench_gl_sharded(H, args, args.clips or 4096)
        line.update({'n_gpus': H.world, 'higher_is_better': True, 'data': 'synthetic'})
    elif args.workload == 'latency':
        line = bench_latency(H, args)
    if H.rank == 0:
        print(json.dumps(line), flush=True)
    H.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-sharded', action='store_true', help='skip the corpus / gl4096 sub-results')
    ap.add_argument('--no-latency', action='store_true', help='skip the single-utterance latency sub-result')
    ap.add_argument('--workload', default='all', choices=['all', 'gl256', 'corpus', 'gl4096', 'latency'],
                    help='all = the driver contract line with every sub-result (default); gl256 = configs[1]+[2] '
                         'only; corpus = configs[3]; gl4096 = configs[4]; latency = single utterance')
    ap.add_argument('--clips', type=int, default=0, help='override the clip count of corpus / gl4096')
    ap.add_argument('--corpus-chunk', type=int, default=16 << 20, help='samples per device batch of the corpus pass')
    ap.add_argument('--ref-items', type=int, default=0,
                    help='--impl reference: utterances per step (default 2 x host cores)')
    ap.add_argument('--cpu-baseline-only', action='store_true',
                    help='print the CPU baselines (oracle on host cores) as one JSON object and exit')
    args = ap.parse_args()
    # a benchmark that hangs is worse than one that fails: after 15 minutes dump every thread's stack and exit
    import faulthandler
    faulthandler.dump_traceback_later(900, exit=True)
    if args.cpu_baseline_only:
        print(json.dumps(cpu_baselines()), flush=True)
        return
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
