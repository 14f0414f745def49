#!/usr/bin/env python
"""Benchmark of the audio hot path (BASELINE.json): batched 50-iteration Griffin-Lim and the
STFT feature pipeline, in audio-seconds per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" is one pass of the hot path over one batch: BASELINE configs[2] -- Griffin-Lim, 50
iterations, 256 ragged synthetic utterances (1-10 s, 22.05 kHz) per GPU -- which is the primary
metric; the feature pipeline (configs[1], the same 256 clips) is measured in the same run and
reported under "features".  `value` is device-resident throughput (inputs in HBM when the timed
region starts, launched through the C ABI), `e2e` goes through the public Python API with host
numpy buffers (pinned staging, H2D, kernels, D2H inside the timed region).  Weak scaling: every
rank owns its own 256 utterances; no data-path collective (utterances are independent).

--impl reference times the reference's CPU implementation of the same path (the numpy oracle
restating librosa 0.6 + audio/synthesis.py; the reference itself cannot be installed: librosa and
TensorFlow 1.8 are unavailable offline) on all host cores on a bounded sample of the workload.
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


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_ncu(kernel, key):
    """One metric of `kernel` from the committed ncu --set full capture (profiles/r1_traffic.json), or None."""
    path = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    try:
        with open(path) as f:
            return json.load(f)[kernel][key]
    except (OSError, KeyError, ValueError):
        return None


def measured_traffic(kernel):
    """dram__bytes_read + dram__bytes_write per launch of `kernel` from the committed ncu
    --set full capture of this workload shape (profiles/r1_traffic.json), or None."""
    path = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    try:
        with open(path) as f:
            return json.load(f)[kernel]['dram_bytes_per_launch']
    except (OSError, KeyError, ValueError):
        return None


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
def _oracle_gl_worker(args):
    os.environ['OMP_NUM_THREADS'] = '1'
    os.environ['MKL_NUM_THREADS'] = '1'
    sys.path.insert(0, ROOT)
    from oracle import reference_audio as ra
    mag, seed = args
    angles = np.exp(2j * np.pi * np.random.RandomState(seed).rand(*mag.shape))
    wav = ra.spectrogram_to_wav(mag, WIN, HOP, NFFT, GL_ITERS, angles=angles)
    return len(wav)


def oracle_sample(n_items, seconds, seed):
    """Bounded sample of the workload: n_items synthetic clips of `seconds` s -> |STFT|."""
    from oracle import librosa_compat as lc
    from single_speaker_tts_b200.synthetic import speech_like_clip
    rng = np.random.default_rng(seed)
    mags = []
    for _ in range(n_items):
        x = speech_like_clip(int(seconds * SR), rng)
        mags.append(np.abs(lc.stft(x, NFFT, HOP, WIN)))
    return mags


def time_oracle_gl(pool, mags, seed0):
    t0 = time.perf_counter()
    lens = pool.map(_oracle_gl_worker, [(m, seed0 + i) for i, m in enumerate(mags)])
    dt = time.perf_counter() - t0
    return sum(lens) / SR, dt


def run_reference(args):
    """--impl reference: the oracle (kind "port": the reference's numpy/librosa code restated; the
    reference itself is not installable offline) on all host cores, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    sample_s = 3.0
    mags = oracle_sample(cores, sample_s, seed=2)
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            time_oracle_gl(pool, mags[:cores], 1000)
        audio, total = 0.0, 0.0
        for s in range(args.steps):
            a, dt = time_oracle_gl(pool, mags, 1000 + s)
            audio += a
            total += dt
    value = audio / total
    sample = '{} utterances x {:.0f} s per step, one process per core, {} iterations'.format(
        cores, sample_s, GL_ITERS)
    line = {
        'impl': 'reference', 'metric': 'griffin_lim_50it_audio_sec_per_sec', 'value': value,
        'unit': 'audio-s/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1000.0 * total / max(1, args.steps), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[2]: Griffin-Lim 50 it, ragged synthetic utterances '
                               '(bounded CPU sample)', 'n_fft': NFFT, 'win': WIN, 'hop': HOP},
        'cpu_baseline': {'value': value, 'unit': 'audio-s/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'audio-s/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from single_speaker_tts_b200 import _lib, _runtime
    from single_speaker_tts_b200.audio import features as feat_api
    from single_speaker_tts_b200.audio import synthesis
    from single_speaker_tts_b200.synthetic import make_clips

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    peak_gbs, peak_src = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed(fn, steps, warmup):
        """warmup untimed calls, then `steps` calls between barriers; CUDA-event time, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms)

    def timed_calls(fn, steps, warmup):
        """Like timed(), plus this rank's per-call times (an event after every call): the end-to-end
        calls synchronise internally, so the per-call list shows host-side hiccups (page-locked
        allocations, other tenants of the host) that the total hides."""
        for _ in range(warmup):
            fn()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[steps])
        per_call = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)]
        barrier()
        return max_over_ranks(ms), per_call

    # ---- workload: configs[1]/[2], 256 ragged clips per rank (weak scaling) ----
    # weak scaling: every rank owns the same 256-clip workload (same seed), so per-GPU work is fixed exactly
    clips = make_clips(N_UTTS, seed=1, pool=16)
    audio_in_s = sum(len(c) for c in clips) / SR
    frames = [1 + len(c) // HOP for c in clips]
    total_frames = sum(frames)
    audio_out_s = sum(HOP * (t - 1) for t in frames) / SR

    # magnitudes |STFT(clip)| produced on the device by our own feature kernel (float64 transform)
    fb = _runtime.stft_features_batch(clips, NFFT, HOP, WIN, want_spec=True, precision='f64',
                                      keep_on_device=True)
    mag_dev = fb.spec.abs().contiguous()                      # (sum T, 1025) float32, frame-major
    del fb
    # host copies of the inputs: page-locked (the e2e contract's "pinned host memory": uploaded with no
    # staging copy) and ordinary pageable numpy memory (what a TF session hands over; staged through
    # pinned buffers by worker threads) -- both are timed, `e2e` reports the pinned one
    import single_speaker_tts_b200 as pkg
    mag_host = pkg.pinned_empty(tuple(mag_dev.shape))
    torch.from_numpy(mag_host).copy_(mag_dev)
    mag_pageable = mag_host.copy()
    foff = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
    mags_host = [mag_host[foff[i]:foff[i + 1]].T for i in range(N_UTTS)]   # (1025, T) views, like spec.T
    mags_pageable = [mag_pageable[foff[i]:foff[i + 1]].T for i in range(N_UTTS)]
    clip_pin = pkg.pinned_empty((sum(len(c) for c in clips),))
    clip_pin[:] = np.concatenate(clips)
    soff_c = np.concatenate([[0], np.cumsum([len(c) for c in clips])])
    clips_pinned = [clip_pin[soff_c[i]:soff_c[i + 1]] for i in range(N_UTTS)]

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

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    gl_ms = timed(gl_step, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    total_audio = sum_over_ranks(audio_out_s)
    gl_value = total_audio * args.steps / (gl_ms / 1000.0)

    # dominant kernel (gl_step_kernel, 50 launches per step): average launch duration from the
    # CUDA-event time of the full step minus the same call with n_iter = 0 (synth + finalize only)
    base_ms = timed(lambda: gl_step(0), args.steps, 1)
    iter_ms = max(1e-9, (gl_ms - base_ms) / (args.steps * GL_ITERS))
    gl_alg_bytes = GL_BYTES_PER_FRAME_ITER * total_frames
    gl_achieved = gl_alg_bytes / (iter_ms / 1000.0) / 1e9

    # ---- e2e Griffin-Lim through the public API (host numpy in, host numpy out) ----
    def gl_e2e():
        synthesis.spectrograms_to_wavs(mags_host, WIN, HOP, NFFT, GL_ITERS, seed=1234)

    def gl_e2e_pageable():
        synthesis.spectrograms_to_wavs(mags_pageable, WIN, HOP, NFFT, GL_ITERS, seed=1234)

    e2e_steps = max(1, min(args.steps, 5))
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

    gl_e2e()
    c0 = alloc_counters()
    gl_e2e_ms, gl_e2e_calls = timed_calls(gl_e2e, e2e_steps, max(1, args.warmup))
    c1 = alloc_counters()
    e2e_allocs = {k: c1[k] - c0[k] for k in c1}      # cudaMalloc / cudaHostAlloc calls inside the e2e loop
    gl_e2e_value = total_audio * e2e_steps / (gl_e2e_ms / 1000.0)
    gl_e2e_pg_ms = timed(gl_e2e_pageable, e2e_steps, max(1, args.warmup))
    h2d = total_frames * N_BINS * 4
    d2h = n_samples * 4

    # ---- feature pipeline (configs[1]) ----
    consts = (35.66, 100.0, 6.02, 99.89)
    feat = {}
    for prec in ('f64', 'f32'):
        fcfg = _runtime._make_config(NFFT, WIN, HOP, prec, SR, 80, 0, 8000)
        fplan = ctypes.c_void_p()
        soff = np.concatenate([[0], np.cumsum([len(c) for c in clips])]).astype(np.int64)
        _lib.check(lib.sstts_feat_plan_create(ctypes.byref(fcfg), N_UTTS,
                                              soff.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 5,
                                              ctypes.byref(fplan)))
        rows = int(lib.sstts_feat_total_rows(fplan))
        wav_in = torch.from_numpy(np.concatenate(clips)).to(dev)
        lin = torch.empty((rows, N_BINS), dtype=torch.float32, device=dev)
        mel = torch.empty((rows, 80), dtype=torch.float32, device=dev)
        out = _lib.FeatOutputs()
        out.lin_db_dev, out.mel_db_dev = lin.data_ptr(), mel.data_ptr()
        out.normalize = 1
        out.lin_ref_db, out.lin_max_db, out.mel_ref_db, out.mel_max_db = consts
        out.mel_power = 1.0

        def feat_step():
            _lib.check(lib.sstts_stft_features(fplan, ctypes.c_void_p(wav_in.data_ptr()), ctypes.byref(out), stream))

        f_ms = timed(feat_step, args.steps * 5, args.warmup)
        per_launch_ms = f_ms / (args.steps * 5)
        alg = 4 * int(soff[-1]) + FEAT_BYTES_PER_FRAME * total_frames
        feat[prec] = {
            'value': sum_over_ranks(audio_in_s) / (per_launch_ms / 1000.0), 'unit': 'audio-s/s',
            'ms_per_step': per_launch_ms,
            'roofline': {'bound': 'hbm', 'achieved': alg / (per_launch_ms / 1000.0) / 1e9, 'peak': peak_gbs,
                         'unit': 'GB/s', 'frac': alg / (per_launch_ms / 1000.0) / 1e9 / peak_gbs,
                         'traffic': measured_traffic('stft_feature_kernel<%s, StaticGeom<1102, 275, 2048>, %d, 1>'
                                                     % (('double', 4) if prec == 'f64' else ('float', 8)))},
        }
        lib.sstts_feat_plan_destroy(fplan)
        del lin, mel, wav_in

    def feat_e2e():
        feat_api.features_batch(clips_pinned, NFFT, HOP, WIN, SR, 80, 0, 8000, *consts, reduction=5)

    def feat_e2e_pageable():
        feat_api.features_batch(clips, NFFT, HOP, WIN, SR, 80, 0, 8000, *consts, reduction=5)

    f_e2e_ms, f_e2e_calls = timed_calls(feat_e2e, e2e_steps, max(1, args.warmup))
    f_e2e_pg_ms = timed(feat_e2e_pageable, e2e_steps, max(1, args.warmup))
    rows5 = sum(-(-t // 5) * 5 for t in frames)
    feat_e2e = {'value': sum_over_ranks(audio_in_s) * e2e_steps / (f_e2e_ms / 1000.0), 'unit': 'audio-s/s',
                'h2d_bytes_per_step': int(sum(len(c) for c in clips)) * 4,
                'd2h_bytes_per_step': rows5 * (N_BINS + 80) * 4, 'inputs': 'pinned host numpy arrays',
                'ms_per_step': f_e2e_ms / e2e_steps, 'per_call_ms_rank0': f_e2e_calls,
                'pageable_inputs_value': sum_over_ranks(audio_in_s) * e2e_steps / (f_e2e_pg_ms / 1000.0)}

    # ---- CPU baseline on rank 0 (oracle, bounded sample) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = len(os.sched_getaffinity(0))
        use = min(6, cores)                   # tacotron/params/inference.py:34: 6 synthesis workers
        mags = oracle_sample(use, 3.0, seed=2)
        with mp.get_context('fork').Pool(use) as pool:
            a, dt = time_oracle_gl(pool, mags, 1000)
        cpu = {'value': a / dt, 'unit': 'audio-s/s', 'cores': use, 'kind': 'port',
               'sample': '{} utterances x 3 s, {} iterations, pool of {} processes (host has {} cores)'.format(
                   use, GL_ITERS, use, cores)}

    lib.sstts_gl_plan_destroy(plan)
    if rank == 0:
        line = {
            'metric': 'griffin_lim_50it_audio_sec_per_sec', 'value': gl_value, 'unit': 'audio-s/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': gl_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'BASELINE configs[2]: Griffin-Lim 50 it over 256 ragged synthetic utterances '
                                   '(1-10 s, 22.05 kHz) per GPU', 'n_fft': NFFT, 'win': WIN, 'hop': HOP,
                       'n_utterances_per_gpu': N_UTTS, 'frames_per_gpu': total_frames,
                       'audio_seconds_per_gpu': audio_out_s, 'l2': 'inputs_exceed_l2 (|S| 4.1 KB/frame + '
                       'waveform state >> 126 MB)', 'sharding': 'by utterance, no collective; every rank runs the same 256-clip set'},
            'e2e': {'value': gl_e2e_value, 'unit': 'audio-s/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': gl_e2e_ms / e2e_steps,
                    'inputs': 'pinned host numpy arrays (pkg.pinned_empty), outputs numpy in pinned memory',
                    'per_call_ms_rank0': gl_e2e_calls, 'allocator_calls_in_loop': e2e_allocs,
                    'pageable_inputs_value': total_audio * e2e_steps / (gl_e2e_pg_ms / 1000.0)},
            'gpu_launches': args.steps * (GL_ITERS + 2),
            'roofline': {'bound': 'hbm', 'achieved': gl_achieved, 'peak': peak_gbs, 'unit': 'GB/s',
                         'frac': gl_achieved / peak_gbs,
                         'traffic': measured_traffic('gl_step_kernel<float, StaticGeom<1102, 275, 2048>, 8, 0, 0>'),
                         'fp32_pipe_active_pct_ncu': measured_ncu('gl_step_kernel<float, StaticGeom<1102, 275, 2048>, 8, 0, 0>',
                                                                  'fp32_pipe_active_pct'),
                         'issue_slots_active_pct_ncu': measured_ncu('gl_step_kernel<float, StaticGeom<1102, 275, 2048>, 8, 0, 0>',
                                                                    'issue_slots_active_pct'),
                         'kernel': 'gl_step_kernel',
                         'peak_source': peak_src, 'ms_per_launch': iter_ms,
                         'algorithmic_bytes_per_launch': gl_alg_bytes,
                         'how': '(CUDA-event time of the 50-iteration call - same call with 0 iterations) / 50'},
            'cpu_baseline': cpu,
            'clocks': clocks,
            'features': {'metric': 'feature_audio_sec_per_sec',
                         'workload': 'BASELINE configs[1]: STFT -> linear + 80-mel dB-normalised features, '
                                     '256 ragged clips per GPU',
                         'f64': feat['f64'], 'f32_fast': feat['f32'], 'e2e': feat_e2e,
                         'dtype': 'f64 (default; f32_fast is the opt-in float32 transform)'},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------------
# BASELINE configs[3] / configs[4]: sharded multi-GPU workloads (torchrun, one rank per GPU)
# ------------------------------------------------------------------------------------------------
def _dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    return torch, dist, rank, world, dev


def run_corpus(args):
    """configs[3]: corpus pass over 13,100 synthetic LJSpeech-length clips sharded by clip:
    (1) per-clip dB statistics (n_fft 1024 / hop 256) + one all-reduce -> corpus constants,
    (2) feature pre-calculation (2048 / 275 / 1102, reduction 5) with those constants.
    Host numpy in, host numpy out (pinned staging); strong scaling (fixed total work)."""
    torch, dist, rank, world, dev = _dist_setup()
    from single_speaker_tts_b200 import distributed
    from single_speaker_tts_b200.audio.features import features_batch
    from single_speaker_tts_b200.synthetic import ClipPlan
    n_total = args.clips or 13100
    plan = ClipPlan(n_total, seed=3, kind='ljspeech', pool=32)
    shards = distributed.shard_by_cost(plan.frames(HOP), world)
    mine = shards[rank]
    wavs = plan.clips(mine)
    audio_total = float(plan.lengths.sum()) / SR

    def one_pass():
        mean4, mn4, mx4, _ = distributed.corpus_decibel_statistics(wavs, mine, n_total, SR, batch_clips=512)
        lin_max, lin_ref, mel_max, mel_ref = mean4          # tacotron/dataset_statistics.py:35-39
        n_rows = 0
        for s in range(0, len(wavs), 256):
            feats = features_batch(wavs[s:s + 256], NFFT, HOP, WIN, SR, 80, 0, 8000,
                                   lin_ref, lin_max, mel_ref, mel_max, reduction=5)
            n_rows += sum(m.shape[0] for m, _ in feats)
        return mean4, n_rows

    for _ in range(max(1, args.warmup)):
        one_pass()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        mean4, n_rows = one_pass()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            'metric': 'corpus_pass_audio_sec_per_sec', 'value': audio_total * args.steps / (ms / 1000.0),
            'unit': 'audio-s/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'BASELINE configs[3]: statistics (1024/256) + all-reduce + feature precalc '
                                   '(2048/275/1102, r=5) over %d LJSpeech-length clips, end to end' % n_total,
                       'sharding': 'by clip, balanced by frames; one all-reduce of the (n_clips, 4) table'},
            'corpus_statistics': {'linear_mag_max_db': mean4[0], 'linear_ref_db': mean4[1],
                                  'mel_mag_max_db': mean4[2], 'mel_mag_ref_db': mean4[3]},
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_gl_sharded(args):
    """configs[4]: Griffin-Lim, 100 iterations, 4,096 ragged utterances sharded by utterance
    (no collective), end to end through spectrograms_to_wavs; strong scaling."""
    torch, dist, rank, world, dev = _dist_setup()
    from single_speaker_tts_b200 import _runtime, distributed
    from single_speaker_tts_b200.audio import synthesis
    from single_speaker_tts_b200.synthetic import ClipPlan
    n_total = args.clips or 4096
    n_iter = 100
    plan = ClipPlan(n_total, seed=4, kind='uniform', pool=32)
    frames = plan.frames(HOP)
    mine = distributed.shard_by_cost(frames, world)[rank]
    audio_total = float((HOP * (frames - 1)).sum()) / SR
    mags = []
    for s in range(0, len(mine), 256):       # |STFT| of this rank's clips via the feature kernel
        idx = mine[s:s + 256]
        fb = _runtime.stft_features_batch(plan.clips(idx), NFFT, HOP, WIN, want_spec=True, precision='f64')
        for i in range(len(idx)):
            mags.append(np.abs(fb.rows(fb.spec, i)).T)

    def one_pass():
        for s in range(0, len(mags), 512):
            synthesis.spectrograms_to_wavs(mags[s:s + 512], WIN, HOP, NFFT, n_iter, seed=1234 + s)

    for _ in range(max(1, args.warmup)):
        one_pass()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            'metric': 'griffin_lim_100it_audio_sec_per_sec', 'value': audio_total * args.steps / (ms / 1000.0),
            'unit': 'audio-s/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'BASELINE configs[4]: Griffin-Lim 100 it over %d ragged utterances, '
                                   'end to end (host numpy in / out)' % n_total,
                       'sharding': 'by utterance, balanced by frames; no collective'},
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='gl256', choices=['gl256', 'corpus', 'gl4096'],
                    help='gl256 = configs[1]+[2] (default, the driver contract); corpus = configs[3]; gl4096 = configs[4]')
    ap.add_argument('--clips', type=int, default=0, help='override the clip count of corpus / gl4096')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'corpus':
        run_corpus(args)
    elif args.workload == 'gl4096':
        run_gl_sharded(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
