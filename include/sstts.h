/* sstts.h -- C ABI of libsstts.so: the B200 (sm_100a) audio hot path of single-speaker-tts.
 *
 * The reference (yweweler/single-speaker-tts) is pure Python; its "operator API" for this path
 * is the set of numpy-in / numpy-out functions of `audio/` plus three `datasets/` methods
 * (SURVEY.md section 8b).  This header is what a binding for those functions would bind.  Every
 * entry point names the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C types only; `*_dev` pointers are device pointers owned by the caller, `*_host`
 *     pointers are host pointers read during the call; `stream` is a cudaStream_t passed as void*
 *     (NULL = the legacy default stream);
 *   - ragged batches are packed back to back and described by int64 prefix-sum offset tables of
 *     length n + 1;
 *   - every function returns 0 on success and a negative sstts_status on failure; the message
 *     of the calling thread's last failure is available from sstts_last_error();
 *   - plans are immutable after creation and may be shared between threads; calls on different
 *     streams may run concurrently as long as their workspaces differ.  No C++ exception
 *     crosses this boundary.
 */
#ifndef SSTTS_H_
#define SSTTS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSTTS_VERSION 100 /* 0.1.0 */

typedef enum sstts_status {
  SSTTS_OK = 0,
  SSTTS_ERR_INVALID = -1,   /* bad argument / unsupported geometry */
  SSTTS_ERR_CUDA = -2,      /* a CUDA runtime call failed */
  SSTTS_ERR_NO_DEVICE = -3, /* no usable sm_100 device: there is NO CPU fallback */
  SSTTS_ERR_ASSERT = -4     /* the reference would raise AssertionError (audio/conversion.py:47-49) */
} sstts_status;

typedef enum sstts_precision {
  SSTTS_F32 = 0, /* transform arithmetic in float32 */
  SSTTS_F64 = 1  /* transform arithmetic in float64 (matches librosa's float64 stft bin for bin) */
} sstts_precision;

/* STFT geometry + mel filterbank.  Mirrors the arguments the reference threads through
 * audio/features.py:5-6,116 and audio/synthesis.py:43 (values: tacotron/params/model.py:13-33). */
typedef struct sstts_stft_config {
  int n_fft;         /* 2048, 1024 or 512.  1024 is transformed natively (512-point complex FFT on half a warp, two
                        frames per warp) by feature and Griffin-Lim plans; n_fft 512 runs embedded in the
                        2048-point transform */
  int win_length;    /* <= n_fft, n_fft - win_length even; periodic Hann, zero-padded centred */
  int hop_length;    /* ceil(win_length / hop_length) <= 5 for Griffin-Lim */
  int sampling_rate; /* mel filterbank only */
  int n_mels;        /* 0: no filterbank */
  double mel_fmin;
  double mel_fmax;   /* <= 0: sampling_rate / 2 */
  int precision;     /* sstts_precision */
} sstts_stft_config;

int sstts_version(void);
const char* sstts_last_error(void);

/* Number of usable CUDA devices (>= 1) or SSTTS_ERR_NO_DEVICE. */
int sstts_device_count(void);

/* ------------------------------------------------------------------------------------------
 * Griffin-Lim -- replaces audio/synthesis.py:43-125 `griffin_lim_v2` (and :5-40
 * `spectrogram_to_wav`) for a ragged batch of utterances.
 * ------------------------------------------------------------------------------------------ */
typedef struct sstts_gl_plan sstts_gl_plan;

/* frame_off_host[n_utts + 1]: prefix sums of the per-utterance frame counts T_u (>= 1).
 * Utterance u produces hop_length * (T_u - 1) samples (librosa.istft centre trimming).
 * Plans are created for the current device.  The configuration's tables (twiddles, window, mel
 * filterbank) are cached per process; a plan owns only its offset / tile tables (one stream-ordered
 * pool allocation).  Destroy a plan only after the work enqueued with it has completed. */
int sstts_gl_plan_create(const sstts_stft_config* cfg, int n_utts, const int64_t* frame_off_host,
                         sstts_gl_plan** plan_out);
void sstts_gl_plan_destroy(sstts_gl_plan* plan);
size_t sstts_gl_workspace_bytes(const sstts_gl_plan* plan);
int64_t sstts_gl_total_frames(const sstts_gl_plan* plan);
int64_t sstts_gl_total_samples(const sstts_gl_plan* plan);
/* Output sample offsets [n_utts + 1] (host memory owned by the plan). */
const int64_t* sstts_gl_sample_offsets(const sstts_gl_plan* plan);

/* mag_dev    : (sum T, n_fft/2 + 1) float32, frame-major -- |S| of every utterance;
 * phase0_dev : (sum T, n_fft/2 + 1) interleaved (re, im) float32 -- initial unit phasors, the
 *              reference's `np.exp(2j * np.pi * np.random.rand(...))` (audio/synthesis.py:85);
 * n_iter     : reconstruction iterations (tacotron/params/model.py:48 uses 50);
 * wav_out_dev: (total_samples) float32;
 * mse_frame_dev: optional (sum T) float64 -- per frame sum over bins of (|S| - |stft|)^2 of the
 *              LAST iteration; the reference's `mse` (audio/synthesis.py:112) is their sum over an
 *              utterance divided by (n_fft/2 + 1) * T.  Requires n_iter >= 1.
 * All work is enqueued on `stream`; nothing synchronises. */
int sstts_griffin_lim(const sstts_gl_plan* plan, const float* mag_dev, const float* phase0_dev,
                      int n_iter, void* workspace_dev, float* wav_out_dev, double* mse_frame_dev,
                      void* stream);

/* Same, with the initial phase drawn inside the first launch instead of read from memory: element
 * i = frame_row * (n_fft/2 + 1) + bin gets the phasor sstts_random_phase_at(seed, first_element, ...)
 * would have written at i (bit-identical results, no (sum T, bins, 2) phase buffer). */
int sstts_griffin_lim_seeded(const sstts_gl_plan* plan, const float* mag_dev, uint64_t seed,
                             int64_t first_element, int n_iter, void* workspace_dev, float* wav_out_dev,
                             double* mse_frame_dev, void* stream);

/* Peak normalisation of every utterance of a Griffin-Lim result in place -- replaces the
 * `librosa.util.normalize(wav, norm=np.inf)` of `save_wav(path, wav, sr, norm=True)`
 * (audio/io.py:33-53; tacotron/inference.py:199): y / max|y| (true division), unchanged when the peak
 * is below float32 tiny.  wav_dev is the wav_out_dev of sstts_griffin_lim for the same plan. */
int sstts_peak_normalize(const sstts_gl_plan* plan, float* wav_dev, void* stream);

/* Unit phasors from host-drawn uniforms -- the device half of audio/synthesis.py:85
 * `np.exp(2j * np.pi * np.random.rand(bins, T))`: the per-item functions keep the reference's use of
 * numpy's global random stream, so the host draws the uniforms and uploads them as they are.
 * uniform_dev: (n_bins, n_frames) float64, bin-major like np.random.rand(bins, T); phase_dev:
 * (n_frames, n_bins) interleaved (re, im) float32 = exp(2 pi i u) in the layout sstts_griffin_lim reads. */
int sstts_phase_from_uniform(const double* uniform_dev, int64_t n_frames, int n_bins, float* phase_dev,
                             void* stream);

/* Fill n unit phasors exp(2 pi i u), u ~ U[0, 1) from a counter-based generator keyed by seed
 * (batched extension: replaces the host-side np.random.rand of audio/synthesis.py:85). */
int sstts_random_phase(uint64_t seed, int64_t n, float* phase_dev, void* stream);
/* Same stream of phasors, starting at element `first` (so a batch can be filled in pieces). */
int sstts_random_phase_at(uint64_t seed, int64_t first, int64_t n, float* phase_dev, void* stream);

/* Call-site glue in front of Griffin-Lim -- replaces tacotron/inference.py:94-101,175 (and
 * tacotron/serve.py:42-59): audio/conversion.py:81-102 `inv_normalize_decibel`, :32-53
 * `decibel_to_magnitude` and `np.power(mag, magnitude_power)` fused into one pass over the model
 * output.  norm_dev / mag_out_dev: n float32 values (frame-major (sum T, bins), same layout in and
 * out; may alias).  *flag_dev (optional int, caller-zeroed) is set to 1 if any dB value is below
 * -100, where the reference raises AssertionError (audio/conversion.py:47-49). */
int sstts_denormalize_magnitude(const float* norm_dev, int64_t n, double ref_db, double max_db,
                                double power, float* mag_out_dev, int* flag_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * STFT features -- replaces audio/features.py:116-145 `linear_scale_spectrogram`, :5-86
 * `mel_scale_spectrogram`, audio/conversion.py:5-29 `magnitude_to_decibel`, :56-78
 * `normalize_decibel`, datasets/statistics.py:11-66 `decibel_statistics` and the core of
 * datasets/lj_speech.py:106-156 `load_audio` for a ragged batch of clips.
 * ------------------------------------------------------------------------------------------ */
typedef struct sstts_feat_plan sstts_feat_plan;

/* sample_off_host[n_clips + 1]: prefix sums of the clip lengths N_c (>= 1).  Clip c yields
 * T_c = 1 + N_c / hop_length frames; with reduction r > 1 its output rows are padded with zero
 * rows to a multiple of r (datasets/dataset_helper.py:357-401). */
int sstts_feat_plan_create(const sstts_stft_config* cfg, int n_clips, const int64_t* sample_off_host,
                           int reduction, sstts_feat_plan** plan_out);
/* Same, for clips that are not packed back to back: clip c occupies
 * wav_dev[clip_start[c] .. clip_start[c] + clip_len[c]) (e.g. the trimmed part of an untrimmed upload). */
int sstts_feat_plan_create_ranges(const sstts_stft_config* cfg, int n_clips, const int64_t* clip_start_host,
                                  const int64_t* clip_len_host, int reduction, sstts_feat_plan** plan_out);
void sstts_feat_plan_destroy(sstts_feat_plan* plan);
int64_t sstts_feat_total_frames(const sstts_feat_plan* plan);
int64_t sstts_feat_total_rows(const sstts_feat_plan* plan);
const int64_t* sstts_feat_frame_offsets(const sstts_feat_plan* plan); /* [n_clips + 1] */
const int64_t* sstts_feat_row_offsets(const sstts_feat_plan* plan);   /* [n_clips + 1] */
/* Dense (n_mels, n_fft/2 + 1) float64 filterbank the plan uses (host memory owned by the plan);
 * equals librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=True) as called at
 * audio/features.py:75-80. */
const double* sstts_feat_mel_basis(const sstts_feat_plan* plan);

typedef struct sstts_feat_outputs {
  float* spec_dev;     /* (rows, bins) interleaved complex64 STFT, or NULL   [features.py:145] */
  float* lin_db_dev;   /* (rows, bins) 20 log10(max(1e-5, |S|)), normalised if `normalize`     */
  float* mel_db_dev;   /* (rows, n_mels) same for mel_basis @ |S| ** mel_power                 */
  double* mel_raw_dev; /* (rows, n_mels) mel_basis @ |S| ** mel_power       [features.py:84]   */
  double* minmax_dev;  /* (n_clips, 4) min lin dB, max lin dB, min mel dB, max mel dB
                          [datasets/statistics.py:63-66]                                        */
  int normalize;       /* apply audio/conversion.py:78 with the constants below                */
  double lin_ref_db, lin_max_db, mel_ref_db, mel_max_db;
  double mel_power;    /* audio/features.py:71 `power` */
  int force_generic;   /* 0: the library picks the fused dB-feature kernel mode when the request is
                          lin_db + mel_db only (n_fft 2048, power 1); 1: always the generic mode
                          (validation / A-B runs; results agree to float32 rounding)              */
} sstts_feat_outputs;

int sstts_stft_features(const sstts_feat_plan* plan, const float* wav_dev,
                        const sstts_feat_outputs* out, void* stream);

/* 16-bit PCM decode -- the sample conversion of `load_wav` (audio/io.py:5-30, int16 / 32768 -> float32)
 * on the device, so clips can be uploaded as 2-byte samples. */
int sstts_pcm16_to_float(const int16_t* pcm_dev, int64_t n, float* out_dev, void* stream);

/* MFCCs of a mel spectrogram -- replaces `librosa.feature.mfcc(S=mel_spec, n_mfcc=...)` at
 * audio/features.py:111: mel_dev (n_frames, n_mels) float64 frame-major -> out_dev (n_frames, n_mfcc)
 * float64, the orthonormal DCT-II of every frame truncated to n_mfcc coefficients. */
int sstts_dct_project(const double* mel_dev, int64_t n_frames, int n_mels, int n_mfcc, double* out_dev,
                      void* stream);

/* Time-stretch glue -- the part of `librosa.core.phase_vocoder(stft, rate)` that audio/effects.py:77-80
 * keeps (its magnitude): spec_dev is a (n_frames, n_bins) interleaved complex64 STFT, frame-major;
 * mag_out_dev receives (sstts_stretch_frames(n_frames, rate), n_bins) float32 magnitudes, linearly
 * interpolated at the fractional frame positions t * rate. */
int64_t sstts_stretch_frames(int64_t n_frames, double rate);
int sstts_stretch_magnitude(const float* spec_dev, int64_t n_frames, int n_bins, double rate,
                            float* mag_out_dev, void* stream);

/* Silence trimming -- replaces `librosa.effects.trim(wav)` as called by datasets/lj_speech.py:119
 * (top_db 60, frame_length 2048, hop_length 512; wrapper audio/effects.py:188-215).
 * clip_start_dev / clip_len_dev: int64[n_clips] on the device; bounds_dev: int64[n_clips * 2]
 * receives (start, end) relative to each clip, (0, 0) for an all-silent clip. */
int sstts_trim_bounds(const float* wav_dev, int n_clips, const int64_t* clip_start_dev,
                      const int64_t* clip_len_dev, double top_db, int frame_length, int hop_length,
                      int64_t* bounds_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSTTS_H_ */
