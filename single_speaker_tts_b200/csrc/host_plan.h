// Host-side planning shared by the CUDA library (sstts.cu) and the CPU emulator used in tests:
// tile tables for ragged batches, offset tables, transform tables and the mel filterbank.
// Plain C++ (no CUDA), so it compiles with g++ as well as nvcc.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "stft_kernels.cuh"

namespace sstts {

// Frames per tile = warps per CTA: every tile is processed in one round, one frame per warp.
#ifndef SSTTS_WARPS
#define SSTTS_WARPS 8
#endif
constexpr int kTileFrames = SSTTS_WARPS;
constexpr int kWarps = SSTTS_WARPS;
// Griffin-Lim float32 kernels: warps per CTA = frames per tile.  Measured (tools/ab_bench.sh): 8 warps,
// 2 CTAs / SM, 128 registers: 0.605 ms per iteration launch; 9 warps (18 resident warps, but the
// register file is split four ways, so 96 registers / thread and spills): 0.664 ms.
// float64 feature kernel: warps per CTA (a tile still holds kTileFrames frames: with fewer warps
// every warp transforms several frames of the tile).  ~250 registers / thread allow 8 warps per SM
// either as one CTA of 8 (0.74 ms per 256-clip batch) or as two independent CTAs of 4 whose phases
// interleave (0.63 ms, tools/ab_bench.sh).
#ifndef SSTTS_FEAT_WARPS_F64
#define SSTTS_FEAT_WARPS_F64 4
#endif
constexpr int kFeatWarpsF64 = SSTTS_FEAT_WARPS_F64;
constexpr int kFeatWarpsF64Native = 4;   // native n_fft 1024 path in float64: two frames per warp, 8-frame tiles
#ifndef SSTTS_GL_WARPS
#define SSTTS_GL_WARPS 8
#endif
constexpr int kGlWarps = SSTTS_GL_WARPS;

// Smallest tile of a multi-tile utterance: (ft + 1) * hop >= win keeps same-parity spans and a
// tile's two edge regions disjoint (>= 4 for win 1102 / hop 275).
inline int min_tile_frames(int win, int hop) {
  int v = (win + hop - 1) / hop - 1;
  return v < 2 ? 2 : v;
}

// Split n_frames into consecutive tiles of at most `tile` frames (one frame per warp); a
// remainder shorter than `min_tile` borrows frames from its predecessor so every tile of a
// multi-tile utterance has >= min_tile frames (needs tile >= 2 * min_tile).
inline void split_frames(int n_frames, int tile, int min_tile, std::vector<std::pair<int, int> >& out) {
  out.clear();
  if (n_frames <= 0) return;
  int a = 0;
  while (a < n_frames) {
    int b = a + tile;
    if (b >= n_frames) b = n_frames;
    else if (n_frames - b < min_tile) b = n_frames - min_tile;   // (tile - (min_tile - r), min_tile)
    out.push_back(std::make_pair(a, b));
    a = b;
  }
}

struct GLPlanHost {
  int n_utts = 0, win = 0, hop = 0, n_fft = NFFT, span_max = 0, max_tile = 0;
  std::vector<long long> frame_off, pad_off, sample_off;
  std::vector<GLTile> tiles;
  long long total_frames = 0, total_pad = 0, total_samples = 0;
};

// Griffin-Lim at n_fft = 1024 runs natively (512-point complex transform per half-warp, two frames per warp,
// tiles of 2 x warps frames) whenever the kernel's shared memory fits the device; SSTTS_GL_NATIVE1024=0 keeps
// the older embedding in the 2048-point transform (A/B switch, also used by the tests to cover that path).
inline bool gl_native_1024_enabled() {
  const char* e = getenv("SSTTS_GL_NATIVE1024");
  return !(e && e[0] == '0');
}
template <typename T>
inline bool gl_native_1024(int n_fft, int win, int hop, int warps, size_t smem_limit) {
  if (n_fft != 1024 || !gl_native_1024_enabled()) return false;
  if (2 * min_tile_frames(win, hop) > 2 * warps) return false;
  const int span_max = (2 * warps - 1) * hop + win;
  return gl_step_smem_bytes<T>(warps, win, hop, span_max, false, true) <= smem_limit;
}

inline bool build_gl_plan(int n_utts, const long long* frame_off, int win, int hop, GLPlanHost& P,
                          std::string& err, int n_fft = NFFT, int tile_frames = kTileFrames) {
  if (n_fft != 2048 && n_fft != 1024 && n_fft != 512) { err = "n_fft must be 2048, 1024 or 512"; return false; }
  if (win < 2 || win > n_fft || hop < 1 || hop > win) { err = "need 1 <= hop <= win <= n_fft"; return false; }
  if ((n_fft - win) % 2 != 0) { err = "n_fft - win_length must be even"; return false; }
  P.n_fft = n_fft;
  const int min_tile = min_tile_frames(win, hop);
  if (2 * min_tile > tile_frames) { err = "win_length / hop_length > 5 is not supported"; return false; }
  P.n_utts = n_utts; P.win = win; P.hop = hop;
  P.frame_off.assign(frame_off, frame_off + n_utts + 1);
  P.pad_off.assign(n_utts + 1, 0);
  P.sample_off.assign(n_utts + 1, 0);
  P.tiles.clear();
  P.span_max = win;
  P.max_tile = 1;
  std::vector<std::pair<int, int> > parts;
  for (int u = 0; u < n_utts; ++u) {
    const long long T = frame_off[u + 1] - frame_off[u];
    if (T < 1 || T > (1 << 22)) { err = "every utterance needs between 1 and 2^22 frames"; return false; }
    const long long padded = n_fft + (long long)hop * (T - 1);
    P.pad_off[u + 1] = P.pad_off[u] + ((padded + 3) & ~3LL);
    P.sample_off[u + 1] = P.sample_off[u] + (long long)hop * (T - 1);
    if (T < 2) continue;  // hop * (T - 1) == 0 output samples: nothing to compute
    split_frames((int)T, tile_frames, min_tile, parts);
    for (size_t i = 0; i < parts.size(); ++i) {
      GLTile t; t.utt = u; t.a = parts[i].first; t.b = parts[i].second; t.parity = (int)(i & 1);
      t.f0 = frame_off[u]; t.poff = P.pad_off[u]; t.n_frames = (int)T; t.reserved = 0; t.soff = P.sample_off[u];
      P.tiles.push_back(t);
      const int ft = t.b - t.a;
      if (ft > P.max_tile) P.max_tile = ft;
      const int span = (ft - 1) * hop + win;
      if (span > P.span_max) P.span_max = span;
    }
  }
  P.total_frames = frame_off[n_utts] - frame_off[0];
  P.total_pad = P.pad_off[n_utts];
  P.total_samples = P.sample_off[n_utts];
  return true;
}

struct FeatPlanHost {
  int n_clips = 0, win = 0, hop = 0, span_max = 0, reduction = 1;
  std::vector<long long> sample_off, sample_len, frame_off, row_off;   // sample_off: clip starts
  std::vector<FeatTile> tiles;
  long long total_frames = 0, total_rows = 0;
};

// clip_start[c] / clip_len[c]: first sample and length of clip c inside the packed wav buffer
// (clips need not be contiguous: trimmed clips keep their place in the untrimmed upload).
// The feature kernel transforms n_fft = 1024 natively (two frames per warp, 16-frame tiles); 2048 and 512 go
// through the 2048-point transform (512 embedded in it).
inline bool feat_native_1024(int n_fft) { return n_fft == 1024; }

// Frames per tile: one frame per warp and round (kTileFrames), or -- native n_fft 1024 path -- two frames per
// warp of the CTA (float32: 8 warps, float64: kFeatWarpsF64 warps; a larger tile would push the float64
// kernel's sample buffers past the shared memory that lets two CTAs share an SM).
#ifndef SSTTS_FEAT_TILE_F64
#define SSTTS_FEAT_TILE_F64 SSTTS_WARPS
#endif
inline int feat_tile_frames(int n_fft, bool f64) {
  if (!feat_native_1024(n_fft)) return f64 ? SSTTS_FEAT_TILE_F64 : kTileFrames;
  const int t = 2 * (f64 ? kFeatWarpsF64Native : kWarps);
  return t < kNativeTileFrames ? t : kNativeTileFrames;
}

inline bool build_feat_plan(int n_clips, const long long* clip_start, const long long* clip_len, int n_fft,
                            int win, int hop, int reduction, FeatPlanHost& P, std::string& err, bool f64 = false) {
  const int tile_frames = feat_tile_frames(n_fft, f64);
  if (n_fft != 2048 && n_fft != 1024 && n_fft != 512) { err = "n_fft must be 2048, 1024 or 512"; return false; }
  if (win < 2 || win > n_fft || hop < 1) { err = "need hop >= 1 and 2 <= win <= n_fft"; return false; }
  if ((n_fft - win) % 2 != 0) { err = "n_fft - win_length must be even"; return false; }
  if (reduction < 1) reduction = 1;
  P.n_clips = n_clips; P.win = win; P.hop = hop; P.reduction = reduction;
  P.sample_off.assign(clip_start, clip_start + n_clips);
  P.sample_len.assign(clip_len, clip_len + n_clips);
  P.frame_off.assign(n_clips + 1, 0);
  P.row_off.assign(n_clips + 1, 0);
  P.tiles.clear();
  P.span_max = win;
  for (int c = 0; c < n_clips; ++c) {
    const long long N = clip_len[c];
    if (clip_start[c] < 0 || N < 1 || N > (1LL << 30)) { err = "every clip needs between 1 and 2^30 samples"; return false; }
    const long long T = 1 + N / hop;
    const long long rows = ((T + reduction - 1) / reduction) * reduction;
    P.frame_off[c + 1] = P.frame_off[c] + T;
    P.row_off[c + 1] = P.row_off[c] + rows;
    for (long long a = 0; a < T; a += tile_frames) {
      FeatTile t; t.clip = c; t.a = (int)a; t.b = (int)((a + tile_frames < T) ? a + tile_frames : T);
      t.last = (t.b == T) ? 1 : 0;
      t.soff = clip_start[c]; t.r0 = P.row_off[c];
      t.n_samples = (int)N; t.n_frames = (int)T; t.n_rows = (int)rows; t.reserved = 0;
      P.tiles.push_back(t);
      const int span = (t.b - t.a - 1) * hop + win;
      if (span > P.span_max) P.span_max = span;
    }
  }
  P.total_frames = P.frame_off[n_clips];
  P.total_rows = P.row_off[n_clips];
  return true;
}

// Transform tables in double; callers narrow to the kernel's arithmetic type.
inline void make_tables(int win, std::vector<double>& tw1024, std::vector<double>& w2048,
                        std::vector<double>& window) {
  const double pi = 3.14159265358979323846264338327950288;
  tw1024.resize(2 * 1024);
  for (int a = 0; a < 32; ++a)
    for (int b = 0; b < 32; ++b) {
      const int e = (a * b) % 1024;
      tw1024[2 * (a * 32 + b)] = std::cos(2.0 * pi * e / 1024.0);
      tw1024[2 * (a * 32 + b) + 1] = -std::sin(2.0 * pi * e / 1024.0);
    }
  w2048.resize(2 * 1024);
  for (int k = 0; k < 1024; ++k) {
    w2048[2 * k] = std::cos(2.0 * pi * k / 2048.0);
    w2048[2 * k + 1] = -std::sin(2.0 * pi * k / 2048.0);
  }
  window.resize(win);
  // scipy.signal.get_window('hann', win, fftbins=True): 0.5 - 0.5 cos(2 pi n / win)
  for (int n = 0; n < win; ++n) window[n] = 0.5 - 0.5 * std::cos(2.0 * pi * n / win);
}

// Tables of the native n_fft = 1024 feature path (halfwarp_fft512), in the slots the kernel loads:
// tw[k1 * 16 + hl] = exp(-2 pi i k1 hl / 512) (k1 < 32, hl < 16; padded to the 1024 entries the prologue
// copies), w[k] = exp(-2 pi i k / 1024) (k < 512), and the periodic Hann window.
inline void make_tables_native1024(int win, std::vector<double>& tw512, std::vector<double>& w1024,
                                   std::vector<double>& window) {
  const double pi = 3.14159265358979323846264338327950288;
  tw512.assign(2 * 1024, 0.0);
  for (int k1 = 0; k1 < 32; ++k1)
    for (int hl = 0; hl < 16; ++hl) {
      const int e = (k1 * hl) % 512;
      tw512[2 * (k1 * 16 + hl)] = std::cos(2.0 * pi * e / 512.0);
      tw512[2 * (k1 * 16 + hl) + 1] = -std::sin(2.0 * pi * e / 512.0);
    }
  w1024.resize(2 * 1024);
  for (int k = 0; k < 1024; ++k) {
    w1024[2 * k] = std::cos(2.0 * pi * k / 1024.0);
    w1024[2 * k + 1] = -std::sin(2.0 * pi * k / 1024.0);
  }
  window.resize(win);
  for (int n = 0; n < win; ++n) window[n] = 0.5 - 0.5 * std::cos(2.0 * pi * n / win);
}

// librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=True, norm=1) (0.6.x), float64, as CSR
// over mel bins (each filter's support is one contiguous run of FFT bins).
// Used by the reference at audio/features.py:75-80.
struct MelCSR {
  std::vector<int> ptr, k0;
  std::vector<double> w;
};
inline void make_mel_csr(int sr, int n_fft, int n_mels, double fmin, double fmax, MelCSR& M,
                         std::vector<double>* dense = nullptr) {
  const int n_bins = 1 + n_fft / 2;
  std::vector<double> fftfreqs(n_bins), mel_f(n_mels + 2);
  // np.linspace(0, sr / 2, n_bins): start + i * step
  const double step = (double(sr) / 2.0) / (n_bins - 1);
  for (int i = 0; i < n_bins; ++i) fftfreqs[i] = i * step;
  fftfreqs[n_bins - 1] = double(sr) / 2.0;
  const double mel_lo = 2595.0 * std::log10(1.0 + fmin / 700.0);
  const double mel_hi = 2595.0 * std::log10(1.0 + fmax / 700.0);
  const double mstep = (mel_hi - mel_lo) / (n_mels + 1);
  for (int i = 0; i < n_mels + 2; ++i) {
    double m = mel_lo + i * mstep;
    if (i == n_mels + 1) m = mel_hi;
    mel_f[i] = 700.0 * (std::pow(10.0, m / 2595.0) - 1.0);
  }
  if (dense) dense->assign((size_t)n_mels * n_bins, 0.0);
  M.ptr.assign(1, 0); M.k0.clear(); M.w.clear();
  for (int i = 0; i < n_mels; ++i) {
    const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    int first = -1, last = -1;
    std::vector<double> row(n_bins);
    for (int k = 0; k < n_bins; ++k) {
      const double lower = -(mel_f[i] - fftfreqs[k]) / fd0;
      const double upper = (mel_f[i + 2] - fftfreqs[k]) / fd1;
      double v = lower < upper ? lower : upper;
      if (v < 0.0) v = 0.0;
      v *= enorm;
      row[k] = v;
      if (v != 0.0) { if (first < 0) first = k; last = k; }
    }
    if (first < 0) { first = 0; last = -1; }
    M.k0.push_back(first);
    for (int k = first; k <= last; ++k) M.w.push_back(row[k]);
    M.ptr.push_back((int)M.w.size());
    if (dense) for (int k = 0; k < n_bins; ++k) (*dense)[(size_t)i * n_bins + k] = row[k];
  }
}

// The filterbank padded for the kernel's dB-feature mode: slot j holds the filters
// m = mbase[j] + lane (top 32 filters first -- filter lengths grow with m, so a slot's filters have
// similar lengths).  A filter is stored as PAIRS of weights starting at the even bin k0 & ~1:
// pair i of filter m is (w[2 * (woff[j] + 32 * i + lane)], w[... + 1]) for bins (k0 & ~1) + 2 i and
// + 2 i + 1, zero outside the filter's support; len / woff / total count pairs.
// ok == false when the layout does not fit the kernel's limits (then the generic mode is used).
struct MelPadded {
  int n_slots = 0, total = 0;
  int len[4] = {0, 0, 0, 0}, woff[4] = {0, 0, 0, 0}, mbase[4] = {0, 0, 0, 0};
  std::vector<float> w;
  bool ok = false;
};
inline void make_mel_padded(const MelCSR& M, int n_mels, int mag_elems, MelPadded& P) {
  P = MelPadded();
  if (n_mels < 1 || n_mels > 128) return;
  P.n_slots = (n_mels + 31) / 32;
  for (int j = 0; j < P.n_slots; ++j) {
    P.mbase[j] = n_mels - 32 * (j + 1);
    int L = 0;
    for (int l = 0; l < 32; ++l) {
      const int m = P.mbase[j] + l;
      if (m < 0 || m >= n_mels) continue;
      const int n = M.ptr[m + 1] - M.ptr[m] + (M.k0[m] & 1);    // leading zero when k0 is odd
      if ((n + 1) / 2 > L) L = (n + 1) / 2;
    }
    P.len[j] = L;
    P.woff[j] = P.total;
    P.total += 32 * L;
  }
  P.w.assign((size_t)2 * P.total, 0.0f);
  for (int j = 0; j < P.n_slots; ++j)
    for (int l = 0; l < 32; ++l) {
      const int m = P.mbase[j] + l;
      if (m < 0 || m >= n_mels) continue;
      const int k0e = M.k0[m] & ~1;
      if (k0e + 2 * P.len[j] > mag_elems) return;    // padded reads would leave the |S| plane
      for (int i = M.ptr[m]; i < M.ptr[m + 1]; ++i) {
        const int e = (i - M.ptr[m]) + (M.k0[m] & 1);     // element index relative to the even start
        P.w[2 * ((size_t)P.woff[j] + 32 * (e >> 1) + l) + (e & 1)] = (float)M.w[i];
      }
    }
  P.ok = P.total > 0;
}

}  // namespace sstts
