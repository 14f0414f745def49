// sm_100a kernels for the audio hot path: Griffin-Lim iterations and the STFT feature pipeline.
//
// Common structure (both halves):
//   * one WARP per frame: the n_fft = 2048 real transform is a 1024-point complex FFT held in
//     the warp's registers (fft_warp.cuh), packed z[n] = x[2n] + i x[2n+1] and (un)tangled with
//     the conjugate-pair identities below; the pair partner lives in lane (32 - l) & 31 and is
//     exchanged with warp shuffles; n_fft = 1024 is a 512-point complex FFT on HALF a warp, two frames per
//     warp (halfwarp_fft512 / halfwarp_ifft512; Griffin-Lim and features), n_fft = 512 is embedded in the
//     2048-point transform;
//   * one CTA per TILE of consecutive frames of one utterance; the tile's sample span is staged
//     in shared memory once (reflect padding resolved while staging), so every waveform sample
//     crosses HBM/L2 once per tile although four to five frames overlap it;
//   * persistent CTAs loop over a host-built tile table; ragged batches are packed with int64
//     offset tables (frames, padded samples, output rows).
//
// Griffin-Lim (reference: audio/synthesis.py:91-123) is re-associated so that the state carried
// between iterations is the overlap-added WAVEFORM instead of the complex spectrum:
//     reference : angles -> istft -> stft -> angles          (x n_iter) -> istft
//     here      : angles0 -> [synth] -> wave_1 ; wave_i -> [stft | E/|E| | x|S| | istft] ->
//                 wave_(i+1)  (x n_iter) ; wave_(n_iter+1) is the reference's final istft.
// The spectrum of a frame never leaves the warp's registers, so per frame and iteration only
// |S| (4.1 KB) is read and ~1.2 KB of waveform is read and written.
//
// Overlap-add without atomics: inside a tile the warps drop their windowed frames into
// per-warp shared-memory slots and all threads then GATHER, per output sample, the frames that
// cover it in ascending frame order.  Across tiles the partially summed edge samples are
// exchanged through two global buffers indexed by tile parity: a tile writes the raw sums of
// its whole span into the buffer of its own parity, and a consumer adds the neighbour tile's
// contribution from the other-parity buffer only inside the two edge regions where the
// neighbour's frames reach in.  Tiles hold at least ceil(win/hop) - 1 (>= 4) frames, which makes
// same-parity spans disjoint and the left/right edge regions of a tile disjoint.
#pragma once
#include "fft_warp.cuh"

namespace sstts {

constexpr int NFFT = 2048;
constexpr int HALF = NFFT / 2;     // centre padding, and #complex points of the packed FFT
constexpr int NBINS = HALF + 1;    // 1025

// ---------------------------------------------------------------------------------------------
// geometry policies: compile-time (win, hop) for the model configuration, run-time otherwise
// ---------------------------------------------------------------------------------------------
// NFFT_EFF < 2048 embeds a shorter transform in the 2048-point one: the frame occupies the first
// NFFT_EFF samples of the zero-padded buffer and only every (2048 / NFFT_EFF)-th bin is kept
// (zero padding in time interpolates the spectrum, so those bins ARE the NFFT_EFF-point DFT).
template <int WIN, int HOP, int NFFT_EFF = 2048> struct StaticGeom {
  static constexpr bool kNative1024 = false;
  SSTTS_HD StaticGeom(int, int, int) {}
  SSTTS_HD constexpr int win() const { return WIN; }
  SSTTS_HD constexpr int hop() const { return HOP; }
  SSTTS_HD constexpr int nfft() const { return NFFT_EFF; }
  SSTTS_HD constexpr int lpad() const { return (NFFT_EFF - WIN) / 2; }   // window offset in the frame
  SSTTS_HD constexpr int cpad() const { return NFFT_EFF / 2; }           // centre (reflect) padding
  SSTTS_HD constexpr int bin_shift() const { return NFFT_EFF == 2048 ? 0 : NFFT_EFF == 1024 ? 1 : 2; }
  // packed elements z[n] (64 samples per register slot) that overlap the window; the rest are zero
  static constexpr int ZLO = ((NFFT_EFF - WIN) / 2) / 64;
  static constexpr int ZHI = ((NFFT_EFF - WIN) / 2 + WIN) / 64 > 31 ? 31 : ((NFFT_EFF - WIN) / 2 + WIN) / 64;   // frames may sit one sample later (SSTTS_GL_FRAME_SHIFT)
};
// n_fft = 1024 transformed natively by the feature kernel: a 512-point complex FFT on half a warp, two
// frames per warp (halfwarp_fft512) -- the geometry of datasets/statistics.py:31-34 and of the STFT in
// audio/effects.py:71-77 -- instead of the 2048-point transform with every other bin dropped.
template <int WIN, int HOP> struct NativeGeom1024 {
  SSTTS_HD NativeGeom1024(int, int, int) {}
  SSTTS_HD constexpr int win() const { return WIN; }
  SSTTS_HD constexpr int hop() const { return HOP; }
  SSTTS_HD constexpr int nfft() const { return 1024; }
  SSTTS_HD constexpr int lpad() const { return (1024 - WIN) / 2; }
  SSTTS_HD constexpr int cpad() const { return 512; }
  SSTTS_HD constexpr int bin_shift() const { return 1; }
  static constexpr int ZLO = 0;
  static constexpr int ZHI = 31;
  static constexpr bool kNative1024 = true;
};
struct DynGeom1024 {
  int win_, hop_;
  SSTTS_HD DynGeom1024(int w, int h, int) : win_(w), hop_(h) {}
  SSTTS_HD int win() const { return win_; }
  SSTTS_HD int hop() const { return hop_; }
  SSTTS_HD constexpr int nfft() const { return 1024; }
  SSTTS_HD int lpad() const { return (1024 - win_) / 2; }
  SSTTS_HD constexpr int cpad() const { return 512; }
  SSTTS_HD constexpr int bin_shift() const { return 1; }
  static constexpr int ZLO = 0;
  static constexpr int ZHI = 31;
  static constexpr bool kNative1024 = true;
};
struct DynGeom {
  static constexpr bool kNative1024 = false;
  int win_, hop_, nfft_;
  SSTTS_HD DynGeom(int w, int h, int n) : win_(w), hop_(h), nfft_(n) {}
  SSTTS_HD int win() const { return win_; }
  SSTTS_HD int hop() const { return hop_; }
  SSTTS_HD int nfft() const { return nfft_; }
  SSTTS_HD int lpad() const { return (nfft_ - win_) / 2; }
  SSTTS_HD int cpad() const { return nfft_ / 2; }
  SSTTS_HD int bin_shift() const { return nfft_ == 2048 ? 0 : nfft_ == 1024 ? 1 : 2; }
  static constexpr int ZLO = 0;
  static constexpr int ZHI = 31;
};

// numpy.pad(mode='reflect') index map for any q (multi-bounce for short signals).
SSTTS_HD int reflect_index(int q, int L) {
  if (L <= 1) return 0;
  const int period = 2 * (L - 1);
  int m = q % period;
  if (m < 0) m += period;
  return m < L ? m : period - m;
}

SSTTS_HD int round_up4(int v) { return (v + 3) & ~3; }

template <typename T> struct StftTables {
  const typename cx_of<T>::type* tw1024;  // [32*32]  exp(-2 pi i a b / 1024) at [a*32+b]
  const typename cx_of<T>::type* w2048;   // [1024]   exp(-2 pi i k / 2048)
  const T* window;                        // [win]    periodic Hann (float64 -> T)
};

// Sum over the frames t in [0, n_frames) covering padded coordinate p of window(p - t hop)^2,
// ascending t (librosa.filters.window_sumsquare restricted to one sample).
template <typename T>
SSTTS_D T window_sumsq(int p, int n_frames, int hop, int win, int lpad, const T* s_win) {
  const int x = p - lpad;
  if (x < 0) return T(0);
  int t_hi = x / hop;
  if (t_hi > n_frames - 1) t_hi = n_frames - 1;
  const int num = x - win + 1;
  const int t_lo = num <= 0 ? 0 : (num + hop - 1) / hop;
  T acc = T(0);
  for (int t = t_lo; t <= t_hi; ++t) {
    const T w = s_win[x - t * hop];
    acc += w * w;
  }
  return acc;
}

// The window lives in shared memory with a zero guard on both sides and shifted by WIN_SHIFT(lpad)
// elements, so that for every even frame position m the pair (w[m - lpad], w[m + 1 - lpad]) is
// one aligned 8 / 16-byte load and reads zeros just outside the window.
constexpr int WIN_TAB_PAD = 4;
SSTTS_HD int win_shift(int lpad) { return (lpad & 1) ? 1 : 2; }
template <typename T>
SSTTS_D T* load_window_table(T* s_wtab, const T* g_window, int win, int lpad, int tid, int nthreads) {
  T* s_win = s_wtab + win_shift(lpad);
  for (int i = tid; i < round_up4(win + WIN_TAB_PAD); i += nthreads) {
    const int j = i - win_shift(lpad);
    s_wtab[i] = (j >= 0 && j < win) ? g_window[j] : T(0);
  }
  return s_win;
}

// np.finfo(np.float32).tiny -- librosa's guard for the window-sum division.
#define SSTTS_F32_TINY 1.17549435e-38f

// Reciprocal window sums for samples all of whose covering frames exist ("interior"): they only
// depend on (p - lpad) mod hop.  Same ascending-frame summation order as window_sumsq().
template <typename T>
SSTTS_D void fill_interior_rwss(T* s_rw, const T* s_win, int hop, int win, int tid, int nthreads, T inv_nfft) {
  for (int r = tid; r < hop; r += nthreads) {
    T acc = T(0);
    for (int j = (win - 1 - r) / hop; j >= 0; --j) {
      const T w = s_win[r + j * hop];
      acc += w * w;
    }
    // the partial-sum buffers hold n_fft x the true overlap-add sums: the 1/n_fft of the inverse
    // transform is folded in here (an exact power-of-two scaling)
    s_rw[r] = (acc > T(SSTTS_F32_TINY) ? T(1) / acc : T(1)) * inv_nfft;
  }
}

// n_fft x overlap-add sum -> waveform sample at padded coordinate p (divide by the window sum
// where it exceeds tiny, as librosa.istft does).
template <typename T>
SSTTS_D T normalise_ola(T v, int p, int n_frames, int hop, int win, int lpad, const T* s_win,
                        const T* s_rw, T inv_nfft) {
  const int x = p - lpad;
  if (x >= win - hop && x / hop <= n_frames - 1) return v * s_rw[x % hop];
  const T wss = window_sumsq<T>(p, n_frames, hop, win, lpad, s_win);
  v *= inv_nfft;
  return wss > T(SSTTS_F32_TINY) ? v / wss : v;
}

// =============================================================================================
// Griffin-Lim
// =============================================================================================
// frames [a, b) of utterance utt, b - a <= warps per CTA.  The record carries the utterance's offsets so that a
// kernel needs ONE independent 48-byte load per tile instead of a tile -> utterance -> offsets chain.
struct alignas(16) GLTile {
  int utt, a, b, parity;
  long long f0, poff;        // frame_off[utt], pad_off[utt]
  int n_frames, reserved;    // frames of the utterance
  long long soff;            // sample_off[utt]
};

template <typename T> struct GLArgs {
  const float* mag;              // (sum T, 1025) frame-major |S|
  const float2* phase0;          // (sum T, bins) initial unit phasors (FROM_PHASE launch only), or nullptr:
  unsigned long long phase_seed; //   then element i = row * bins + bin gets seeded_phasor(phase_seed,
  long long phase_first;         //   phase_first + i)
  const T* pin0; const T* pin1;  // partial overlap-add sums of the previous step, by tile parity
  T* pout0; T* pout1;            // ... of this step
  const long long* frame_off;    // [n_utts + 1]
  const long long* pad_off;      // [n_utts + 1] offsets of each utterance's padded axis
  const GLTile* tiles;
  int n_tiles;
  StftTables<T> tab;
  double* mse_frame;             // [sum T] per-frame sum_k (|S| - |E|)^2 (WANT_MSE launch only)
  int win, hop, span_max;
  int n_fft;                     // 2048, or 1024 / 512 embedded in the 2048-point transform
};

template <typename T> SSTTS_D T fast_rsqrt(T x);
template <> SSTTS_D float fast_rsqrt<float>(float x) { return sstts_rsqrt_approx(x); }
template <> SSTTS_D double fast_rsqrt<double>(double x) { return 1.0 / sqrt(x); }

// Unit phasor of (xr, xi) times s;  (1, 0) * s when the bin is exactly zero
// (np.exp(1j * np.angle(0)) == 1, audio/synthesis.py:109).  Branch-free.
template <typename T>
SSTTS_D void replace_magnitude(T xr, T xi, T s, T& yr, T& yi, T& m2) {
  m2 = fma(xr, xr, xi * xi);   // explicit contraction: identical bits in every instantiation
  // |x|^2 below 1e-30 (|x| < 1e-15, far under the float32 FFT noise floor) counts as zero, which
  // lets the float path use the single-instruction flush-to-zero MUFU.RSQ
  const bool nz = m2 > T(1e-30);
  const T inv = nz ? fast_rsqrt<T>(m2) * s : T(0);
  yr = nz ? xr * inv : s;
  yi = xi * inv;
}

// 1: a frame of the iteration kernel whose first sample sits at an odd offset of the staged span is transformed
// ONE SAMPLE LATER in its zero-padded n_fft buffer (position lpad + 1 instead of lpad), so that the packed pairs
// (x[2n], x[2n+1]) are aligned 8 / 16-byte shared-memory loads instead of two conflicting scalar ones.  A circular
// shift only multiplies the spectrum by a phase ramp: |X| is unchanged, E/|E| * |S| carries the same ramp, and the
// inverse transform returns the frame one sample later as well, where the output store and the gather expect it.
// (Exactly-zero bins get the phase (1, 0) in the shifted buffer; an all-zero analysis frame with |S| != 0 does
// not occur.)  Needs lpad + 1 + win <= 2048.
#ifndef SSTTS_GL_FRAME_SHIFT
#define SSTTS_GL_FRAME_SHIFT 1
#endif
// 1: the synthesis window is applied by the overlap-add gather (an FMA with the thread's <= 5 window values, which
// only depend on its hop residue) instead of by the frame's warp before the store: no window loads and multiplies
// on the output side of the transform.
#ifndef SSTTS_GL_WINDOW_IN_GATHER
#define SSTTS_GL_WINDOW_IN_GATHER 1
#endif
// 1: the float32 n_fft 2048 Griffin-Lim kernels exchange (re, im) pairs through a complex transpose plane
// (warp_fft1024_cx) that overlaps the warp's |S| row buffer.
#ifndef SSTTS_GL_COMPLEX_PLANE
#define SSTTS_GL_COMPLEX_PLANE 1
#endif
constexpr int MAX_OVERLAP = 5;  // frames covering one sample: ceil(win / hop) <= 5 (host_plan.h)
constexpr int MAGROW = 1032;  // per-warp staging of one |S| row: 1025 + up to 3 alignment floats
constexpr int NATIVE_MAGROW = 516;  // native n_fft 1024 path: 513 + up to 3 alignment floats per frame (16-byte multiple)

// Asynchronously stage one n-float row (n = 1025, 513 or 257) into shared memory (16-byte LDGSTS for the aligned
// body, plain loads for the <= 3 + 3 ragged elements).  Element e lands at dst[e + mis]; returns
// mis.  Completion: sstts_cp_async_wait_all() + __syncwarp().
SSTTS_D int stage_row_async(float* dst, const float* __restrict__ g, int lane, int n = NBINS) {
  const int mis = (int)((reinterpret_cast<uintptr_t>(g) >> 2) & 3);
  const int c_lo = (mis + 3) >> 2;
  const int c_hi = ((n - 4 + mis) >> 2) + 1;          // chunks c_lo <= c < c_hi lie inside the row
  const float* gal = g - mis;
  for (int c = c_lo + lane; c < c_hi; c += 32) sstts_cp_async16(dst + 4 * c, gal + 4 * c);
  const int head = 4 * c_lo - mis;                     // elements [0, head)
  if (lane < head) sstts_cp_async4(dst + lane + mis, g + lane);
  const int tail0 = 4 * c_hi - mis;                    // elements [tail0, n)
  if (tail0 + lane < n) sstts_cp_async4(dst + tail0 + lane + mis, g + tail0 + lane);
  return mis;
}

// Counter-based initial phase of the batched API (replaces np.random.rand at audio/synthesis.py:85):
// element index -> two rounds of the murmur3 32-bit finaliser over (index, seed) -> 24-bit uniform u
// -> exp(2 pi i (u - 1/2)) with the SFU sine / cosine (|error| ~ 5e-7: the phase only has to be
// random).  Shared by the stand-alone generator (sstts_random_phase_at) and the synthesis launch
// that draws the phase itself, so both give the same bits.
SSTTS_D unsigned fmix32(unsigned h) {
  h ^= h >> 16; h *= 0x85ebca6bu;
  h ^= h >> 13; h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
SSTTS_D float2 seeded_phasor(unsigned long long seed, long long index) {
  const unsigned lo = (unsigned)index, hi = (unsigned)((unsigned long long)index >> 32);
  unsigned h = fmix32(lo ^ (unsigned)seed);
  h = fmix32(h ^ hi ^ (unsigned)(seed >> 32) ^ 0x9e3779b9u);
  const float a = ((float)(h >> 8) * (1.0f / 16777216.0f) - 0.5f) * 6.283185307179586f;   // [-pi, pi)
  float2 r; r.x = sstts_cos_approx(a); r.y = sstts_sin_approx(a);
  return r;
}

// The per-frame core of one Griffin-Lim step, on the packed spectrum held by the warp.
//   in  (FROM_PHASE = false): re/im = Z = FFT1024(z), element 32 r + lane in slot r
//   out: re/im = 2 Z' (packed spectrum of |S| * E/|E|), same slots
// Conjugate-pair identities (N = 1024, w = exp(-2 pi i k / 2048), Zn = Z[N-k]):
//   2 X[k]   = (Zk + conj Zn) + w * (-i)(Zk - conj Zn)
//   2 X[N-k] = conj((Zk + conj Zn) - w * (-i)(Zk - conj Zn))
//   2 Z'[k]   = (Yk + conj Yn) + i conj(w) (Yk - conj Yn)          (Y = |S| X/|X|)
//   2 Z'[N-k] = conj(Yk + conj Yn) + i conj(conj(w) (Yk - conj Yn))
// Lane l owns k = l + 32 k2; for k2 < 16 its partner bin N - k sits in lane (32 - l) & 31,
// slot 31 - k2 (lane 0: its own slot 32 - k2; k = 0 pairs with the Nyquist bin; k = 512 is
// self-conjugate).  Each lane processes its 16 low pairs and swaps results with its partner.
// srow[k] is |S| of this frame (shared memory for the iteration kernel, global for the synth).
// bshift > 0: a shorter transform (n_fft = 2048 >> bshift) embedded in the 2048-point one -- only
// every (1 << bshift)-th bin exists (srow / prow are indexed by k >> bshift), the others are forced
// to zero, which makes the inverse transform periodic with period n_fft (its first period is used).
// OUT_BREV: the result is written to (ro, io) in BIT-REVERSED slots (what a decimation-in-time first
// pass of the inverse transform takes), otherwise in natural slots.  Slot indices are compile-time, so
// the permutation is only a renaming of registers; ro / io may not alias re / im.
template <typename T, bool FROM_PHASE, bool WANT_MSE, bool OUT_BREV>
SSTTS_D void gl_frame_core(T (&re)[32], T (&im)[32], T (&ro)[32], T (&io)[32], const float* srow,
                           const float2* __restrict__ prow,
                           const typename cx_of<T>::type* s_w2k, int lane, double& mse_acc, int bshift = 0,
                           unsigned long long phase_seed = 0, long long phase_base = 0) {
  typedef typename cx_of<T>::type C;
  const int partner = (32 - lane) & 31;
  const int bmask = (1 << bshift) - 1;
  const bool real_bin = (lane & bmask) == 0;   // k = lane + 32 k2 and 1024 - k share lane's residue
  T prev_r = T(0), prev_i = T(0);              // 2 Z'[N-k] received in the previous pair step
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const int sl_mine = k2;
    const int sl_part = 31 - k2;
    const int sl_alt = (32 - k2) & 31;
    const int k = lane + 32 * k2;
    const int kn = HALF - k;
    const C w = s_w2k[k];
    const T sk = real_bin ? fabs((T)srow[k >> bshift]) : T(0);
    const T sn = real_bin ? fabs((T)srow[kn >> bshift]) : T(0);
    T ykr, yki, ynr, yni;
    if (!FROM_PHASE) {
      const T zr = re[sl_mine], zi = im[sl_mine];
      // lane 0 is its own partner and pairs slot k2 with slot 32 - k2: select the source, then shuffle
      const T pr = __shfl_sync(0xffffffffu, lane == 0 ? re[sl_alt] : re[sl_part], partner);
      const T pi = __shfl_sync(0xffffffffu, lane == 0 ? im[sl_alt] : im[sl_part], partner);
      const T er = zr + pr, ei = zi - pi;   // Zk + conj Zn
      const T dr = zr - pr, di = zi + pi;   // Zk - conj Zn
      const T wor = w.x * di + w.y * dr;    // w * (di - i dr)
      const T woi = w.y * di - w.x * dr;
      const T xkr = er + wor, xki = ei + woi;      // 2 X[k]
      const T xnr = er - wor, xni = woi - ei;      // 2 X[N-k]
      T m2k, m2n;
      replace_magnitude<T>(xkr, xki, sk, ykr, yki, m2k);
      replace_magnitude<T>(xnr, xni, sn, ynr, yni, m2n);
      if (WANT_MSE && real_bin) {
        const double ek = (double)sk - 0.5 * sqrt((double)m2k);
        const double en = (double)sn - 0.5 * sqrt((double)m2n);
        mse_acc += ek * ek + en * en;
      }
    } else {
      float2 pk = make_float2(0.0f, 0.0f), pn = make_float2(0.0f, 0.0f);
      if (real_bin) {
        if (prow) { pk = prow[k >> bshift]; pn = prow[kn >> bshift]; }
        else {   // prow == nullptr: the phase is drawn here (element index = phase_base + bin)
          pk = seeded_phasor(phase_seed, phase_base + (k >> bshift));
          pn = seeded_phasor(phase_seed, phase_base + (kn >> bshift));
        }
      }
      ykr = sk * (T)pk.x; yki = sk * (T)pk.y;
      ynr = sn * (T)pn.x; yni = sn * (T)pn.y;
      if (lane == 0 && k2 == 0) { yki = T(0); yni = T(0); }  // ifft(...).real drops Im of DC/Nyquist
    }
    const T e2r = ykr + ynr, e2i = yki - yni;   // Yk + conj Yn
    const T d2r = ykr - ynr, d2i = yki + yni;   // Yk - conj Yn
    const T o2r = w.x * d2r + w.y * d2i;        // conj(w) * (Yk - conj Yn)
    const T o2i = w.x * d2i - w.y * d2r;
    const T zkr = e2r - o2i, zki = e2i + o2r;   // 2 Z'[k]
    const T znr = e2r + o2i, zni = o2r - e2i;   // 2 Z'[N-k]
    const T rr = __shfl_sync(0xffffffffu, znr, partner);
    const T ri = __shfl_sync(0xffffffffu, zni, partner);
    // slot 32 - k2 (k2 >= 1) takes the partner's value of the PREVIOUS step for lanes >= 1 and lane 0's
    // own value of this step (lane 0's shuffle returns its own znr); slot 16 is completed below
    ro[OUT_BREV ? brev5(sl_mine) : sl_mine] = zkr; io[OUT_BREV ? brev5(sl_mine) : sl_mine] = zki;
    if (k2 != 0) {
      ro[OUT_BREV ? brev5(sl_alt) : sl_alt] = lane == 0 ? rr : prev_r;
      io[OUT_BREV ? brev5(sl_alt) : sl_alt] = lane == 0 ? ri : prev_i;
    }
    prev_r = rr; prev_i = ri;
  }
  // slot 16: lanes >= 1 take the partner's value of the last step; lane 0 holds k = 512 there
  // (self-conjugate): X = conj(Z), Z' = conj(Y).
  ro[OUT_BREV ? brev5(16) : 16] = prev_r; io[OUT_BREV ? brev5(16) : 16] = prev_i;
  if (lane == 0) {
    const int sl = 16;
    const T s = fabs((T)srow[(HALF / 2) >> bshift]);
    T yr, yi;
    if (!FROM_PHASE) {
      T m2;
      replace_magnitude<T>(T(2) * re[sl], T(-2) * im[sl], s, yr, yi, m2);
      if (WANT_MSE) {
        const double e = (double)s - 0.5 * sqrt((double)m2);
        mse_acc += e * e;
      }
    } else {
      const float2 p = prow ? prow[(HALF / 2) >> bshift] : seeded_phasor(phase_seed, phase_base + ((HALF / 2) >> bshift));
      yr = s * (T)p.x; yi = s * (T)p.y;
    }
    ro[OUT_BREV ? brev5(sl) : sl] = T(2) * yr; io[OUT_BREV ? brev5(sl) : sl] = T(-2) * yi;
  }
}

// The same core for the native n_fft = 1024 path: the 16 lanes of a half-warp hold the packed spectrum of one
// frame as halfwarp_fft512 leaves it -- slot s of lane hl is Z[k], k = hl + (s & 16) + 32 (s & 15) -- and
// N = 512, w = exp(-2 pi i k / 1024) in the identities above.  The partner Z[512 - k] of a lower-set bin
// (slot j < 16) of lane hl >= 1 is in slot 31 - j of lane 16 - hl of the same half; lane hl == 0 holds both
// members of its pairs (slots (j, 16 - j) and (16 + j, 31 - j)), the DC / Nyquist pair in slot 0 and the
// self-conjugate bin 256 in slot 8.  Every lane processes 16 pairs and hands 2 Z'[512 - k] to its partner.
// Output: 2 Z' in slot (s & 16) + brev4(s & 15), the order halfwarp_ifft512 takes.
template <typename T, bool FROM_PHASE, bool WANT_MSE>
SSTTS_D void gl_frame_core_native(T (&re)[32], T (&im)[32], T (&ro)[32], T (&io)[32], const float* srow,
                                  const float2* __restrict__ prow, const typename cx_of<T>::type* s_w2k, int lane,
                                  double& mse_acc, unsigned long long phase_seed = 0, long long phase_base = 0) {
  typedef typename cx_of<T>::type C;
  const int hl = lane & 15;
  const int partner = (lane & 16) | ((16 - hl) & 15);
  const bool h0 = hl == 0;
  T zkr_[16], zki_[16], rr_[16], ri_[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int sa = j < 8 ? j : 8 + j;                      // lane 0's own pairs
    const int sb = j == 0 ? 0 : (j < 8 ? 16 - j : 39 - j);
    const int k = !h0 ? hl + 32 * j : (j < 8 ? 32 * j : 32 * j - 240);
    const int kn = 512 - k;
    const C w = s_w2k[k];
    const T sk = fabs((T)srow[k]);
    const T sn = fabs((T)srow[kn]);
    T ykr, yki, ynr, yni;
    if (!FROM_PHASE) {
      const T zr = (j >= 8 && h0) ? re[sa] : re[j];
      const T zi = (j >= 8 && h0) ? im[sa] : im[j];
      const T pr = __shfl_sync(0xffffffffu, h0 ? re[sb] : re[31 - j], partner);
      const T pi = __shfl_sync(0xffffffffu, h0 ? im[sb] : im[31 - j], partner);
      const T er = zr + pr, ei = zi - pi;   // Zk + conj Zn
      const T dr = zr - pr, di = zi + pi;   // Zk - conj Zn
      const T wor = w.x * di + w.y * dr;    // w * (di - i dr)
      const T woi = w.y * di - w.x * dr;
      const T xkr = er + wor, xki = ei + woi;      // 2 X[k]
      const T xnr = er - wor, xni = woi - ei;      // 2 X[512 - k]
      T m2k, m2n;
      replace_magnitude<T>(xkr, xki, sk, ykr, yki, m2k);
      replace_magnitude<T>(xnr, xni, sn, ynr, yni, m2n);
      if (WANT_MSE) {
        const double ek = (double)sk - 0.5 * sqrt((double)m2k);
        const double en = (double)sn - 0.5 * sqrt((double)m2n);
        mse_acc += ek * ek + en * en;
      }
    } else {
      float2 pk, pn;
      if (prow) { pk = prow[k]; pn = prow[kn]; }
      else {   // prow == nullptr: the phase is drawn here (element index = phase_base + bin)
        pk = seeded_phasor(phase_seed, phase_base + k);
        pn = seeded_phasor(phase_seed, phase_base + kn);
      }
      ykr = sk * (T)pk.x; yki = sk * (T)pk.y;
      ynr = sn * (T)pn.x; yni = sn * (T)pn.y;
      if (h0 && j == 0) { yki = T(0); yni = T(0); }  // ifft(...).real drops Im of DC/Nyquist
    }
    const T e2r = ykr + ynr, e2i = yki - yni;   // Yk + conj Yn
    const T d2r = ykr - ynr, d2i = yki + yni;   // Yk - conj Yn
    const T o2r = w.x * d2r + w.y * d2i;        // conj(w) * (Yk - conj Yn)
    const T o2i = w.x * d2i - w.y * d2r;
    zkr_[j] = e2r - o2i; zki_[j] = e2i + o2r;   // 2 Z'[k]
    const T znr = e2r + o2i, zni = o2r - e2i;   // 2 Z'[512 - k]
    rr_[j] = __shfl_sync(0xffffffffu, znr, partner);   // lane hl == 0 is its own partner
    ri_[j] = __shfl_sync(0xffffffffu, zni, partner);
  }
  // self-conjugate bin 256 (lane hl == 0, slot 8): X = conj(Z), Z' = conj(Y)
  T scr = T(0), sci = T(0);
  {
    const T s = fabs((T)srow[256]);
    T yr, yi;
    if (!FROM_PHASE) {
      T m2;
      replace_magnitude<T>(T(2) * re[8], T(-2) * im[8], s, yr, yi, m2);
      if (WANT_MSE && h0) {
        const double e = (double)s - 0.5 * sqrt((double)m2);
        mse_acc += e * e;
      }
    } else {
      const float2 p = prow ? prow[256] : seeded_phasor(phase_seed, phase_base + 256);
      yr = s * (T)p.x; yi = s * (T)p.y;
    }
    scr = T(2) * yr; sci = T(-2) * yi;
  }
  // assemble the slots: lanes hl >= 1 keep 2 Z'[k] of step j in slot j and receive slot 31 - j; lane hl == 0
  // keeps step j in slot sa(j) and its own 2 Z'[512 - k] in slot sb(j)
#pragma unroll
  for (int s = 0; s < 32; ++s) {
    const int os = (s & 16) + brev4(s & 15);
    T vr, vi;
    if (s < 8) { vr = zkr_[s]; vi = zki_[s]; }
    else if (s == 8) { vr = h0 ? scr : zkr_[8]; vi = h0 ? sci : zki_[8]; }
    else if (s < 16) { vr = h0 ? rr_[16 - s] : zkr_[s]; vi = h0 ? ri_[16 - s] : zki_[s]; }
    else if (s < 24) { vr = h0 ? zkr_[s - 8] : rr_[31 - s]; vi = h0 ? zki_[s - 8] : ri_[31 - s]; }
    else { vr = h0 ? rr_[39 - s] : rr_[31 - s]; vi = h0 ? ri_[39 - s] : ri_[31 - s]; }
    ro[os] = vr; io[os] = vi;
  }
}

// Windowed, packed frame into the registers of a DIT first pass (bit-reversed slots): element n1 of lane l is
// z = (x[m] w[m - lpad], x[m + 1] w[m + 1 - lpad]), m = STRIDE n1 + 2 l (STRIDE 64: one frame per warp, 32: one per
// half-warp), zero outside the window.  PAIRED: fin + m is aligned for a two-sample load (the caller checked that
// the frame starts at an even element of the 8-byte aligned buffer), else two scalar loads.
template <typename T, typename S, int STRIDE, bool PAIRED>
SSTTS_D void load_windowed_frame(T (&re)[32], T (&im)[32], const S* fin, const T* s_win, int lpad, int win, int l) {
  typedef typename cx_of<T>::type C;
  typedef typename cx_of<S>::type S2;
#pragma unroll
  for (int n1 = 0; n1 < 32; ++n1) {
    const int m = STRIDE * n1 + 2 * l;
    const int i = m - lpad;
    C w2; w2.x = T(0); w2.y = T(0);
    S xa = S(0), xb = S(0);
    if (i + 1 >= 0 && i < win) {
      SSTTS_CHECK_ALIGNED(s_win + i, sizeof(C));
      w2 = *reinterpret_cast<const C*>(s_win + i);       // window pair (zero outside the window) in one aligned load
      if (PAIRED) { SSTTS_CHECK_ALIGNED(fin + m, sizeof(S2)); const S2 v = *reinterpret_cast<const S2*>(fin + m); xa = v.x; xb = v.y; }
      else { if (i >= 0) xa = fin[m]; if (i + 1 < win) xb = fin[m + 1]; }
    }
    re[brev5(n1)] = (i >= 0 && i < win) ? (T)xa * w2.x : T(0);
    im[brev5(n1)] = (i + 1 >= 0 && i + 1 < win) ? (T)xb * w2.y : T(0);
  }
}

// Shared-memory carve-up of the Griffin-Lim step kernel.
// Two staging variants of the iteration kernel (template parameter BULK of gl_step_kernel):
//   BULK = false  every thread loads, normalises and stores its samples of the next tile after the gather
//                 (the span was pulled into L2 by prefetch hints while the tile was being transformed);
//   BULK = true   interior tiles take their input span with bulk asynchronous copies (cp.async.bulk +
//                 mbarrier, the 1-D form of TMA) issued by ONE thread while the CTA is still transforming
//                 the previous tile; the normalisation is folded into the window table.
// Measured on B200 (profiles/experiments/r2_ab_bulk_staging.txt): 0.575 ms (bulk) vs 0.572 ms per iteration
// launch on the 256-utterance batch, 0.916 vs 0.845 ms for one 1000-frame utterance (one tile per CTA: the
// copy latency is exposed) -- the kernel is bound by issue slots and the shared-memory pipe, not by this
// phase, so the library runs BULK = false unless SSTTS_GL_STAGING=bulk is set (sstts.cu).
template <typename T> struct GLSmem {
  typedef typename cx_of<T>::type C;
  int plane_elems;   // per-warp region: transpose tile(s), later the windowed output frame(s)
  int frame_pitch;   // distance between the output frames of consecutive frames of the tile
  int edge_elems;    // one neighbour edge region ((win - hop) samples + alignment slack)
  int mag_in_plane = 0;   // > 0: offset (elements) of the warp's |S| row inside its region (complex-plane layout)
  size_t off_w2k, off_win, off_win2, off_wr, off_rw, off_plane, off_mag, off_yin, off_edge, off_bar, total;
  // native: the n_fft = 1024 path with two frames per warp.  A warp's region then holds, one after the other in
  // time, the two half-warp transpose planes (2 x HPLANE_ELEMS) next to the two staged |S| rows
  // (2 x NATIVE_MAGROW floats), and -- once the core has consumed the rows and the inverse transform its
  // planes -- the two windowed output frames (2 x frame_pitch), which the gather then reads.
  SSTTS_HD GLSmem(int warps, int win, int hop, int span_max, bool bulk, bool native = false) {
    if (native) {
      const size_t work = sizeof(T) * 2 * HPLANE_ELEMS + sizeof(float) * 2 * NATIVE_MAGROW;
      const int need = (int)((work + 2 * sizeof(T) - 1) / (2 * sizeof(T)));
      frame_pitch = round_up4(win + 2) > need ? round_up4(win + 2) : round_up4(need);
      plane_elems = 2 * frame_pitch;
    } else {
      plane_elems = round_up4(win + 2) > XPLANE_ELEMS ? round_up4(win + 2) : round_up4(XPLANE_ELEMS);
      mag_in_plane = 0;
      if (SSTTS_GL_COMPLEX_PLANE && sizeof(T) == 4) {
        // float32: the |S| row is staged right behind the frame part of the warp's region, and the COMPLEX
        // transpose plane (2 x XPLANE_ELEMS floats) spans both -- the row is copied in after the forward
        // transform's exchange and is dead before the inverse transform's
        mag_in_plane = plane_elems;
        plane_elems += round_up4(XPLANE_ELEMS);
      }
      frame_pitch = plane_elems;
    }
    edge_elems = round_up4(win - hop > 0 ? win - hop : 0) + 8;
    size_t o = sizeof(C) * (native ? 512 : 1024);
    off_w2k = o; o += sizeof(C) * 512;
    off_win = o; o += sizeof(T) * round_up4(win + WIN_TAB_PAD);
    off_win2 = o; if (SSTTS_GL_FRAME_SHIFT && !native && !bulk) o += sizeof(T) * round_up4(win + WIN_TAB_PAD);   // other pair alignment
    off_wr = o; if (bulk) o += sizeof(T) * round_up4(win + WIN_TAB_PAD);
    off_rw = o; o += sizeof(T) * round_up4(hop);
    off_plane = o; o += sizeof(T) * (size_t)warps * plane_elems;
    off_mag = o; if (!native && !mag_in_plane) o += sizeof(float) * (size_t)warps * MAGROW;
    off_yin = o; o += sizeof(T) * (round_up4(span_max) + 8);      // + slack for the 16-byte alignment shift
    off_edge = o; if (bulk) o += sizeof(T) * 2 * (size_t)edge_elems;
    off_bar = o; o += 16;                                          // mbarrier + arrival counter
    total = o;
  }
};

#ifndef SSTTS_CORE_BREV_OUT
#define SSTTS_CORE_BREV_OUT 1
#endif
#ifndef SSTTS_STAGE_PREFETCH
#define SSTTS_STAGE_PREFETCH 1
#endif
// One Griffin-Lim step over all tiles.  FROM_PHASE = true is the initial synthesis from the
// random phase (no analysis half).  W warps per CTA, one frame per warp, tiles of <= W frames -- or, with a
// native n_fft 1024 geometry (G::kNative1024), two frames per warp (a 512-point complex transform per half-warp)
// and tiles of <= 2 W frames.
template <typename T, typename G, int W, bool FROM_PHASE, bool WANT_MSE, bool USE_BULK = false>
__global__ void __launch_bounds__(W * 32, sizeof(T) == 4 ? 2 : 1) gl_step_kernel(const GLArgs<T> A) {
  typedef typename cx_of<T>::type C;
  const G g(A.win, A.hop, A.n_fft);
  const int win = g.win(), hop = g.hop(), lpad = g.lpad(), cpad = g.cpad();
  const int bshift = g.bin_shift();
  const int n_bins = (HALF >> bshift) + 1;
  const T inv_nfft = T(1.0) / T(g.nfft());
  const int mlo = lpad & ~1;          // even base of the output-frame slot (8-byte stores)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NT = W * 32;
  constexpr bool NATIVE = G::kNative1024;
  constexpr int F = NATIVE ? 2 * W : W;          // frames per tile
  static_assert(!(NATIVE && USE_BULK), "the bulk-copy staging variant exists for the 2048-point kernels only");

  SSTTS_DYN_SMEM(smem);
  const GLSmem<T> L(W, win, hop, A.span_max, USE_BULK, NATIVE);
  C* s_tw = reinterpret_cast<C*>(smem);
  C* s_w2k = reinterpret_cast<C*>(smem + L.off_w2k);
  T* s_win = load_window_table<T>(reinterpret_cast<T*>(smem + L.off_win), A.tab.window, win, lpad, tid, W * 32);
  T* s_rw = reinterpret_cast<T*>(smem + L.off_rw);
  T* s_planes = reinterpret_cast<T*>(smem + L.off_plane);
  T* s_yin = reinterpret_cast<T*>(smem + L.off_yin);
  T* plane = s_planes + warp * L.plane_elems;
  constexpr bool CXP = (SSTTS_GL_COMPLEX_PLANE != 0) && sizeof(T) == 4 && !NATIVE;
  float* s_mag = NATIVE ? reinterpret_cast<float*>(plane + 2 * HPLANE_ELEMS)
                 : CXP  ? reinterpret_cast<float*>(plane + L.mag_in_plane)
                        : reinterpret_cast<float*>(smem + L.off_mag) + warp * MAGROW;
  constexpr bool BULK = !FROM_PHASE && USE_BULK;
  // frames at odd offsets of the span are transformed one sample later (see SSTTS_GL_FRAME_SHIFT)
  constexpr bool SHIFT = (SSTTS_GL_FRAME_SHIFT != 0) && !FROM_PHASE && !USE_BULK && !NATIVE;
  const bool shift_ok = SHIFT && lpad + 1 + win <= NFFT;
  auto frame_shift = [&](int f) -> int { return (SHIFT && shift_ok) ? ((f * hop - lpad) & 1) : 0; };
  T* s_win2 = s_win;   // window table with the other pair alignment (for lpad + 1)
  if constexpr (SHIFT) s_win2 = load_window_table<T>(reinterpret_cast<T*>(smem + L.off_win2), A.tab.window, win, lpad + 1, tid, W * 32);
  T* s_wrtab = reinterpret_cast<T*>(smem + L.off_wr);           // window x reciprocal window sum (BULK)
  T* s_wr = s_wrtab + win_shift(lpad);
  T* s_edge = reinterpret_cast<T*>(smem + L.off_edge);          // raw neighbour sums of the two edge regions
  sstts_mbar_t* s_bar = reinterpret_cast<sstts_mbar_t*>(smem + L.off_bar);
  int* s_cnt = reinterpret_cast<int*>(smem + L.off_bar + 8);    // warps that have consumed s_yin this round

  // (the complex-plane transform reads its twiddles two at a time: paired table layout)
  for (int i = tid; i < (NATIVE ? 512 : 1024); i += NT) s_tw[CXP ? paired_twiddle_index(i >> 5, i & 31) : i] = A.tab.tw1024[i];
  for (int i = tid; i < 512; i += NT) s_w2k[i] = A.tab.w2048[i];
  if (BULK && tid == 0) { sstts_mbar_init(s_bar, 1); *s_cnt = 0; }
  __syncthreads();
  fill_interior_rwss<T>(s_rw, s_win, hop, win, tid, NT, inv_nfft);
  __syncthreads();
  if (BULK) {
    // interior samples: x = (own + neighbour) * rw[(m - lpad) mod hop], then * window[m - lpad] in the window
    // load; the residue only depends on the position inside the frame, so both factors fold into one table
    for (int i = tid; i < round_up4(win + WIN_TAB_PAD); i += NT) {
      const int j = i - win_shift(lpad);
      s_wrtab[i] = (j >= 0 && j < win) ? s_win[j] * s_rw[j % hop] : T(0);
    }
  }

  // Tile records are single independent 48-byte loads, issued two rounds ahead.
  struct TileCtx { int a, b, parity, n_frames; long long f0, poff; };
  auto load_ctx = [&](int t) {
    TileCtx c;
    const GLTile tl = A.tiles[t];
    c.a = tl.a; c.b = tl.b; c.parity = tl.parity;
    c.f0 = tl.f0; c.n_frames = tl.n_frames; c.poff = tl.poff;
    return c;
  };

  // common case ("plain" tile): no reflection inside the span and every sample covered by a full set of frames
  auto is_plain = [&](const TileCtx& tl) -> bool {
    const int L_out = hop * (tl.n_frames - 1);
    const int span_lo = tl.a * hop + lpad;
    const int span = (tl.b - tl.a - 1) * hop + win;
    return (span_lo >= cpad) && (span_lo + span - cpad <= L_out) && (tl.a * hop >= win - hop) &&
           ((tl.a * hop + span - 1) / hop <= tl.n_frames - 1);
  };
  // BULK: one thread hands the span of a plain tile to the copy engine -- the tile's own-parity sums into
  // s_yin and the neighbour-parity sums of the two edge regions into s_edge -- and arms the barrier with
  // the byte count.  Sources are rounded down to 16 bytes: sample s of the span lands at s_yin[s + mis],
  // left-edge sample s at s_edge[s + mis], right-edge sample s (>= rb) at s_edge[edge_elems + s - rb + mis_r].
  auto bulk_issue = [&](const TileCtx& tl) {
    const int span_lo = tl.a * hop + lpad;
    const int span = (tl.b - tl.a - 1) * hop + win;
    const T* own = (tl.parity ? A.pin1 : A.pin0) + tl.poff + span_lo;
    const T* oth = (tl.parity ? A.pin0 : A.pin1) + tl.poff + span_lo;
    const int mis = span_lo & 3;                       // poff and the buffers are 16-byte aligned
    const int le = win - hop, rb = (tl.b - tl.a) * hop;
    const int mis_r = (span_lo + rb) & 3;
    const unsigned b_own = (unsigned)sizeof(T) * (unsigned)round_up4(span + mis);
    const unsigned b_le = (unsigned)sizeof(T) * (unsigned)round_up4(le + mis);
    const unsigned b_re = (unsigned)sizeof(T) * (unsigned)round_up4(span - rb + mis_r);
    sstts_fence_proxy_async();
    sstts_mbar_arrive_expect_tx(s_bar, b_own + b_le + b_re);
    sstts_bulk_g2s(s_yin, own - mis, b_own, s_bar);
    sstts_bulk_g2s(s_edge, oth - mis, b_le, s_bar);
    sstts_bulk_g2s(s_edge + L.edge_elems, oth + rb - mis_r, b_re, s_bar);
    sstts_mbar_phase_done(s_bar);
  };
  // ... and everybody completes it: wait for the bytes, add the neighbour's contribution inside the two edge
  // regions (the normalisation is folded into the window table, see s_wr above)
  unsigned bar_parity = 0;
  auto bulk_finish = [&](const TileCtx& tl) {
    const int span_lo = tl.a * hop + lpad;
    const int span = (tl.b - tl.a - 1) * hop + win;
    const int mis = span_lo & 3;
    const int le = win - hop, rb = (tl.b - tl.a) * hop;
    const int mis_r = (span_lo + rb) & 3;
    sstts_mbar_wait(s_bar, bar_parity);
    bar_parity ^= 1u;
    for (int s = tid; s < le; s += NT) s_yin[s + mis] += s_edge[s + mis];
    const T* er = s_edge + L.edge_elems + mis_r - rb;
    for (int s = rb + tid; s < span; s += NT) s_yin[s + mis] += er[s];
  };

  // Stage the analysis input of a tile: x_pad[span] = y_norm[reflect], y_norm = OLA sum / wss.
  auto stage = [&](const TileCtx& tl) {
    const int n_frames = tl.n_frames;
    const long long poff = tl.poff;
    const int a = tl.a, b = tl.b;
    const int L_out = hop * (n_frames - 1);
    const int span_lo = a * hop + lpad;
    const int span = (b - a - 1) * hop + win;
    const T* pin_own = (tl.parity ? A.pin1 : A.pin0) + poff;
    const T* pin_oth = (tl.parity ? A.pin0 : A.pin1) + poff;
    // tile-relative edge regions: below `le` the previous tile's frames reach in, from `rb` on the
    // next tile's (disjoint because tiles hold >= min_tile frames)
    const int le = a > 0 ? win - hop : 0;
    const int rb = b < n_frames ? (b - a) * hop : 0x7fffffff;
    const bool plain = is_plain(tl);
    if (plain) {
      const T* own = pin_own + span_lo;
      const T* oth = pin_oth + span_lo;
      // all global loads of the thread first (independent, in flight together), then the stores;
      // residue of sample s: (span_lo - lpad + s) mod hop = s mod hop since span_lo - lpad = a hop
      constexpr int NS = F + MAX_OVERLAP - 1;
      T x[NS], e[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int s = tid + i * NT;
        x[i] = s < span ? own[s] : T(0);
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int s = tid + i * NT;
        e[i] = (s < span && (s < le || s >= rb)) ? oth[s] : T(0);
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const int s = tid + i * NT;
        if (s < span) s_yin[s] = (x[i] + e[i]) * s_rw[s % hop];
      }
      for (int s = tid + NS * NT; s < span; s += NT) {   // spans longer than NS * NT (run-time geometry)
        T v = own[s];
        if (s < le || s >= rb) v += oth[s];
        s_yin[s] = v * s_rw[s % hop];
      }
    } else {
      for (int s = tid; s < span; s += NT) {
        int q = span_lo + s - cpad;
        if (q < 0 || q >= L_out) q = reflect_index(q, L_out);
        const int p = q + cpad;
        const int sp = p - span_lo;      // reflected position relative to this tile's span
        T v = pin_own[p];
        if (sp < le || sp >= rb) v += pin_oth[p];
        s_yin[s] = normalise_ola<T>(v, p, n_frames, hop, win, lpad, s_win, s_rw, inv_nfft);
      }
    }
  };

  int tile = blockIdx.x;
  // records of this round's tile, the next one and the one after: a record is loaded two rounds
  // before its first use, so no warp ever waits for it
  TileCtx cur = {0, 0, 0, 0, 0, 0}, nxt = {0, 0, 0, 0, 0, 0}, nxt2 = {0, 0, 0, 0, 0, 0};
  if (tile < A.n_tiles) cur = load_ctx(tile);
  if (tile + (int)gridDim.x < A.n_tiles) nxt = load_ctx(tile + gridDim.x);
  bool cur_bulk = false;          // the current tile's s_yin holds raw sums at an alignment offset (BULK path)
  if (!FROM_PHASE && tile < A.n_tiles) {
    cur_bulk = BULK && is_plain(cur);
    if (cur_bulk) {
      __syncthreads();            // s_wr and the barrier initialisation are visible
      if (tid == 0) bulk_issue(cur);
      bulk_finish(cur);
    } else {
      stage(cur);
    }
  }
  __syncthreads();

  while (tile < A.n_tiles) {
    const int next = tile + gridDim.x;
    if (next + (int)gridDim.x < A.n_tiles) nxt2 = load_ctx(next + gridDim.x);
    const long long f0 = cur.f0;
    const long long poff = cur.poff;
    const int a = cur.a, FT = cur.b - cur.a;
    const int span_lo = a * hop + lpad;
    const int span = (FT - 1) * hop + win;
    T* pout_own = (cur.parity ? A.pout1 : A.pout0) + poff;

    if constexpr (NATIVE) {
      // two frames per warp: half-warp h transforms frame 2 warp + h of the tile (an odd tile's last half redoes
      // its sibling's frame and stores nothing)
      if (2 * warp < FT) {
        const int half = lane >> 4, hl = lane & 15;
        const bool live = 2 * warp + half < FT;
        const int fr = live ? 2 * warp + half : 2 * warp;
        const long long row = f0 + a + fr;
        const float2* prow = (FROM_PHASE && A.phase0) ? A.phase0 + row * n_bins : nullptr;
        // |S| rows of both frames: asynchronous copies by the whole warp, consumed by the core
        stage_row_async(s_mag, A.mag + (f0 + a + 2 * warp) * n_bins, lane, n_bins);
        if (2 * warp + 1 < FT) stage_row_async(s_mag + NATIVE_MAGROW, A.mag + (f0 + a + 2 * warp + 1) * n_bins, lane, n_bins);
        const float* srow = s_mag + (fr - 2 * warp) * NATIVE_MAGROW +
                            (int)((reinterpret_cast<uintptr_t>(A.mag + row * n_bins) >> 2) & 3);
        T re[32], im[32];
        if (!FROM_PHASE) {
          const T* fin = s_yin + fr * hop - lpad;                // fin[m], m in [lpad, lpad + win)
          // frames at even offsets of the span (all of them at hop 256 / win 1024) load their sample pairs aligned
          if (((fr * hop - lpad) & 1) == 0) load_windowed_frame<T, T, 32, true>(re, im, fin, s_win, lpad, win, hl);
          else load_windowed_frame<T, T, 32, false>(re, im, fin, s_win, lpad, win, hl);
          if (SSTTS_STAGE_PREFETCH && next < A.n_tiles) {        // see the 2048-point branch below
            const int nspan = (nxt.b - nxt.a - 1) * hop + win;
            const int lines = (nspan * (int)sizeof(T) + 127) / 128 + 1;
            const long long base = nxt.poff + nxt.a * hop + lpad;
            for (int t = tid; t < 2 * lines; t += NT) {
              const int line = t < lines ? t : t - lines;
              const T* src = ((t < lines) == (nxt.parity != 0) ? A.pin1 : A.pin0) + base;
              sstts_prefetch_l2(reinterpret_cast<const char*>(src) + 128 * line);
            }
          }
          halfwarp_fft512<T>(re, im, plane, half, s_tw, hl);
        }
        sstts_cp_async_wait_all();
        __syncwarp();
        double mse_acc = 0.0;
        T ro[32], io[32];
        gl_frame_core_native<T, FROM_PHASE, WANT_MSE>(re, im, ro, io, srow, prow, s_w2k, lane, mse_acc, A.phase_seed,
                                                      A.phase_first + row * n_bins);
        if (WANT_MSE) {
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) mse_acc += __shfl_xor_sync(0xffffffffu, mse_acc, o);
          if (hl == 0 && live) A.mse_frame[row] = mse_acc;
        }
        halfwarp_ifft512<T>(ro, io, plane, half, s_tw, hl);
        // windowed output frames into the warp's region (transpose planes and |S| rows are dead by now)
        T* oplane = plane + half * L.frame_pitch;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
          const int m = 32 * n1 + 2 * hl;
          const int i = m - lpad;
          if (m >= mlo && m < lpad + win + 1) {
            C vv;
            if (SSTTS_GL_WINDOW_IN_GATHER) { vv.x = ro[n1]; vv.y = io[n1]; (void)i; }
            else {
              const C w2 = *reinterpret_cast<const C*>(s_win + i);   // zeros outside the window
              vv.x = ro[n1] * w2.x; vv.y = io[n1] * w2.y;
            }
            SSTTS_CHECK_ALIGNED(oplane + (m - mlo), sizeof(C));
            *reinterpret_cast<C*>(oplane + (m - mlo)) = vv;        // m - mlo is even
          }
        }
      }
    }
    if constexpr (!NATIVE) if (warp < FT) {
      const long long row = f0 + a + warp;
      const float* mrow = A.mag + row * n_bins;
      const float2* prow = (FROM_PHASE && A.phase0) ? A.phase0 + row * n_bins : nullptr;
      const float* srow = mrow;
      T re[32], im[32];
      const int sh = frame_shift(warp);
      const int lp = lpad + sh;                                   // position of the frame in its n_fft buffer
      if (FROM_PHASE) {
        // the frame's |S| row goes through shared memory here too: one asynchronous copy and one wait
        // instead of 33 dependent global loads inside the pair loop
        const int mis = stage_row_async(s_mag, mrow, lane, n_bins);
        sstts_cp_async_wait_all();
        __syncwarp();
        srow = s_mag + mis;
      } else {
        srow = s_mag + (int)((reinterpret_cast<uintptr_t>(mrow) >> 2) & 3);
        if (!CXP) stage_row_async(s_mag, mrow, lane, n_bins);
        else sstts_prefetch_l2(reinterpret_cast<const char*>(mrow) + 128 * lane);   // copied after the exchange
        // bulk-staged tiles hold raw sums shifted by the alignment offset; their normalisation is in s_wr
        const T* fin = s_yin + (cur_bulk ? ((a * hop + lpad) & 3) : 0) + warp * hop - lp;  // fin[m], m in [lp, lp + win)
        const T* wtab = cur_bulk ? s_wr : (sh ? s_win2 : s_win);
        // with the shift available every frame starts at an even offset of the 16-byte aligned span: (x[m], x[m+1])
        // is one aligned load like the window pair
        const bool paired = SHIFT && shift_ok;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
          const int m = 64 * n1 + 2 * lane;
          const int i = m - lp;
          // window pair (zero outside the window) in one aligned load; DIT pass: bit-reversed slots.
          // Slots entirely inside / outside the window for every lane and both frame positions need no guards
          // (compile-time for the model geometry).
          const int i_min = 64 * n1 - lpad - 1, i_max = 64 * n1 + 62 - lpad;
          if (i_max + 1 < 0 || i_min >= win) {
            re[brev5(n1)] = T(0); im[brev5(n1)] = T(0);
          } else if (paired && i_min >= 0 && i_max + 1 < win) {
            SSTTS_CHECK_ALIGNED(wtab + i, sizeof(C));
            SSTTS_CHECK_ALIGNED(fin + m, sizeof(C));
            const C w2 = *reinterpret_cast<const C*>(wtab + i);
            const C x2 = *reinterpret_cast<const C*>(fin + m);
            re[brev5(n1)] = x2.x * w2.x;
            im[brev5(n1)] = x2.y * w2.y;
          } else {
            C w2; w2.x = T(0); w2.y = T(0);
            C x2; x2.x = T(0); x2.y = T(0);
            if (i + 1 >= 0 && i < win) {
              SSTTS_CHECK_ALIGNED(wtab + i, sizeof(C));
              w2 = *reinterpret_cast<const C*>(wtab + i);
              if (paired) { SSTTS_CHECK_ALIGNED(fin + m, sizeof(C)); x2 = *reinterpret_cast<const C*>(fin + m); }
              else { x2.x = i >= 0 ? fin[m] : T(0); x2.y = i + 1 < win ? fin[m + 1] : T(0); }
            }
            re[brev5(n1)] = (i >= 0 && i < win) ? x2.x * w2.x : T(0);
            im[brev5(n1)] = (i + 1 >= 0 && i + 1 < win) ? x2.y * w2.y : T(0);
          }
        }
        if (SSTTS_STAGE_PREFETCH && !BULK && next < A.n_tiles) {
          // the next tile's span of both parity buffers is pulled into L2 while this tile is transformed
          // (one 128-byte line per thread, no registers held), so the synchronous staging after the
          // gather waits for an L2 hit instead of DRAM.  Placed after the window load: the tile record
          // loaded at the top of the round has arrived by now; tiles with fewer than 6 frames issue
          // only part of the hints
          const int nspan = (nxt.b - nxt.a - 1) * hop + win;
          const int lines = (nspan * (int)sizeof(T) + 127) / 128 + 1;
          const long long base = nxt.poff + nxt.a * hop + lpad;
          const int line = tid < lines ? tid : tid - lines;
          if (tid < 2 * lines) {
            const T* src = ((tid < lines) == (nxt.parity != 0) ? A.pin1 : A.pin0) + base;
            sstts_prefetch_l2(reinterpret_cast<const char*>(src) + 128 * line);
          }
        }
        if constexpr (CXP)
          warp_fft1024_cx<T, false, true, true, G::ZLO, G::ZHI>(re, im, reinterpret_cast<C*>(plane), s_tw, lane,
                                                                  [&]() { stage_row_async(s_mag, mrow, lane, n_bins); });
        else
          warp_fft1024<T, false, true, true, G::ZLO, G::ZHI>(re, im, plane, s_tw, lane);
        if (BULK) {
          // this warp's samples have been consumed by the transform: the last of the tile's warps to get
          // here hands s_yin back to the copy engine for the NEXT tile, whose span then arrives while the
          // CTA is busy with the core, the inverse transform and the gather of this one
          if (lane == 0 && atomicAdd(s_cnt, 1) == FT - 1) {
            *s_cnt = 0;
            if (next < A.n_tiles && is_plain(nxt)) bulk_issue(nxt);
          }
        }
        sstts_cp_async_wait_all();
        __syncwarp();
      }
      double mse_acc = 0.0;
      T ro[32], io[32];
      gl_frame_core<T, FROM_PHASE, WANT_MSE, SSTTS_CORE_BREV_OUT != 0>(re, im, ro, io, srow, prow, s_w2k, lane, mse_acc,
                                                                        bshift, A.phase_seed, A.phase_first + row * n_bins);
      if (WANT_MSE) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mse_acc += __shfl_xor_sync(0xffffffffu, mse_acc, o);
        if (lane == 0) A.mse_frame[row] = mse_acc;
      }
      // inverse transform: with the core's output in bit-reversed slots both passes are FMA-fused DIT
      if constexpr (CXP)
        warp_fft1024_cx<T, true, SSTTS_CORE_BREV_OUT != 0, true, 0, 31>(ro, io, reinterpret_cast<C*>(plane), s_tw, lane, []() {});
      else
        warp_fft1024<T, true, SSTTS_CORE_BREV_OUT != 0, true>(ro, io, plane, s_tw, lane);
      // windowed output frame into the (now dead) plane; slot index = m - mlo, zero outside
      // the window so that the pair store needs no per-element guard
      const T* wout = sh ? s_win2 : s_win;
      const int mlo_w = lp & ~1;
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) {
        const int m = 64 * n1 + 2 * lane;
        const int i = m - lp;
        // slots inside the window for every lane and both frame positions store unconditionally
        const bool inside = 64 * n1 >= mlo + 2 && 64 * n1 + 62 < lpad + win + 1;
        const bool outside = 64 * n1 + 62 < mlo || 64 * n1 >= lpad + win + 2;
        if (!outside && (inside || (m >= mlo_w && m < lp + win + 1))) {
          C vv;
          if (SSTTS_GL_WINDOW_IN_GATHER) { vv.x = ro[n1]; vv.y = io[n1]; (void)i; (void)wout; }
          else {
            const C w2 = *reinterpret_cast<const C*>(wout + i);  // zeros outside the window
            vv.x = ro[n1] * w2.x; vv.y = io[n1] * w2.y;
          }
          SSTTS_CHECK_ALIGNED(plane + (m - mlo_w), sizeof(C));
          *reinterpret_cast<C*>(plane + (m - mlo_w)) = vv;       // m - mlo_w is even
        }
      }
    }
    __syncthreads();
    // Gather overlap-add (ascending frame order, no atomics) straight to the parity buffer:
    // sample s = q * hop + r receives frame q - j at offset r + j * hop, j = jmax .. 0.
    // column form: a thread owns the samples s = q * hop + r of one residue r; every slot element
    // is read exactly once and all indices are compile-time
    {
      const int pe = L.frame_pitch;
      // first sample of frame f inside its output slot: (lpad + shift) & 1
      auto slot_off = [&](int f) -> int { return (lpad + frame_shift(f)) & 1; };
      T* dst = pout_own + span_lo;
      // residues beyond the first NT (hop 275 vs 256 threads: 19 of them) would keep one warp busy
      // for a whole second pass while the others wait at the barrier: they are spread over all
      // threads in row form instead, one output sample per thread
      const int r_cols = hop <= NT ? hop : (hop / NT) * NT;
      const int n_left = hop - r_cols;
      if (n_left > 0) {
        constexpr int NQ = F + MAX_OVERLAP - 1;
        for (int t = tid; t < n_left * NQ; t += NT) {
          const int nl = n_left > 0 ? n_left : 1;      // (compile-time hop = NT: the branch is dead, keep the division defined)
          const int q = t / nl, r = r_cols + t % nl;
          const int s = q * hop + r;
          if (s < span) {
            T acc = T(0);
            for (int j = MAX_OVERLAP - 1; j >= 0; --j) {      // ascending frame order f = q - j
              const int f = q - j, off = r + j * hop;
              if (f >= 0 && f < FT && off < win) {
                if (SSTTS_GL_WINDOW_IN_GATHER) acc = fma(s_planes[f * pe + off + slot_off(f)], s_win[off], acc);
                else acc += s_planes[f * pe + off + slot_off(f)];
              }
            }
            dst[s] = acc;
          }
        }
      }
      for (int r = tid; r < r_cols; r += NT) {
        T acc[F + MAX_OVERLAP - 1];
#pragma unroll
        for (int q = 0; q < F + MAX_OVERLAP - 1; ++q) acc[q] = T(0);
        T wj[MAX_OVERLAP];      // the window at this residue's offsets inside a frame
#pragma unroll
        for (int j = 0; j < MAX_OVERLAP; ++j) wj[j] = (SSTTS_GL_WINDOW_IN_GATHER && r + j * hop < win) ? s_win[r + j * hop] : T(0);
#pragma unroll
        for (int f = 0; f < F; ++f) {
          if (f < FT) {
            const int dl = slot_off(f);
#pragma unroll
            for (int j = 0; j < MAX_OVERLAP; ++j) {
              const int off = r + j * hop;
              if (off < win) {
                if (SSTTS_GL_WINDOW_IN_GATHER) acc[f + j] = fma(s_planes[f * pe + off + dl], wj[j], acc[f + j]);
                else acc[f + j] += s_planes[f * pe + off + dl];
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < F + MAX_OVERLAP - 1; ++q) {
          const int s = q * hop + r;
          if (s < span) dst[s] = acc[q];
        }
      }
    }
    if (!FROM_PHASE && next < A.n_tiles) {
      cur_bulk = BULK && is_plain(nxt);
      if (cur_bulk) bulk_finish(nxt); else stage(nxt);
    }
    __syncthreads();
    tile = next;
    cur = nxt;
    nxt = nxt2;
  }
}

template <typename T> SSTTS_HD size_t gl_step_smem_bytes(int warps, int win, int hop, int span_max, bool bulk = false,
                                                         bool native = false) {
  return GLSmem<T>(warps, win, hop, span_max, bulk, native).total;
}

// Partial sums -> normalised, centre-trimmed float32 waveform (the reference's final istft
// epilogue: divide by the window sum where > tiny, drop n_fft/2 samples on both sides).
template <typename T> struct GLFinalArgs {
  const T* pin0; const T* pin1;
  const long long* frame_off; const long long* pad_off; const long long* sample_off;
  const GLTile* tiles; int n_tiles;
  const T* window;
  float* wav_out;
  int win, hop, n_fft;
};

template <typename T, typename G, int NT>
__global__ void __launch_bounds__(NT) gl_finalize_kernel(const GLFinalArgs<T> A) {
  const G g(A.win, A.hop, A.n_fft);
  const int win = g.win(), hop = g.hop(), lpad = g.lpad(), cpad = g.cpad();
  const T inv_nfft = T(1.0) / T(g.nfft());
  const int tid = threadIdx.x;
  SSTTS_DYN_SMEM(smem);
  T* s_win = reinterpret_cast<T*>(smem);
  T* s_rw = s_win + round_up4(win);
  for (int i = tid; i < win; i += NT) s_win[i] = A.window[i];
  __syncthreads();
  fill_interior_rwss<T>(s_rw, s_win, hop, win, tid, NT, inv_nfft);
  __syncthreads();
  for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
    const GLTile tl = A.tiles[tile];
    const long long f0 = A.frame_off[tl.utt];
    const int n_frames = (int)(A.frame_off[tl.utt + 1] - f0);
    const long long poff = A.pad_off[tl.utt];
    const int L_out = hop * (n_frames - 1);
    const int a = tl.a, b = tl.b;
    const T* pin_own = (tl.parity ? A.pin1 : A.pin0) + poff;
    const T* pin_oth = (tl.parity ? A.pin0 : A.pin1) + poff;
    float* out = A.wav_out + A.sample_off[tl.utt];
    int p_lo = a * hop + lpad;
    if (p_lo < cpad) p_lo = cpad;
    int p_hi = (b < n_frames) ? b * hop + lpad : cpad + L_out;
    if (p_hi > cpad + L_out) p_hi = cpad + L_out;
    const int left_end = (a - 1) * hop + lpad + win;
    for (int p = p_lo + tid; p < p_hi; p += NT) {
      T v = pin_own[p];
      if (a > 0 && p < left_end) v += pin_oth[p];
      out[p - cpad] = (float)normalise_ola<T>(v, p, n_frames, hop, win, lpad, s_win, s_rw, inv_nfft);
    }
  }
}

// =============================================================================================
// STFT feature pipeline
// =============================================================================================
// frames [a, b) of clip; last = tile holds frame T-1.  Like GLTile the record is self-contained (the clip's
// sample range and output rows ride along), so a kernel needs ONE independent 48-byte load per tile -- fetched
// asynchronously into shared memory two rounds ahead -- instead of a tile -> clip -> offsets chain of dependent
// global loads at every tile start (12 % of the feature kernel's stall samples in round 1).
struct alignas(16) FeatTile {
  int clip, a, b, last;
  long long soff;            // first sample of the clip in the packed wav buffer
  long long r0;              // first output row of the clip
  int n_samples, n_frames, n_rows, reserved;
};

template <typename T> struct FeatArgs {
  const float* wav;               // packed clips
  const long long* sample_off;    // [n_clips] first sample of every clip in `wav`
  const long long* sample_len;    // [n_clips] clip lengths N_c
  const long long* frame_off;     // [n_clips + 1] (frame counts T = 1 + N / hop)
  const long long* row_off;       // [n_clips + 1] output row offsets (>= frames: zero pad rows)
  const FeatTile* tiles;
  int n_tiles;
  StftTables<T> tab;
  const int* mel_ptr;             // [n_mels + 1] CSR row pointers
  const int* mel_k0;              // [n_mels]     first FFT bin of each mel filter
  const T* mel_w;                 // [nnz]        filter weights (float64 -> T)
  int n_mels, mel_nnz;
  // the same filterbank padded for the fused dB-feature mode (FeatMode::kDbFeatures): slot j holds
  // the filters m = melp_mbase[j] + lane (32 per slot, top filters first).  Filters start at the even
  // bin k0 & ~1 and are stored as PAIRS of weights: pair i of filter m is the float2
  // melp_w[2 * (melp_woff[j] + 32 * i + lane)] covering bins (k0 & ~1) + 2 i, + 2 i + 1 (zero outside the
  // filter's support), i < melp_len[j]; melp_total counts pairs.
  const float* melp_w;
  int melp_slots, melp_total;
  int melp_len[4], melp_woff[4], melp_mbase[4];
  float2* spec_out;               // (rows, 1025) complex64 STFT, or nullptr
  float* lin_out;                 // (rows, 1025) linear dB (normalised if normalize), or nullptr
  float* mel_out;                 // (rows, n_mels) mel dB (normalised if normalize), or nullptr
  double* melraw_out;             // (rows, n_mels) mel_basis @ |S|**power, or nullptr
  long long* minmax_out;          // [n_clips * 4] order-preserving int64 codes of
                                  //   (min lin dB, max lin dB, min mel dB, max mel dB), or nullptr
  float lin_ref_db, lin_range_db; // normalisation constants: ref, |ref| + |max|
  double mel_ref_db, mel_range_db;
  float mel_power;
  int normalize;
  int win, hop, span_max;
  int n_fft;                      // effective transform size: 2048, 1024 or 512
};

SSTTS_D long long encode_ordered(double v) {
  long long b;
#ifdef SSTTS_CPU_EMU
  std::memcpy(&b, &v, 8);
#else
  b = __double_as_longlong(v);
#endif
  return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}

// dB epilogue constants: 20 log10(x) = kDbPerLog2Mag * log2(x); on |X|^2 it is half of that.
// log2 / sqrt use the SFU approximations (MUFU.LG2 / MUFU.SQRT, ~1e-7 relative): their error is
// ~2e-6 dB, four orders of magnitude inside the 1e-4 normalised tolerance (1e-4 * 135 dB).
#define kDbPerLog2Mag 6.020599913279624f
#define kDbPerLog2Pow 3.010299956639812f

constexpr int FEAT_PLANE_ELEMS = 1056;         // per-warp plane: >= XPLANE_ELEMS and >= NBINS floats
constexpr int FEAT_PLANE_ELEMS_NATIVE = 1120;  // native n_fft 1024 path: >= 2 * HPLANE_ELEMS (and 2 * HMAG floats)
template <typename G> SSTTS_HD constexpr int feat_plane_elems() {
  return G::kNative1024 ? FEAT_PLANE_ELEMS_NATIVE : FEAT_PLANE_ELEMS;
}
constexpr int kNativeTileFrames = 16;   // frames per tile of the native n_fft 1024 path (two per warp and round)
constexpr int HMAG = 544;               // per-half |S| row of the native path: 513 bins + zero slack for padded filters

// MODE = FeatMode::kGeneric: every output is optional and selected at run time.
// MODE = FeatMode::kDbFeatures: the pre-calculation configuration (datasets/lj_speech.py:106-156) --
// linear dB + mel dB (power 1), optionally normalised, n_fft = 2048, nothing else: the dB / normalise
// chain collapses to max -> MUFU.LG2 -> one FFMA -> clamp per bin (the transform's factor 2 and the
// dB scale are folded into the FFMA constants), the mel projection runs in float32 over a padded,
// lane-per-filter table without index arithmetic, and the staging of interior tiles uses 16-byte
// copies.  Results differ from kGeneric by float32 rounding only (~1e-7 of the normalised value).
struct FeatMode { enum { kGeneric = 0, kDbFeatures = 1 }; };

#ifndef SSTTS_FEAT_MINBLOCKS
#define SSTTS_FEAT_MINBLOCKS 2
#endif
template <typename T, typename G, int W, int MODE>
__global__ void __launch_bounds__(W * 32, sizeof(T) == 4 ? SSTTS_FEAT_MINBLOCKS : (W <= 6 ? 2 : 1)) stft_feature_kernel(const FeatArgs<T> A) {
  constexpr bool FAST = MODE == FeatMode::kDbFeatures;
  typedef typename cx_of<T>::type C;
  const G g(A.win, A.hop, A.n_fft);
  const int win = g.win(), hop = g.hop(), lpad = g.lpad(), cpad = g.cpad();
  const int bshift = g.bin_shift(), bmask = (1 << bshift) - 1;
  const int n_bins = (HALF >> bshift) + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NT = W * 32;

  SSTTS_DYN_SMEM(smem);
  C* s_tw = reinterpret_cast<C*>(smem);
  C* s_w2k = s_tw + 1024;
  FeatTile* s_rec = reinterpret_cast<FeatTile*>(s_w2k + 512);      // ring of 3 tile records
  T* s_planes = reinterpret_cast<T*>(s_rec + 3);
  constexpr int PLANE = feat_plane_elems<G>();
  T* s_win = load_window_table<T>(s_planes + W * PLANE, A.tab.window, win, lpad, tid, W * 32);
  float* s_x = reinterpret_cast<float*>(s_planes + W * PLANE + round_up4(win + WIN_TAB_PAD));

  for (int i = tid; i < 1024; i += NT) s_tw[i] = A.tab.tw1024[i];
  for (int i = tid; i < 512; i += NT) s_w2k[i] = A.tab.w2048[i];
  __syncthreads();

  // mel filterbank in shared memory, after the two sample-span buffers: CSR (generic mode) or the
  // padded lane-per-filter table (dB-feature mode)
  const int xbuf_elems = round_up4(A.span_max) + 8;     // + slack for the 16-byte staging shift
  T* s_mel_w = reinterpret_cast<T*>(s_x + 2 * xbuf_elems);
  int* s_mel_ptr = reinterpret_cast<int*>(s_mel_w + round_up4(A.mel_nnz));
  int* s_mel_k0 = s_mel_ptr + round_up4(A.n_mels + 1);
  float* s_melp_w = reinterpret_cast<float*>(s_x + 2 * xbuf_elems);
  if (FAST) {
    s_mel_k0 = reinterpret_cast<int*>(s_melp_w + round_up4(2 * A.melp_total));
    for (int i = tid; i < 2 * A.melp_total; i += NT) s_melp_w[i] = A.melp_w[i];
    for (int i = tid; i < A.n_mels; i += NT) s_mel_k0[i] = A.mel_k0[i] & ~1;   // even start of the pairs
  } else {
    for (int i = tid; i < A.mel_nnz; i += NT) s_mel_w[i] = A.mel_w[i];
    for (int i = tid; i < A.n_mels; i += NT) { s_mel_ptr[i] = A.mel_ptr[i]; s_mel_k0[i] = A.mel_k0[i]; }
    if (tid == 0 && A.n_mels > 0) s_mel_ptr[A.n_mels] = A.mel_ptr[A.n_mels];
  }
  __syncthreads();

  T* plane = s_planes + warp * PLANE;
  const bool want_lin = A.lin_out != nullptr;
  const bool want_lin_db = want_lin || (A.minmax_out != nullptr);
  const bool want_mel = (A.n_mels > 0) && ((A.mel_out != nullptr) || (A.melraw_out != nullptr) ||
                                           (A.minmax_out != nullptr));
  // audio/conversion.py:78  clip(1 + (db - ref) / range, 0, 1) as one fma: db * scale + shift
  const float lin_scale = 1.0f / A.lin_range_db, lin_shift = 1.0f - A.lin_ref_db / A.lin_range_db;
  const float mel_scale = (float)(1.0 / A.mel_range_db), mel_shift = (float)(1.0 - A.mel_ref_db / A.mel_range_db);
  const int pmode = A.mel_power == 1.0f ? 0 : (A.mel_power == 2.0f ? 1 : 2);
  // dB-feature mode works on 2 X (|2 X|^2 = 4 |X|^2, mel of 2 |X|): value = log2(.) * a + b
  const float fl_a = kDbPerLog2Pow * (A.normalize ? lin_scale : 1.0f);
  const float fl_b = (A.normalize ? lin_shift : 0.0f) - 2.0f * fl_a;
  const float fm_a = kDbPerLog2Mag * (A.normalize ? mel_scale : 1.0f);
  const float fm_b = (A.normalize ? mel_shift : 0.0f) - fm_a;
  const float clip_lo = A.normalize ? 0.0f : -3.0e38f, clip_hi = A.normalize ? 1.0f : 3.0e38f;

  // Sample spans are staged with 4-byte LDGSTS (reflect padding resolved per element) into a
  // double buffer: the span of the NEXT tile is in flight while the current tile is transformed.
  // Interior tiles (no reflection, >= 3 samples of the clip on both sides of the span) are copied in
  // 16-byte chunks: the span starts `mis` floats past a 16-byte boundary of the packed wav buffer,
  // so sample s lands at buf[s + mis]; issue_stage returns mis (0 for the per-element path).
  // the two sample buffers are addressed as s_x + k * xbuf_elems (NOT through an array of pointers: indexing one
  // with a run-time value makes the compiler fall back to generic LD for every sample load)
  auto issue_stage = [&](const FeatTile& tl, float* buf) -> int {
    const long long soff = tl.soff;
    const int n_samples = tl.n_samples;
    const int span_lo = tl.a * hop + lpad;
    const int span = (tl.b - tl.a - 1) * hop + win;
    const float* x = A.wav + soff;
    const int q0 = span_lo - cpad;
    if (q0 >= 3 && q0 + span + 3 <= n_samples) {
      const float* g = x + q0;
      const int mis = (int)((reinterpret_cast<uintptr_t>(g) >> 2) & 3);
      const float* gal = g - mis;
      const int nch = (span + mis + 3) >> 2;
      for (int c = tid; c < nch; c += NT) sstts_cp_async16(buf + 4 * c, gal + 4 * c);
      return mis;
    }
    for (int s = tid; s < span; s += NT) {
      int q = q0 + s;
      if (q < 0 || q >= n_samples) q = reflect_index(q, n_samples);
      sstts_cp_async4(buf + s, x + q);
    }
    return 0;
  };
  // tile records travel two rounds ahead of their use: three 16-byte asynchronous copies into slot round % 3
  auto fetch_record = [&](int t, int slot) {
    if (tid < 3 && t < A.n_tiles)
      sstts_cp_async16(reinterpret_cast<char*>(s_rec + slot) + 16 * tid, reinterpret_cast<const char*>(A.tiles + t) + 16 * tid);
  };
  int cur = 0;
  int mis_cur = 0, mis_nxt = 0;
  const int grid = (int)gridDim.x;
  fetch_record(blockIdx.x, 0);
  fetch_record(blockIdx.x + grid, 1);
  sstts_cp_async_wait_all();
  __syncthreads();
  if ((int)blockIdx.x < A.n_tiles) mis_nxt = issue_stage(s_rec[0], s_x);
  sstts_cp_async_commit();

  int round = 0;
  for (int tile = blockIdx.x; tile < A.n_tiles; tile += grid, ++round) {
    sstts_cp_async_wait_all();    // this thread's part of the current span (and of the next record) has landed
    __syncthreads();              // ... everyone's has, and the other buffer / the oldest record slot are free
    const FeatTile tl = s_rec[round % 3];
    const int n_frames = tl.n_frames;
    const long long r0 = tl.r0;
    const int n_rows = tl.n_rows;
    const int a = tl.a, b = tl.b, FT = b - a;
    const float* s_xc = s_x + cur * xbuf_elems;
    mis_cur = mis_nxt;
    fetch_record(tile + 2 * grid, (round + 2) % 3);
    if (tile + grid < A.n_tiles) mis_nxt = issue_stage(s_rec[(round + 1) % 3], s_x + (cur ^ 1) * xbuf_elems);
    sstts_cp_async_commit();
    cur ^= 1;
    // zero rows appended by apply_reduction_padding (datasets/dataset_helper.py:383-393)
    if (tl.last && n_rows > n_frames) {
      const long long zr0 = r0 + n_frames;
      const int nz = n_rows - n_frames;
      if (A.lin_out) for (int i = tid; i < nz * n_bins; i += NT) A.lin_out[zr0 * n_bins + i] = 0.0f;
      if (A.mel_out) for (int i = tid; i < nz * A.n_mels; i += NT) A.mel_out[zr0 * A.n_mels + i] = 0.0f;
      if (A.spec_out) for (int i = tid; i < nz * n_bins; i += NT) A.spec_out[zr0 * n_bins + i] = make_float2(0.0f, 0.0f);
      if (A.melraw_out) for (int i = tid; i < nz * A.n_mels; i += NT) A.melraw_out[zr0 * A.n_mels + i] = 0.0;
    }

    float mn_lin = 3.0e38f, mx_lin = -3.0e38f, mn_mel = 3.0e38f, mx_mel = -3.0e38f;
    if constexpr (G::kNative1024) {
      // n_fft = 1024 natively: two frames per warp, the 16 lanes of a half own one 512-point complex transform
      // (halfwarp_fft512).  Z[k] sits in slot s of lane hl with k = hl + (s & 16) + 32 (s & 15); the conjugate
      // partner Z[512 - k] of a lower-set bin (s < 16) is in the upper set of lane (16 - hl) & 15, slot 31 - s;
      // lane hl == 0 holds both members of its pairs: slots (j, 16 - j) and (16 + j, 31 - j), plus the DC /
      // Nyquist pair in slot 0 and the self-conjugate bin 256 in slot 8.
      // FAST (fused mode): float32 epilogue like the n_fft 2048 dB-feature mode -- any of linear dB, mel dB and
      // the per-clip extrema (the statistics pass, datasets/statistics.py:54-66: the extrema of the dB values
      // are the dB values of the extrema, so only min / max of |X|^2 and of the mel sums are tracked per bin).
      const int half = lane >> 4, hl = lane & 15;
      float* s_mag2 = reinterpret_cast<float*>(plane);          // after the transposes: |S| rows of both frames
      const int partner = (lane & 16) | ((16 - hl) & 15);
      float pmin = 3.0e38f, pmax = 0.0f, mmin = 3.0e38f, mmax = 0.0f;   // FAST: extrema of 4 |X|^2 and of 2 mel
      for (int jp = warp; 2 * jp < FT; jp += W) {
        const bool live = 2 * jp + half < FT;
        const int jr = live ? 2 * jp + half : 2 * jp;           // an odd tile's last half redoes its sibling, stores nothing
        const long long row = r0 + a + jr;
        T re[32], im[32];
        const float* fin = s_xc + mis_cur + jr * hop - lpad;
        // (8-byte sample-pair loads for frames at even offsets were measured here: 2-4 % SLOWER in both precisions,
        // profiles/experiments/r2_ab_paired_loads.txt -- scalar loads)
        load_windowed_frame<T, float, 32, false>(re, im, fin, s_win, lpad, win, hl);
        halfwarp_fft512<T>(re, im, plane, half, s_tw, hl);
        float* s_mag = s_mag2 + half * HMAG;
        float* lin_row = (FAST && A.lin_out) ? A.lin_out + row * n_bins : nullptr;
        // one bin, given 2 X[ko]; librosa stores X as complex64, everything downstream is float32
        auto emit = [&](int ko, T xr2, T xi2) {
          if (FAST) {
            const float fr = (float)xr2, fi = (float)xi2;       // an exact power-of-two multiple of the complex64 value
            const float p4 = fmaf(fr, fr, fi * fi);             // 4 |X|^2
            s_mag[ko] = sstts_sqrt_approx(p4);                  // 2 |X| for the mel projection
            pmin = fminf(pmin, p4);
            pmax = fmaxf(pmax, p4);
            if (lin_row) lin_row[ko] = fminf(fmaxf(fmaf(sstts_log2_ftz(fmaxf(4e-10f, p4)), fl_a, fl_b), clip_lo), clip_hi);
          } else {
            const float fr = (float)(T(0.5) * xr2), fi = (float)(T(0.5) * xi2);
            if (A.spec_out) A.spec_out[row * n_bins + ko] = make_float2(fr, fi);
            const float p = fmaf(fr, fr, fi * fi);
            if (want_mel) s_mag[ko] = sstts_sqrt_approx(p);
            if (want_lin_db) {
              const float d = kDbPerLog2Pow * sstts_log2_approx(fmaxf(1e-10f, p));
              mn_lin = fminf(mn_lin, d);
              mx_lin = fmaxf(mx_lin, d);
              if (want_lin) A.lin_out[row * n_bins + ko] = A.normalize ? fminf(fmaxf(fmaf(d, lin_scale, lin_shift), 0.0f), 1.0f) : d;
            }
          }
        };
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int sa = j < 8 ? j : 8 + j;                      // lane 0's own pairs
          const int sb = j == 0 ? 0 : (j < 8 ? 16 - j : 39 - j);
          const T zr = (j >= 8 && hl == 0) ? re[sa] : re[j];
          const T zi = (j >= 8 && hl == 0) ? im[sa] : im[j];
          const T pr = __shfl_sync(0xffffffffu, hl == 0 ? re[sb] : re[31 - j], partner);
          const T pi = __shfl_sync(0xffffffffu, hl == 0 ? im[sb] : im[31 - j], partner);
          const int k = hl != 0 ? hl + 32 * j : (j < 8 ? 32 * j : 32 * j - 240);
          const int kn = 512 - k;
          const C w = s_w2k[k];                                  // exp(-2 pi i k / 1024)
          const T er = zr + pr, ei = zi - pi;                    // Zk + conj Zn
          const T dr = zr - pr, di = zi + pi;                    // Zk - conj Zn
          const T wor = w.x * di + w.y * dr;
          const T woi = w.y * di - w.x * dr;
          if (live) {
            emit(k, er + wor, ei + woi);                         // 2 X[k]
            emit(kn, er - wor, woi - ei);                        // 2 X[512 - k]
          }
        }
        if (hl == 0 && live) emit(256, T(2) * re[8], T(-2) * im[8]);     // self-conjugate bin: X = conj(Z)
        if (FAST) {
          s_mag[513 + hl] = 0.0f;                                // zero slack read (times zero weights) by padded filters
          __syncwarp();
          for (int f = 0; f < 2 && 2 * jp + f < FT; ++f) {
            const float* smf = s_mag2 + f * HMAG;
            float* mel_row = A.mel_out ? A.mel_out + (r0 + a + 2 * jp + f) * A.n_mels : nullptr;
            for (int j = 0; j < A.melp_slots; ++j) {
              const int m = A.melp_mbase[j] + lane;
              const bool valid = m >= 0 && m < A.n_mels;
              const float2* wp = reinterpret_cast<const float2*>(s_melp_w) + A.melp_woff[j] + lane;
              const float2* mg = reinterpret_cast<const float2*>(smf + (valid ? s_mel_k0[m] : 0));
              SSTTS_CHECK_ALIGNED(mg, sizeof(float2));
              SSTTS_CHECK_ALIGNED(wp, sizeof(float2));
              const int len = A.melp_len[j];
              float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 4
              for (int i = 0; i < len; ++i) {
                const float2 w = wp[32 * i], v = mg[i];
                acc0 = fmaf(w.x, v.x, acc0);
                acc1 = fmaf(w.y, v.y, acc1);
              }
              const float m2 = acc0 + acc1;                      // 2 x mel
              if (valid) {
                mmin = fminf(mmin, m2);
                mmax = fmaxf(mmax, m2);
                if (mel_row) mel_row[m] = fminf(fmaxf(fmaf(sstts_log2_ftz(fmaxf(2e-5f, m2)), fm_a, fm_b), clip_lo), clip_hi);
              }
            }
          }
          __syncwarp();
          continue;
        }
        __syncwarp();
        if (want_mel) {
          for (int f = 0; f < 2 && 2 * jp + f < FT; ++f) {
            const long long frow = r0 + a + 2 * jp + f;
            const float* smf = s_mag2 + f * HMAG;
            for (int m = lane; m < A.n_mels; m += 32) {
              const int p0 = s_mel_ptr[m], p1 = s_mel_ptr[m + 1];
              const float* mg = smf + s_mel_k0[m] - p0;
              T acc = T(0);
              if (pmode == 0) {
                for (int i = p0; i < p1; ++i) acc += s_mel_w[i] * (T)mg[i];
              } else if (pmode == 1) {
                for (int i = p0; i < p1; ++i) acc += s_mel_w[i] * (T)(mg[i] * mg[i]);
              } else {
                for (int i = p0; i < p1; ++i) acc += s_mel_w[i] * (T)powf(mg[i], A.mel_power);
              }
              if (A.melraw_out) A.melraw_out[frow * A.n_mels + m] = (double)acc;
              const float db = kDbPerLog2Mag * sstts_log2_approx(fmaxf(1e-5f, fabsf((float)acc)));
              mn_mel = fminf(mn_mel, db);
              mx_mel = fmaxf(mx_mel, db);
              if (A.mel_out)
                A.mel_out[frow * A.n_mels + m] = A.normalize ? fminf(fmaxf(fmaf(db, mel_scale, mel_shift), 0.0f), 1.0f) : db;
            }
          }
        }
        __syncwarp();
      }
      if (FAST) {
        // 20 log10(max(1e-5, |X|)) from 4 |X|^2, 20 log10(max(1e-5, mel)) from 2 mel (audio/conversion.py:29)
        mn_lin = kDbPerLog2Pow * (sstts_log2_ftz(fmaxf(4e-10f, pmin)) - 2.0f);
        mx_lin = kDbPerLog2Pow * (sstts_log2_ftz(fmaxf(4e-10f, pmax)) - 2.0f);
        mn_mel = kDbPerLog2Mag * (sstts_log2_ftz(fmaxf(2e-5f, mmin)) - 1.0f);
        mx_mel = kDbPerLog2Mag * (sstts_log2_ftz(fmaxf(2e-5f, mmax)) - 1.0f);
      }
    }
    if constexpr (!G::kNative1024)
    for (int jr = warp; jr < FT; jr += W) {
      const long long row = r0 + a + jr;
      T re[32], im[32];
      const float* fin = s_xc + mis_cur + jr * hop - lpad;
      // float32: frames that start at an even element of the staged span load their sample pairs with one 8-byte
      // load (fused mode 0.409 -> 0.397 ms); the float64 instances got 1-3 % slower with it and keep scalar loads
      if (sizeof(T) == 4 && ((mis_cur + jr * hop - lpad) & 1) == 0) load_windowed_frame<T, float, 64, true>(re, im, fin, s_win, lpad, win, lane);
      else load_windowed_frame<T, float, 64, false>(re, im, fin, s_win, lpad, win, lane);
      warp_fft1024<T, false, true, true, G::ZLO, G::ZHI>(re, im, plane, s_tw, lane);
      float* s_mag = reinterpret_cast<float*>(plane);  // transpose plane is dead: |S| of this frame
      float* lin_row = FAST ? A.lin_out + row * NBINS : nullptr;
      const int partner = (32 - lane) & 31;
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const int sl_mine = k2;
        const int sl_part = 31 - k2;
        const int sl_alt = (32 - k2) & 31;
        const int k = lane + 32 * k2;
        const int kn = HALF - k;
        const C w = s_w2k[k];
        const T zr = re[sl_mine], zi = im[sl_mine];
        T pr = __shfl_sync(0xffffffffu, re[sl_part], partner);
        T pi = __shfl_sync(0xffffffffu, im[sl_part], partner);
        if (lane == 0) { pr = re[sl_alt]; pi = im[sl_alt]; }
        const T er = zr + pr, ei = zi - pi;
        const T dr = zr - pr, di = zi + pi;
        const T wor = w.x * di + w.y * dr;
        const T woi = w.y * di - w.x * dr;
        if (FAST) {
          // 2 X[k], 2 X[N-k] rounded to float32 (an exact power-of-two multiple of the complex64
          // value librosa stores); |2 X|^2 -> 2 |X| for the mel projection and the fused dB value
          const float fkr = (float)(er + wor), fki = (float)(ei + woi);
          const float fnr = (float)(er - wor), fni = (float)(woi - ei);
          const float pk = fmaf(fkr, fkr, fki * fki), pn = fmaf(fnr, fnr, fni * fni);
          s_mag[k] = sstts_sqrt_approx(pk);
          s_mag[kn] = sstts_sqrt_approx(pn);
          const float lk = sstts_log2_ftz(fmaxf(4e-10f, pk)), ln = sstts_log2_ftz(fmaxf(4e-10f, pn));
          lin_row[k] = fminf(fmaxf(fmaf(lk, fl_a, fl_b), clip_lo), clip_hi);
          lin_row[kn] = fminf(fmaxf(fmaf(ln, fl_a, fl_b), clip_lo), clip_hi);
          continue;
        }
        const T xkr = T(0.5) * (er + wor), xki = T(0.5) * (ei + woi);   // X[k]
        const T xnr = T(0.5) * (er - wor), xni = T(0.5) * (woi - ei);   // X[N-k]
        if ((k & bmask) == 0) {   // kn = 1024 - k shares k's residue
          // librosa stores the float64 transform as complex64; everything downstream is float32
          const float fkr = (float)xkr, fki = (float)xki, fnr = (float)xnr, fni = (float)xni;
          const int ko = k >> bshift, no = kn >> bshift;
          if (A.spec_out) {
            A.spec_out[row * n_bins + ko] = make_float2(fkr, fki);
            A.spec_out[row * n_bins + no] = make_float2(fnr, fni);
          }
          const float pk = fmaf(fkr, fkr, fki * fki), pn = fmaf(fnr, fnr, fni * fni);   // |X|^2
          if (want_mel) { s_mag[ko] = sstts_sqrt_approx(pk); s_mag[no] = sstts_sqrt_approx(pn); }
          if (want_lin_db) {
            // 20 log10(max(1e-5, |X|)) = 10 log10(2) log2(max(1e-10, |X|^2))   (conversion.py:29)
            const float dk = kDbPerLog2Pow * sstts_log2_approx(fmaxf(1e-10f, pk));
            const float dn = kDbPerLog2Pow * sstts_log2_approx(fmaxf(1e-10f, pn));
            mn_lin = fminf(mn_lin, fminf(dk, dn));
            mx_lin = fmaxf(mx_lin, fmaxf(dk, dn));
            if (want_lin) {
              A.lin_out[row * n_bins + ko] = A.normalize ? fminf(fmaxf(fmaf(dk, lin_scale, lin_shift), 0.0f), 1.0f) : dk;
              A.lin_out[row * n_bins + no] = A.normalize ? fminf(fmaxf(fmaf(dn, lin_scale, lin_shift), 0.0f), 1.0f) : dn;
            }
          }
        }
      }
      if (FAST) {
        if (lane == 0) {   // k = 512: X = conj(Z); keep the 2 X convention of the pair loop
          const float fr = 2.0f * (float)re[16], fi = 2.0f * (float)im[16];
          const float ph = fmaf(fr, fr, fi * fi);
          s_mag[HALF / 2] = sstts_sqrt_approx(ph);
          lin_row[HALF / 2] = fminf(fmaxf(fmaf(sstts_log2_ftz(fmaxf(4e-10f, ph)), fl_a, fl_b), clip_lo), clip_hi);
        } else {
          s_mag[NBINS - 1 + lane] = 0.0f;   // zero slack read (times zero weights) by padded filters
        }
        __syncwarp();
        // mel_basis @ |S| in float32: lane = filter, padded weights [i][lane] (conflict-free), |S|
        // taken from this warp's plane; all terms are >= 0, so float32 accumulation is accurate to
        // ~n eps (1e-6 relative, 1e-5 dB)
        float* mel_row = A.mel_out + row * A.n_mels;
        for (int j = 0; j < A.melp_slots; ++j) {
          const int m = A.melp_mbase[j] + lane;
          const bool valid = m >= 0 && m < A.n_mels;
          // one 8-byte load of two weights and one of two magnitudes per step (the plane is 8-byte
          // aligned and the pairs start at an even bin)
          const float2* wp = reinterpret_cast<const float2*>(s_melp_w) + A.melp_woff[j] + lane;
          const float2* mg = reinterpret_cast<const float2*>(s_mag + (valid ? s_mel_k0[m] : 0));
          SSTTS_CHECK_ALIGNED(mg, sizeof(float2));
          SSTTS_CHECK_ALIGNED(wp, sizeof(float2));
          const int len = A.melp_len[j];
          float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 4
          for (int i = 0; i < len; ++i) {
            const float2 w = wp[32 * i], v = mg[i];
            acc0 = fmaf(w.x, v.x, acc0);
            acc1 = fmaf(w.y, v.y, acc1);
          }
          const float lm = sstts_log2_ftz(fmaxf(2e-5f, acc0 + acc1));
          if (valid) mel_row[m] = fminf(fmaxf(fmaf(lm, fm_a, fm_b), clip_lo), clip_hi);
        }
        __syncwarp();
        continue;
      }
      if (lane == 0) {   // k = 512: X = conj(Z)
        const int sl = 16;
        const int kh = (HALF / 2) >> bshift;
        const float fr = (float)re[sl], fi = -(float)im[sl];
        if (A.spec_out) A.spec_out[row * n_bins + kh] = make_float2(fr, fi);
        const float ph = fmaf(fr, fr, fi * fi);
        if (want_mel) s_mag[kh] = sstts_sqrt_approx(ph);
        if (want_lin_db) {
          const float dh = kDbPerLog2Pow * sstts_log2_approx(fmaxf(1e-10f, ph));
          mn_lin = fminf(mn_lin, dh);
          mx_lin = fmaxf(mx_lin, dh);
          if (want_lin) A.lin_out[row * n_bins + kh] = A.normalize ? fminf(fmaxf(fmaf(dh, lin_scale, lin_shift), 0.0f), 1.0f) : dh;
        }
      }
      __syncwarp();
      if (want_mel) {
        // mel_basis @ |S| ** power: one filter per lane and pass; filters are contiguous runs of
        // FFT bins (CSR staged in shared memory); accumulation in T (float64 in the exact mode,
        // like the reference's float64 np.dot, audio/features.py:84)
        for (int m = lane; m < A.n_mels; m += 32) {
          const int p0 = s_mel_ptr[m], p1 = s_mel_ptr[m + 1];
          const float* mg = s_mag + s_mel_k0[m] - p0;
          T acc = T(0);
          if (pmode == 0) {
            for (int i = p0; i < p1; ++i) acc += s_mel_w[i] * (T)mg[i];
          } else if (pmode == 1) {
            for (int i = p0; i < p1; ++i) acc += s_mel_w[i] * (T)(mg[i] * mg[i]);
          } else {
            for (int i = p0; i < p1; ++i) acc += s_mel_w[i] * (T)powf(mg[i], A.mel_power);
          }
          if (A.melraw_out) A.melraw_out[row * A.n_mels + m] = (double)acc;
          const float db = kDbPerLog2Mag * sstts_log2_approx(fmaxf(1e-5f, fabsf((float)acc)));
          mn_mel = fminf(mn_mel, db);
          mx_mel = fmaxf(mx_mel, db);
          if (A.mel_out)
            A.mel_out[row * A.n_mels + m] = A.normalize ? fminf(fmaxf(fmaf(db, mel_scale, mel_shift), 0.0f), 1.0f) : db;
        }
      }
      __syncwarp();
    }
    if (A.minmax_out) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn_lin = fminf(mn_lin, __shfl_xor_sync(0xffffffffu, mn_lin, o));
        mx_lin = fmaxf(mx_lin, __shfl_xor_sync(0xffffffffu, mx_lin, o));
        mn_mel = fminf(mn_mel, __shfl_xor_sync(0xffffffffu, mn_mel, o));
        mx_mel = fmaxf(mx_mel, __shfl_xor_sync(0xffffffffu, mx_mel, o));
      }
      if (lane == 0 && (G::kNative1024 ? 2 * warp < FT : warp < FT)) {
        long long* mm = A.minmax_out + 4LL * tl.clip;
        atomicMin(mm + 0, encode_ordered((double)mn_lin));
        atomicMax(mm + 1, encode_ordered((double)mx_lin));
        atomicMin(mm + 2, encode_ordered((double)mn_mel));
        atomicMax(mm + 3, encode_ordered((double)mx_mel));
      }
    }
    // no barrier here: the one at the top of the next iteration orders the buffer hand-over
  }
}

// melp_total > 0 (weight pairs) selects the dB-feature layout (padded float2 table + k0) instead of the CSR one.
template <typename T>
SSTTS_HD size_t stft_feature_smem_bytes(int warps, int win, int span_max, int n_mels, int mel_nnz,
                                        int melp_total = 0, int plane_elems = FEAT_PLANE_ELEMS) {
  const size_t mel = melp_total > 0
      ? sizeof(float) * (size_t)round_up4(2 * melp_total) + sizeof(int) * (size_t)round_up4(n_mels)
      : sizeof(T) * (size_t)round_up4(mel_nnz) + sizeof(int) * (size_t)(round_up4(n_mels + 1) + round_up4(n_mels));
  return sizeof(typename cx_of<T>::type) * (size_t)(1024 + 512) + 3 * sizeof(FeatTile) +
         sizeof(T) * (size_t)(warps * plane_elems + round_up4(win + WIN_TAB_PAD)) +
         sizeof(float) * 2 * (size_t)(round_up4(span_max) + 8) + mel;
}

// =============================================================================================
// Silence trimming -- librosa.effects.trim(y, top_db, ref=np.max, frame_length, hop_length) as
// called by datasets/lj_speech.py:119 (defaults 60 / 2048 / 512): frame mean-square energy over the
// centre-reflect-padded clip, frames within top_db of the loudest are "non-silent", the clip is cut
// to [first * hop, min(N, (last + 1) * hop)).  One CTA per clip; sums in float64.
// =============================================================================================
constexpr int TRIM_CACHE_FRAMES = 2048;   // per-frame energies kept in shared memory (recomputed beyond)

template <int NT>
__global__ void __launch_bounds__(NT) trim_bounds_kernel(const float* __restrict__ wav,
                                                         const long long* __restrict__ clip_start,
                                                         const long long* __restrict__ clip_len, int n_clips,
                                                         int frame_length, int hop_length, double top_db,
                                                         long long* __restrict__ bounds) {
  __shared__ double s_ms[TRIM_CACHE_FRAMES];
  __shared__ double s_red[NT / 32];
  __shared__ int s_first, s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  for (int clip = blockIdx.x; clip < n_clips; clip += gridDim.x) {
    const float* x = wav + clip_start[clip];
    const int N = (int)clip_len[clip];
    const int n_frames = 1 + N / hop_length;
    const int half = frame_length / 2;
    auto frame_ms = [&](int f) -> double {   // warp-collective: mean square of frame f
      double acc = 0.0;
      const int base = f * hop_length - half;
      for (int i = lane; i < frame_length; i += 32) {
        int q = base + i;
        if (q < 0 || q >= N) q = reflect_index(q, N);
        const double v = (double)x[q];
        acc += v * v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      return acc / (double)frame_length;
    };
    // pass 1: energies (cached) and their maximum
    double mx = 0.0;
    for (int f = warp; f < n_frames; f += NW) {
      const double ms = frame_ms(f);
      if (lane == 0 && f < TRIM_CACHE_FRAMES) s_ms[f] = ms;
      mx = fmax(mx, ms);
    }
    if (lane == 0) s_red[warp] = mx;
    if (tid == 0) { s_first = 0x7fffffff; s_last = -1; }
    __syncthreads();
    mx = 0.0;
    for (int w = 0; w < NW; ++w) mx = fmax(mx, s_red[w]);
    // power_to_db(mse, ref=np.max, amin=1e-10, top_db=None) > -top_db
    const double ref_db = 10.0 * log10(fmax(1e-10, mx));
    for (int f = warp; f < n_frames; f += NW) {
      const double ms = f < TRIM_CACHE_FRAMES ? s_ms[f] : frame_ms(f);
      const bool loud = 10.0 * log10(fmax(1e-10, ms)) - ref_db > -top_db;
      if (lane == 0 && loud) { atomicMin(&s_first, f); atomicMax(&s_last, f); }
    }
    __syncthreads();
    if (tid == 0) {
      long long start = 0, end = 0;
      if (s_last >= 0) {
        start = (long long)s_first * hop_length;
        end = (long long)(s_last + 1) * hop_length;
        if (end > N) end = N;
      }
      bounds[2 * clip] = start;
      bounds[2 * clip + 1] = end;
    }
    __syncthreads();
  }
}

}  // namespace sstts
