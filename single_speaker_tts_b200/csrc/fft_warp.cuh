// Warp-resident complex FFT of 1024 points (= real FFT of n_fft = 2048) for sm_100a.
//
// Layout: one warp owns one transform.  Lane l holds 32 complex values in registers,
// element index = 32 * r + l (r = register slot, l = lane).  The transform is the classic
// 32 x 32 Cooley-Tukey split:
//     pass 1  32-point FFT over the register index, entirely in registers
//     twiddle W_1024^(l * k1)   (table in shared memory, conflict-free [k1][l] layout)
//     32 x 32 transpose through a padded per-warp shared-memory plane (the only exchange)
//     pass 2  32-point FFT over the register index again
// Input and output use the same "element = 32 * slot + lane" indexing.  A 32-point pass is
// either DIF (natural slot order in, bit-reversed out) or DIT (bit-reversed in, natural out).
// Slot order is free wherever data comes from or goes to memory, and wherever the producer is
// fully unrolled code (slot indices are then compile-time, so writing to slot brev5(k) instead of k
// is only a renaming of registers): the kernels therefore use DIT (FMA-fused butterflies,
// zero-padding pruned for free) for all four passes of a Griffin-Lim iteration -- the window load
// and the conjugate-pair core hand their results over in bit-reversed slots.  The DIF stage is kept
// for completeness and for the emulator's transform tests.  No register permutation is ever executed.
// No 1/N scaling is applied here -- callers fold it into the synthesis window.
//
// The 32-point register FFT is radix-2 with the trivial twiddles (1, -i, (1-i)/sqrt2, ...)
// special-cased at compile time.  Twiddles are literals rounded from extended precision, so
// the float32 transform error stays at the ~1e-7 level the Griffin-Lim tolerance budget
//  assumes; --use_fast_math is never used.
#pragma once
#include "simt_compat.h"

// 1: apply the inter-pass twiddles after the transpose, fused into the first butterfly stage of pass 2
// (-32 instructions per transform).  Measured SLOWER on B200 (Griffin-Lim step 0.572 -> 0.579 ms, float32
// features 0.43 -> 0.44 ms: the twiddle loads land behind the transpose loads on the critical path), so
// the default keeps the separate twiddle loop before the transpose.
#ifndef SSTTS_FUSE_TWIDDLE
#define SSTTS_FUSE_TWIDDLE 0
#endif

namespace sstts {

template <typename T> struct cx_of;
template <> struct cx_of<float> { typedef float2 type; };
template <> struct cx_of<double> { typedef double2 type; };

SSTTS_HD constexpr int brev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// cos(2 pi k / 32), k = 0..8, as literals (the rest of the circle by symmetry).
template <typename T> SSTTS_HD constexpr T cos32_q(int k) {
  return k == 0 ? T(1.0)
       : k == 1 ? T(0.9807852804032304491261822)
       : k == 2 ? T(0.9238795325112867561281832)
       : k == 3 ? T(0.8314696123025452370787884)
       : k == 4 ? T(0.7071067811865475244008444)
       : k == 5 ? T(0.5555702330196022247428308)
       : k == 6 ? T(0.38268343236508977172846)
       : k == 7 ? T(0.1950903220161282678482849)
       : T(0.0);
}
// cos / sin of 2 pi k / 32 for k = 0..15.
template <typename T> SSTTS_HD constexpr T cos32(int k) { return k <= 8 ? cos32_q<T>(k) : -cos32_q<T>(16 - k); }
template <typename T> SSTTS_HD constexpr T sin32(int k) { return k <= 8 ? cos32_q<T>(8 - k) : cos32_q<T>(k - 8); }

// (dr + i di) * W, W = exp(-/+ 2 pi i idx / 32) (forward: minus, INV: plus), idx in [0, 16).
template <typename T, bool INV>
SSTTS_HD void mul_w32(T dr, T di, int idx, T& outr, T& outi) {
  if (idx == 0) {
    outr = dr; outi = di;
  } else if (idx == 8) {
    if (!INV) { outr = di; outi = -dr; } else { outr = -di; outi = dr; }
  } else if (idx == 4) {
    const T h = cos32_q<T>(4);
    if (!INV) { outr = (dr + di) * h; outi = (di - dr) * h; }
    else      { outr = (dr - di) * h; outi = (dr + di) * h; }
  } else if (idx == 12) {
    const T h = cos32_q<T>(4);
    if (!INV) { outr = (di - dr) * h; outi = -(dr + di) * h; }
    else      { outr = -(dr + di) * h; outi = (dr - di) * h; }
  } else {
    const T c = cos32<T>(idx), s = sin32<T>(idx);
    if (!INV) { outr = dr * c + di * s; outi = di * c - dr * s; }
    else      { outr = dr * c - di * s; outi = di * c + dr * s; }
  }
}

// One decimation-in-frequency stage: (a, b) -> (a + b, (a - b) W).
template <typename T, bool INV, int LEN>
SSTTS_HD void dif_stage(T (&re)[32], T (&im)[32]) {
  constexpr int HALF = LEN / 2;
  constexpr int TSTEP = 32 / LEN;
#pragma unroll
  for (int g = 0; g < 32; g += LEN) {
#pragma unroll
    for (int j = 0; j < HALF; ++j) {
      const int i0 = g + j, i1 = g + j + HALF;
      const T ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
      re[i0] = ar + br;
      im[i0] = ai + bi;
      mul_w32<T, INV>(ar - br, ai - bi, j * TSTEP, re[i1], im[i1]);
    }
  }
}

// FMA-fused decimation-in-time butterfly (a, b) -> (a + b W, a - b W): the sum is built with
// chained FMAs and the difference as 2 a - (a + b W), i.e. 6 instructions for a general twiddle
// instead of 8 (complex multiply + two complex adds).
template <typename T, bool INV>
SSTTS_HD void dit_butterfly(T& ar, T& ai, T& br, T& bi, int idx) {
  T sr, si;   // a + b W
  if (idx == 0) {
    sr = ar + br; si = ai + bi;
    br = ar - br; bi = ai - bi;
    ar = sr; ai = si;
    return;
  }
  if (idx == 8) {   // W = -/+ i
    if (!INV) { sr = ar + bi; si = ai - br; const T dr = ar - bi, di = ai + br; br = dr; bi = di; }
    else      { sr = ar - bi; si = ai + br; const T dr = ar + bi, di = ai - br; br = dr; bi = di; }
    ar = sr; ai = si;
    return;
  }
  const T c = cos32<T>(idx), s = INV ? -sin32<T>(idx) : sin32<T>(idx);
  // b W = (br c + bi s, bi c - br s) with W = c - i s
  sr = fma(br, c, fma(bi, s, ar));
  si = fma(bi, c, fma(-br, s, ai));
  br = fma(T(2), ar, -sr);
  bi = fma(T(2), ai, -si);
  ar = sr; ai = si;
}

// One decimation-in-time stage over all 16 butterflies of span LEN.
template <typename T, bool INV, int LEN>
SSTTS_HD void dit_stage(T (&re)[32], T (&im)[32]) {
  constexpr int HALF = LEN / 2;
  constexpr int TSTEP = 32 / LEN;
#pragma unroll
  for (int g = 0; g < 32; g += LEN) {
#pragma unroll
    for (int j = 0; j < HALF; ++j) dit_butterfly<T, INV>(re[g + j], im[g + j], re[g + j + HALF], im[g + j + HALF], j * TSTEP);
  }
}

// First DIT stage (span 2, W = 1) when the input elements outside [ZLO, ZHI] are known to be zero
// (zero-padded window): slot 2g holds element e = brev5(2g) < 16 and slot 2g + 1 element e + 16,
// so most butterflies degenerate to copies / negations and cost nothing.
template <typename T, int ZLO, int ZHI>
SSTTS_HD void dit_stage2_pruned(T (&re)[32], T (&im)[32]) {
#pragma unroll
  for (int g = 0; g < 32; g += 2) {
    const int e = brev5(g);
    const bool a_zero = e < ZLO || e > ZHI, b_zero = e + 16 < ZLO || e + 16 > ZHI;
    if (a_zero && b_zero) {
      re[g] = T(0); im[g] = T(0); re[g + 1] = T(0); im[g + 1] = T(0);
    } else if (a_zero) {
      re[g] = re[g + 1]; im[g] = im[g + 1]; re[g + 1] = -re[g + 1]; im[g + 1] = -im[g + 1];
    } else if (b_zero) {
      re[g + 1] = re[g]; im[g + 1] = im[g];
    } else {
      const T ar = re[g], ai = im[g], br = re[g + 1], bi = im[g + 1];
      re[g] = ar + br; im[g] = ai + bi; re[g + 1] = ar - br; im[g + 1] = ai - bi;
    }
  }
}

// First DIT stage (span 2, W = 1) fused with the inter-pass twiddles of the 1024-point transform:
// slot 2g holds element e = brev5(2g) < 16 and slot 2g + 1 element e + 16; both are first multiplied by
// their twiddle t_e = tw[32 e + lane] (conjugated for the inverse), then combined:
//   A = x_e t_e,  out0 = A + x_(e+16) t_(e+16),  out1 = 2 A - out0
// 10 instructions per butterfly instead of 8 (two twiddle multiplies) + 4 (butterfly).
template <typename T, bool INV>
SSTTS_D void dit_stage2_twiddled(T (&re)[32], T (&im)[32], const typename cx_of<T>::type* tw, int lane) {
  typedef typename cx_of<T>::type C;
#pragma unroll
  for (int g = 0; g < 32; g += 2) {
    const int e = brev5(g);
    T ar = re[g], ai = im[g];
    if (e != 0) {   // t_0 = 1
      const C ta = tw[e * 32 + lane];
      const T sy = INV ? -ta.y : ta.y;
      const T tr = ar * ta.x - ai * sy, ti = ar * sy + ai * ta.x;
      ar = tr; ai = ti;
    }
    const C tb = tw[(e + 16) * 32 + lane];
    const T sb = INV ? -tb.y : tb.y;
    const T br = re[g + 1], bi = im[g + 1];
    const T sr = fma(br, tb.x, fma(-bi, sb, ar));
    const T si = fma(br, sb, fma(bi, tb.x, ai));
    re[g] = sr; im[g] = si;
    re[g + 1] = fma(T(2), ar, -sr);
    im[g + 1] = fma(T(2), ai, -si);
  }
}

// In-register 32-point FFT.  DIT = false: element k in slot k -> result k in slot brev5(k).
//                            DIT = true : element k in slot brev5(k) -> result k in slot k;
//                            input elements outside [ZLO, ZHI] must be zero (default: none are).
// SKIP2: the span-2 stage has already been applied (dit_stage2_twiddled).
template <typename T, bool INV, bool DIT, int ZLO = 0, int ZHI = 31, bool SKIP2 = false>
SSTTS_HD void fft32(T (&re)[32], T (&im)[32]) {
  if (!DIT) {
    dif_stage<T, INV, 32>(re, im);
    dif_stage<T, INV, 16>(re, im);
    dif_stage<T, INV, 8>(re, im);
    dif_stage<T, INV, 4>(re, im);
    dif_stage<T, INV, 2>(re, im);
  } else {
    if (SKIP2) {}
    else if (ZLO > 0 || ZHI < 31) dit_stage2_pruned<T, ZLO, ZHI>(re, im);
    else dit_stage<T, INV, 2>(re, im);
    dit_stage<T, INV, 4>(re, im);
    dit_stage<T, INV, 8>(re, im);
    dit_stage<T, INV, 16>(re, im);
    dit_stage<T, INV, 32>(re, im);
  }
}

SSTTS_HD constexpr int brev4(int i) { return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3); }

// In-register 16-point DIT FFT over the slots BASE .. BASE + 15: element n in slot BASE + brev4(n) ->
// result k in slot BASE + k.  W_16^j = W_32^(2 j), so the butterflies are the ones of the 32-point pass.
template <typename T, bool INV, int BASE>
SSTTS_HD void fft16_dit(T (&re)[32], T (&im)[32]) {
#pragma unroll
  for (int g = 0; g < 16; g += 2) dit_butterfly<T, INV>(re[BASE + g], im[BASE + g], re[BASE + g + 1], im[BASE + g + 1], 0);
#pragma unroll
  for (int g = 0; g < 16; g += 4) {
#pragma unroll
    for (int j = 0; j < 2; ++j) dit_butterfly<T, INV>(re[BASE + g + j], im[BASE + g + j], re[BASE + g + j + 2], im[BASE + g + j + 2], j * 8);
  }
#pragma unroll
  for (int g = 0; g < 16; g += 8) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dit_butterfly<T, INV>(re[BASE + g + j], im[BASE + g + j], re[BASE + g + j + 4], im[BASE + g + j + 4], j * 4);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) dit_butterfly<T, INV>(re[BASE + j], im[BASE + j], re[BASE + j + 8], im[BASE + j + 8], j * 2);
}

// 512-point complex FFT (= real FFT of n_fft = 1024) on HALF a warp: the 16 lanes of a half own one
// transform, so a warp carries two frames at once.  Element index = 16 * slot + hl (hl = lane & 15) on the
// way in (bit-reversed slots, like the DIT first pass of warp_fft1024); 512 = 32 x 16:
//     pass 1  32-point FFT over the slot index (registers)
//     twiddle W_512^(hl * k1)          tw[k1 * 16 + hl]
//     32 x 16 transpose through the half's shared-memory plane (pitch 17; the second half's plane starts
//             16 banks further, so the 32 lanes of a warp-wide access never collide)
//     pass 2  two 16-point FFTs: slots 0..15 for k1 = hl, slots 16..31 for k1 = hl + 16
// Result: slot s holds Z[k], k = hl + (s & 16) + 32 * (s & 15).
constexpr int HPITCH = 17;
constexpr int HPLANE_ELEMS = 32 * HPITCH + 16;      // 560: + 16 puts the second half on the other 16 banks
template <typename T>
SSTTS_D void halfwarp_fft512(T (&re)[32], T (&im)[32], T* plane, int half, const typename cx_of<T>::type* tw, int hl) {
  typedef typename cx_of<T>::type C;
  // indexed from the warp's plane (not through a per-half pointer) so that every access compiles to LDS / STS
  T* xh = plane + half * HPLANE_ELEMS;
  fft32<T, false, true>(re, im);
#pragma unroll
  for (int k1 = 1; k1 < 32; ++k1) {
    const C w = tw[k1 * 16 + hl];
    const T vr = re[k1], vi = im[k1];
    re[k1] = vr * w.x - vi * w.y;
    im[k1] = vr * w.y + vi * w.x;
  }
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) xh[k1 * HPITCH + hl] = re[k1];
  __syncwarp();
#pragma unroll
  for (int s = 0; s < 32; ++s) re[(s & 16) + brev4(s & 15)] = xh[(hl + (s & 16)) * HPITCH + (s & 15)];
  __syncwarp();
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) xh[k1 * HPITCH + hl] = im[k1];
  __syncwarp();
#pragma unroll
  for (int s = 0; s < 32; ++s) im[(s & 16) + brev4(s & 15)] = xh[(hl + (s & 16)) * HPITCH + (s & 15)];
  __syncwarp();
  fft16_dit<T, false, 0>(re, im);
  fft16_dit<T, false, 16>(re, im);
}

// Inverse of halfwarp_fft512 (unscaled: 512 x the inverse DFT), the same steps backwards:
//     pass 1  two 16-point inverse FFTs over k2 (slots 0..15: k1 = hl, slots 16..31: k1 = hl + 16)
//     16 x 32 transpose back through the half's plane (lane hl then holds n2 = hl and all 32 k1)
//     twiddle conj W_512^(hl * k1)     same table walk as the forward transform
//     pass 2  32-point inverse FFT over k1
// Input : Z[k], k = hl + (s & 16) + 32 k2, in slot (s & 16) + brev4(k2) (bit-reversed inside each 16-block --
//         the producer is unrolled code, so this is only a renaming of registers).
// Output: z[16 * slot + hl] in natural slots.
template <typename T>
SSTTS_D void halfwarp_ifft512(T (&re)[32], T (&im)[32], T* plane, int half, const typename cx_of<T>::type* tw, int hl) {
  typedef typename cx_of<T>::type C;
  T* xh = plane + half * HPLANE_ELEMS;
  fft16_dit<T, true, 0>(re, im);
  fft16_dit<T, true, 16>(re, im);
#pragma unroll
  for (int s = 0; s < 32; ++s) xh[(hl + (s & 16)) * HPITCH + (s & 15)] = re[s];
  __syncwarp();
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) re[brev5(k1)] = xh[k1 * HPITCH + hl];
  __syncwarp();
#pragma unroll
  for (int s = 0; s < 32; ++s) xh[(hl + (s & 16)) * HPITCH + (s & 15)] = im[s];
  __syncwarp();
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) im[brev5(k1)] = xh[k1 * HPITCH + hl];
  __syncwarp();
#pragma unroll
  for (int k1 = 1; k1 < 32; ++k1) {
    const C w = tw[k1 * 16 + hl];
    const T vr = re[brev5(k1)], vi = im[brev5(k1)];
    re[brev5(k1)] = vr * w.x + vi * w.y;
    im[brev5(k1)] = vi * w.x - vr * w.y;
  }
  fft32<T, true, true>(re, im);
}

// Per-warp transpose tile: one scalar plane of 32 x 33 (pitch 33 keeps the column-wise stores
// and the row-wise loads bank-conflict free).  Real and imaginary parts go through the same
// plane one after the other, which halves the shared memory per warp compared with a complex
// tile at the same number of shared-memory wavefronts.
constexpr int XPITCH = 33;
constexpr int XPLANE_ELEMS = 32 * XPITCH;

// 1024-point complex FFT across one warp; element index = 32 * slot + lane on both sides.
// Each 32-point pass is either DIF (natural slots in, bit-reversed out) or DIT (bit-reversed in,
// natural out); P1_DIT / P2_DIT select them.  The slot order between the passes is free because
// the data goes through the shared-memory transpose, so the only constraint is on the caller's
// side: pass-1 input in (P1_DIT ? bit-reversed : natural) slots, result in
// (P2_DIT ? natural : bit-reversed) slots.  DIT passes use the FMA-fused butterfly.
// With P1_DIT, input elements (slots before bit reversal) outside [ZLO, ZHI] must be zero.
// tw[a * 32 + b] = exp(-2 pi i a b / 1024); xp is this warp's private XPLANE_ELEMS plane.
template <typename T, bool INV, bool P1_DIT, bool P2_DIT, int ZLO = 0, int ZHI = 31>
SSTTS_D void warp_fft1024(T (&re)[32], T (&im)[32], T* xp, const typename cx_of<T>::type* tw,
                          int lane) {
  typedef typename cx_of<T>::type C;
  fft32<T, INV, P1_DIT, (P1_DIT ? ZLO : 0), (P1_DIT ? ZHI : 31)>(re, im);
  // The twiddles W^(k1 n2) are symmetric in (k1, n2): with a DIT second pass they are applied AFTER the
  // transpose, fused into its first butterfly stage (tw[n2 * 32 + lane], same conflict-free table walk).
  constexpr bool FUSE_TW = P2_DIT && (SSTTS_FUSE_TWIDDLE != 0);
  if (!FUSE_TW) {
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) {
      const int p = P1_DIT ? k1 : brev5(k1);  // slot holding pass-1 result k1
      const C w = tw[k1 * 32 + lane];
      const T vr = re[p], vi = im[p];
      if (!INV) { re[p] = vr * w.x - vi * w.y; im[p] = vr * w.y + vi * w.x; }
      else      { re[p] = vr * w.x + vi * w.y; im[p] = vi * w.x - vr * w.y; }
    }
  }
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) xp[k1 * XPITCH + lane] = re[P1_DIT ? k1 : brev5(k1)];
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) re[P2_DIT ? brev5(n2) : n2] = xp[lane * XPITCH + n2];
  __syncwarp();
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) xp[k1 * XPITCH + lane] = im[P1_DIT ? k1 : brev5(k1)];
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) im[P2_DIT ? brev5(n2) : n2] = xp[lane * XPITCH + n2];
  __syncwarp();
  if (FUSE_TW) {
    dit_stage2_twiddled<T, INV>(re, im, tw, lane);
    fft32<T, INV, true, 0, 31, true>(re, im);
  } else {
    fft32<T, INV, P2_DIT>(re, im);
  }
}

// The same transform with a COMPLEX transpose plane (32 x 33 values of 8 / 16 bytes): real and imaginary parts
// travel together, so the exchange is 32 + 32 shared-memory instructions instead of 64 + 64 at the same number
// of wavefronts (column-wise stores are contiguous; row-wise 8-byte loads at pitch 33 are conflict-free per
// half-warp).  `mid()` runs between the exchange and the second pass -- the Griffin-Lim kernel issues the
// asynchronous copy of the frame's |S| row there, into memory the plane overlaps.
// The twiddle table of this variant is PAIRED (paired_twiddle_index): entry (k1 >> 1) * 32 + lane holds the twiddles
// of k1 and k1 + 1 for this lane, so one 16 / 32-byte load fetches two of them (16 table loads instead of 31).
template <typename C> struct alignas(2 * sizeof(C)) TwiddlePair { C a, b; };
SSTTS_HD constexpr int paired_twiddle_index(int k1, int lane) { return (((k1 >> 1) * 32 + lane) << 1) | (k1 & 1); }

template <typename T, bool INV, bool P1_DIT, bool P2_DIT, int ZLO, int ZHI, typename Mid>
SSTTS_D void warp_fft1024_cx(T (&re)[32], T (&im)[32], typename cx_of<T>::type* xc,
                             const typename cx_of<T>::type* tw, int lane, Mid mid) {
  typedef typename cx_of<T>::type C;
  fft32<T, INV, P1_DIT, (P1_DIT ? ZLO : 0), (P1_DIT ? ZHI : 31)>(re, im);
  const TwiddlePair<C>* tw2 = reinterpret_cast<const TwiddlePair<C>*>(tw);
  SSTTS_CHECK_ALIGNED(tw2, sizeof(TwiddlePair<C>));
  SSTTS_CHECK_ALIGNED(xc, sizeof(C));
#pragma unroll
  for (int kp = 0; kp < 32; kp += 2) {
    const TwiddlePair<C> wp = tw2[(kp >> 1) * 32 + lane];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k1 = kp + h;
      if (k1 == 0) continue;      // W^0 = 1
      const int p = P1_DIT ? k1 : brev5(k1);
      const C w = h ? wp.b : wp.a;
      const T vr = re[p], vi = im[p];
      if (!INV) { re[p] = vr * w.x - vi * w.y; im[p] = vr * w.y + vi * w.x; }
      else      { re[p] = vr * w.x + vi * w.y; im[p] = vi * w.x - vr * w.y; }
    }
  }
  __syncwarp();   // the plane may overlap data other lanes were still reading (the |S| row in the core)
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    C v; v.x = re[P1_DIT ? k1 : brev5(k1)]; v.y = im[P1_DIT ? k1 : brev5(k1)];
    xc[k1 * XPITCH + lane] = v;
  }
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) {
    const C v = xc[lane * XPITCH + n2];
    re[P2_DIT ? brev5(n2) : n2] = v.x; im[P2_DIT ? brev5(n2) : n2] = v.y;
  }
  __syncwarp();
  mid();
  fft32<T, INV, P2_DIT>(re, im);
}

}  // namespace sstts
