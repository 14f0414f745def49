// CPU emulation driver -- TEST INFRASTRUCTURE ONLY (see cpu_simt.h).
//
// Compiles the real kernel sources (stft_kernels.cuh, fft_warp.cuh, host_plan.h) with g++ and
// replays the launch sequence of sstts.cu on the SIMT emulator, so that the CPU-only test tier
// can check the kernels' index maths, barrier structure and shuffle pairing against the oracle
// on small inputs.  Built by tests/emu/build.py into tests/emu/libsstts_emu.so.
#define SSTTS_CPU_EMU 1
#include <math.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "cpu_simt.h"
#include "host_plan.h"
#include "stft_kernels.cuh"

using namespace sstts;

namespace {

template <typename T> struct HostTables {
  std::vector<typename cx_of<T>::type> tw, w2;
  std::vector<T> win;
  StftTables<T> view() const {
    StftTables<T> t; t.tw1024 = tw.data(); t.w2048 = w2.data(); t.window = win.data(); return t;
  }
};
template <typename T> void fill_tables(int win, HostTables<T>& H, bool native1024 = false) {
  std::vector<double> tw, w2, wn;
  if (native1024) make_tables_native1024(win, tw, w2, wn);
  else make_tables(win, tw, w2, wn);
  H.tw.resize(1024); H.w2.resize(1024); H.win.resize(win);
  for (int i = 0; i < 1024; ++i) {
    H.tw[i].x = (T)tw[2 * i]; H.tw[i].y = (T)tw[2 * i + 1];
    H.w2[i].x = (T)w2[2 * i]; H.w2[i].y = (T)w2[2 * i + 1];
  }
  for (int i = 0; i < win; ++i) H.win[i] = (T)wn[i];
}

// one-warp test kernel around warp_fft1024
template <typename T, bool INV, bool DIT>
void fft_test_kernel(const T* in, T* out, const typename cx_of<T>::type* tw) {
  typedef typename cx_of<T>::type C;
  SSTTS_DYN_SMEM(smem);
  T* xt = reinterpret_cast<T*>(smem);
  const int lane = threadIdx.x & 31;
  T re[32], im[32];
  // forward: DIT/DIT (bit-reversed slots in, natural out); inverse: DIF/DIT (natural in and out)
  constexpr bool P1_DIT = !INV;
  for (int r = 0; r < 32; ++r) {
    const int slot = P1_DIT ? brev5(r) : r;
    re[slot] = in[2 * (32 * r + lane)];
    im[slot] = in[2 * (32 * r + lane) + 1];
  }
  warp_fft1024<T, INV, P1_DIT, true>(re, im, xt, tw, lane);
  for (int r = 0; r < 32; ++r) {
    out[2 * (32 * r + lane)] = re[r];
    out[2 * (32 * r + lane) + 1] = im[r];
  }
}

template <typename T, typename G, int W>
int emu_gl_run(int win, int hop, int n_utts, const long long* frame_off, const float* mag,
               const float* phase0, int n_iter, float* wav_out, double* mse_frame, int grid_cap, int n_fft) {
  GLPlanHost H;
  std::string err;
  constexpr bool NATIVE = G::kNative1024;
  if (!build_gl_plan(n_utts, frame_off, win, hop, H, err, n_fft, NATIVE ? 2 * W : W)) { fprintf(stderr, "plan: %s\n", err.c_str()); return -1; }
  if (H.tiles.empty()) return 0;
  HostTables<T> tabs;
  fill_tables<T>(win, tabs, NATIVE);
  std::vector<T> ws(4 * (size_t)H.total_pad, (T)NAN);  // poison: stale reads show up as NaN
  T* buf[4] = {ws.data(), ws.data() + H.total_pad, ws.data() + 2 * H.total_pad, ws.data() + 3 * H.total_pad};
  GLArgs<T> A;
  A.mag = mag; A.phase0 = reinterpret_cast<const float2*>(phase0);
  A.phase_seed = 0x5eedULL; A.phase_first = 0;     // used when phase0 == nullptr
  A.frame_off = H.frame_off.data(); A.pad_off = H.pad_off.data();
  A.tiles = H.tiles.data(); A.n_tiles = (int)H.tiles.size();
  A.tab = tabs.view(); A.mse_frame = nullptr;
  A.win = win; A.hop = hop; A.span_max = H.span_max; A.n_fft = n_fft;
  // SSTTS_GL_STAGING=bulk: the bulk-copy staging variant of the float32 iteration kernel (tests run both)
  const char* stg = getenv("SSTTS_GL_STAGING");
  const bool bulk = !NATIVE && sizeof(T) == 4 && stg && std::string(stg) == "bulk";
  const size_t smem = gl_step_smem_bytes<T>(W, win, hop, H.span_max, bulk, NATIVE);
  int grid = A.n_tiles < grid_cap ? A.n_tiles : grid_cap;
  // one launch of the step kernel in the selected staging variant (the bulk variant exists for the 2048-point kernels)
  auto step = [&](auto from_phase, auto want_mse) {
    constexpr bool FP = decltype(from_phase)::value, MSE = decltype(want_mse)::value;
    if constexpr (!NATIVE) {
      if (bulk) { emu::launch(dim3(grid), dim3(W * 32), smem, [&]() { gl_step_kernel<T, G, W, FP, MSE, true>(A); }); return; }
    }
    emu::launch(dim3(grid), dim3(W * 32), smem, [&]() { gl_step_kernel<T, G, W, FP, MSE>(A); });
  };
  A.pin0 = nullptr; A.pin1 = nullptr; A.pout0 = buf[0]; A.pout1 = buf[1];
  step(std::true_type(), std::false_type());
  int cur = 0;
  for (int it = 0; it < n_iter; ++it) {
    A.pin0 = buf[2 * cur]; A.pin1 = buf[2 * cur + 1];
    A.pout0 = buf[2 * (cur ^ 1)]; A.pout1 = buf[2 * (cur ^ 1) + 1];
    if (it == n_iter - 1 && mse_frame) {
      A.mse_frame = mse_frame;
      step(std::false_type(), std::true_type());
    } else {
      step(std::false_type(), std::false_type());
    }
    cur ^= 1;
  }
  GLFinalArgs<T> F;
  F.pin0 = buf[2 * cur]; F.pin1 = buf[2 * cur + 1];
  F.frame_off = H.frame_off.data(); F.pad_off = H.pad_off.data(); F.sample_off = H.sample_off.data();
  F.tiles = H.tiles.data(); F.n_tiles = A.n_tiles; F.window = tabs.win.data(); F.wav_out = wav_out;
  F.win = win; F.hop = hop; F.n_fft = n_fft;
  emu::launch(dim3(grid), dim3(256), sizeof(T) * (round_up4(win) + round_up4(hop)), [&]() { gl_finalize_kernel<T, G, 256>(F); });
  return 0;
}

template <typename T, typename G, int W>
int emu_feat_run(int n_fft, int win, int hop, int sr, int n_mels, double fmin, double fmax, int n_clips,
                 const long long* sample_off, int reduction, const float* wav, float* spec, float* lin,
                 float* mel, double* melraw, double* minmax, int normalize, double lin_ref,
                 double lin_max, double mel_ref, double mel_max, double power, int grid_cap, int fast_mode) {
  FeatPlanHost H;
  std::string err;
  std::vector<long long> cs(n_clips), cl(n_clips);
  for (int c = 0; c < n_clips; ++c) { cs[c] = sample_off[c]; cl[c] = sample_off[c + 1] - sample_off[c]; }
  if (!build_feat_plan(n_clips, cs.data(), cl.data(), n_fft, win, hop, reduction, H, err, sizeof(T) == 8)) { fprintf(stderr, "plan: %s\n", err.c_str()); return -1; }
  HostTables<T> tabs;
  fill_tables<T>(win, tabs, G::kNative1024);
  MelCSR M;
  std::vector<T> mw;
  MelPadded MP;
  if (n_mels > 0) {
    make_mel_csr(sr, n_fft, n_mels, fmin, fmax > 0 ? fmax : sr / 2.0, M);
    mw.assign(M.w.begin(), M.w.end());
    make_mel_padded(M, n_mels, G::kNative1024 ? HMAG : FEAT_PLANE_ELEMS, MP);
  }
  // same selection rule as sstts.cu:run_features
  const bool fast = fast_mode && MP.ok && !spec && !melraw && power == 1.0 &&
                    (G::kNative1024 ? (lin || mel || minmax) : (n_fft == NFFT && lin && mel && !minmax));
  if (fast_mode && !fast) { fprintf(stderr, "dB-feature mode does not apply to this request\n"); return -2; }
  std::vector<long long> mm(4 * (size_t)n_clips);
  FeatArgs<T> A;
  A.wav = wav; A.sample_off = H.sample_off.data(); A.sample_len = H.sample_len.data(); A.frame_off = H.frame_off.data();
  A.row_off = H.row_off.data(); A.tiles = H.tiles.data(); A.n_tiles = (int)H.tiles.size();
  A.tab = tabs.view();
  A.mel_ptr = M.ptr.data(); A.mel_k0 = M.k0.data(); A.mel_w = mw.data(); A.n_mels = n_mels; A.mel_nnz = (int)mw.size();
  A.spec_out = reinterpret_cast<float2*>(spec); A.lin_out = lin; A.mel_out = mel; A.melraw_out = melraw;
  A.minmax_out = minmax ? mm.data() : nullptr;
  A.lin_ref_db = (float)lin_ref; A.lin_range_db = (float)(fabs(lin_ref) + fabs(lin_max));
  A.mel_ref_db = mel_ref; A.mel_range_db = fabs(mel_ref) + fabs(mel_max);
  A.mel_power = (float)power; A.normalize = normalize;
  A.win = win; A.hop = hop; A.span_max = H.span_max; A.n_fft = n_fft;
  A.melp_w = MP.w.data(); A.melp_slots = fast ? MP.n_slots : 0; A.melp_total = fast ? MP.total : 0;
  for (int j = 0; j < 4; ++j) { A.melp_len[j] = MP.len[j]; A.melp_woff[j] = MP.woff[j]; A.melp_mbase[j] = MP.mbase[j]; }
  if (minmax) for (size_t i = 0; i < mm.size(); ++i) mm[i] = (i & 1) ? encode_ordered(-1e300) : encode_ordered(1e300);
  const size_t smem = stft_feature_smem_bytes<T>(W, win, H.span_max, n_mels, (int)mw.size(), A.melp_total,
                                                 feat_plane_elems<G>());
  int grid = A.n_tiles < grid_cap ? A.n_tiles : grid_cap;
  if (fast) emu::launch(dim3(grid), dim3(W * 32), smem, [&]() { stft_feature_kernel<T, G, W, FeatMode::kDbFeatures>(A); });
  else emu::launch(dim3(grid), dim3(W * 32), smem, [&]() { stft_feature_kernel<T, G, W, FeatMode::kGeneric>(A); });
  if (minmax) for (size_t i = 0; i < mm.size(); ++i) {
    long long c = mm[i]; c = c >= 0 ? c : (c ^ 0x7fffffffffffffffLL);
    std::memcpy(&minmax[i], &c, 8);
  }
  return 0;
}

}  // namespace

extern "C" {

// mode: 0 = forward DIF, 1 = inverse DIT (unscaled); prec: 0 float, 1 double (buffers are double)
int emu_fft1024(const double* in, double* out, int inverse, int prec) {
  if (prec == 0) {
    HostTables<float> tabs; fill_tables<float>(1102, tabs);
    std::vector<float> fi(2048), fo(2048);
    for (int i = 0; i < 2048; ++i) fi[i] = (float)in[i];
    const size_t smem = sizeof(float) * XPLANE_ELEMS;
    if (!inverse) emu::launch(dim3(1), dim3(32), smem, [&]() { fft_test_kernel<float, false, false>(fi.data(), fo.data(), tabs.tw.data()); });
    else emu::launch(dim3(1), dim3(32), smem, [&]() { fft_test_kernel<float, true, true>(fi.data(), fo.data(), tabs.tw.data()); });
    for (int i = 0; i < 2048; ++i) out[i] = fo[i];
  } else {
    HostTables<double> tabs; fill_tables<double>(1102, tabs);
    const size_t smem = sizeof(double) * XPLANE_ELEMS;
    if (!inverse) emu::launch(dim3(1), dim3(32), smem, [&]() { fft_test_kernel<double, false, false>(in, out, tabs.tw.data()); });
    else emu::launch(dim3(1), dim3(32), smem, [&]() { fft_test_kernel<double, true, true>(in, out, tabs.tw.data()); });
  }
  return 0;
}

int emu_griffin_lim(int win, int hop, int prec, int n_utts, const long long* frame_off, const float* mag,
                    const float* phase0, int n_iter, float* wav_out, double* mse_frame, int grid_cap, int n_fft) {
  const bool model = (win == 1102 && hop == 275 && n_fft == 2048);
#define GL_ARGS win, hop, n_utts, frame_off, mag, phase0, n_iter, wav_out, mse_frame, grid_cap, n_fft
  // same selection rule as sstts.cu (232448 bytes: opt-in shared memory per block on sm_100)
  const bool native = prec == 1 ? gl_native_1024<double>(n_fft, win, hop, kWarps, 232448)
                                : gl_native_1024<float>(n_fft, win, hop, kGlWarps, 232448);
  if (native) {
    const bool stats = win == 1024 && hop == 256;
    if (prec == 1)
      return stats ? emu_gl_run<double, NativeGeom1024<1024, 256>, kWarps>(GL_ARGS) : emu_gl_run<double, DynGeom1024, kWarps>(GL_ARGS);
    return stats ? emu_gl_run<float, NativeGeom1024<1024, 256>, kGlWarps>(GL_ARGS) : emu_gl_run<float, DynGeom1024, kGlWarps>(GL_ARGS);
  }
  if (prec == 1)
    return model ? emu_gl_run<double, StaticGeom<1102, 275, 2048>, kWarps>(GL_ARGS) : emu_gl_run<double, DynGeom, kWarps>(GL_ARGS);
  return model ? emu_gl_run<float, StaticGeom<1102, 275, 2048>, kGlWarps>(GL_ARGS) : emu_gl_run<float, DynGeom, kGlWarps>(GL_ARGS);
#undef GL_ARGS
}

int emu_stft_features(int n_fft, int win, int hop, int prec, int sr, int n_mels, double fmin, double fmax, int n_clips,
                      const long long* sample_off, int reduction, const float* wav, float* spec, float* lin,
                      float* mel, double* melraw, double* minmax, int normalize, double lin_ref,
                      double lin_max, double mel_ref, double mel_max, double power, int grid_cap, int fast_mode) {
  const bool model = (n_fft == 2048 && win == 1102 && hop == 275);
  const bool stats = (n_fft == 1024 && win == 1024 && hop == 256);
#define FEAT_ARGS n_fft, win, hop, sr, n_mels, fmin, fmax, n_clips, sample_off, reduction, wav, spec, lin, mel, melraw, minmax, normalize, lin_ref, lin_max, mel_ref, mel_max, power, grid_cap, fast_mode
  if (prec == 1) {
    if (model) return emu_feat_run<double, StaticGeom<1102, 275, 2048>, kFeatWarpsF64>(FEAT_ARGS);
    if (stats) return emu_feat_run<double, NativeGeom1024<1024, 256>, kFeatWarpsF64Native>(FEAT_ARGS);
    if (feat_native_1024(n_fft)) return emu_feat_run<double, DynGeom1024, kFeatWarpsF64Native>(FEAT_ARGS);
    return emu_feat_run<double, DynGeom, kFeatWarpsF64>(FEAT_ARGS);
  }
  if (model) return emu_feat_run<float, StaticGeom<1102, 275, 2048>, kWarps>(FEAT_ARGS);
  if (stats) return emu_feat_run<float, NativeGeom1024<1024, 256>, kWarps>(FEAT_ARGS);
  if (feat_native_1024(n_fft)) return emu_feat_run<float, DynGeom1024, kWarps>(FEAT_ARGS);
  return emu_feat_run<float, DynGeom, kWarps>(FEAT_ARGS);
#undef FEAT_ARGS
}

int emu_trim_bounds(const float* wav, int n_clips, const long long* clip_start, const long long* clip_len,
                    double top_db, int frame_length, int hop_length, long long* bounds) {
  emu::launch(dim3(n_clips < 3 ? n_clips : 3), dim3(256), 0, [&]() {
    trim_bounds_kernel<256>(wav, clip_start, clip_len, n_clips, frame_length, hop_length, top_db, bounds);
  });
  return 0;
}

int emu_mel_basis(int sr, int n_fft, int n_mels, double fmin, double fmax, double* dense_out) {
  MelCSR M;
  std::vector<double> dense;
  make_mel_csr(sr, n_fft, n_mels, fmin, fmax, M, &dense);
  std::memcpy(dense_out, dense.data(), dense.size() * sizeof(double));
  return 0;
}

}  // extern "C"
