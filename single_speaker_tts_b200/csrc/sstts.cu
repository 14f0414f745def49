// libsstts.so -- C ABI (include/sstts.h) over the sm_100a kernels in stft_kernels.cuh.
//
// Host side of the library: argument checking, tile/offset planning (host_plan.h), device
// tables, kernel dispatch on (precision, geometry) and launch configuration.  No torch types,
// no C++ exceptions across the boundary, no CPU fallback: without a CUDA device every compute
// entry point fails with SSTTS_ERR_NO_DEVICE / SSTTS_ERR_CUDA.
#include "../../include/sstts.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <tuple>
#include <vector>

#include "host_plan.h"
#include "stft_kernels.cuh"

using namespace sstts;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? SSTTS_ERR_NO_DEVICE
                                                                          : SSTTS_ERR_CUDA,
              std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                             \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call);    \
  } while (0)

template <typename T>
int upload(const std::vector<T>& h, T** d) {
  *d = nullptr;
  if (h.empty()) return 0;
  CU(cudaMalloc((void**)d, h.size() * sizeof(T)));
  CU(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// Device copies of the transform tables in one precision.
struct DeviceTables {
  void* tw1024 = nullptr;
  void* w2048 = nullptr;
  void* window = nullptr;
};

template <typename T>
int upload_tables(int win, bool native1024, DeviceTables& D) {
  std::vector<double> tw, w2, wn;
  if (native1024) make_tables_native1024(win, tw, w2, wn);
  else make_tables(win, tw, w2, wn);
  std::vector<T> a(tw.begin(), tw.end()), b(w2.begin(), w2.end()), c(wn.begin(), wn.end());
  T *da, *db, *dc;
  int rc;
  if ((rc = upload(a, &da))) return rc;
  D.tw1024 = da;
  if ((rc = upload(b, &db))) return rc;
  D.w2048 = db;
  if ((rc = upload(c, &dc))) return rc;
  D.window = dc;
  return 0;
}

template <typename T>
StftTables<T> tables_view(const DeviceTables& D) {
  StftTables<T> t;
  t.tw1024 = reinterpret_cast<const typename cx_of<T>::type*>(D.tw1024);
  t.w2048 = reinterpret_cast<const typename cx_of<T>::type*>(D.w2048);
  t.window = reinterpret_cast<const T*>(D.window);
  return t;
}

// Everything that depends only on the configuration -- transform tables, window, mel filterbank --
// lives in a process-wide cache keyed by (device, precision, geometry, filterbank): a plan for a new
// batch shape then costs one allocation and one copy (its offset / tile tables, packed into one blob)
// instead of a dozen.  Entries are immutable and never freed (a few hundred KB per configuration).
struct StaticTables {
  DeviceTables tab;
  MelCSR mel;
  MelPadded melp;
  std::vector<double> mel_dense;
  int* d_mel_ptr = nullptr;
  int* d_mel_k0 = nullptr;
  void* d_mel_w = nullptr;
  float* d_melp_w = nullptr;
};
typedef std::tuple<int, int, int, int, int, int, double, double, int> StaticKey;
std::mutex g_static_mu;
std::map<StaticKey, std::shared_ptr<StaticTables> > g_static;

// native1024: tables of the native n_fft = 1024 transform (halfwarp_fft512) instead of the 2048-point ones.
int get_static_tables(const sstts_stft_config* cfg, int device, bool with_mel, bool native1024,
                      std::shared_ptr<StaticTables>* out) {
  const bool f64 = cfg->precision == SSTTS_F64;
  const int n_mels = with_mel ? cfg->n_mels : 0;
  const double fmax = n_mels > 0 ? (cfg->mel_fmax > 0 ? cfg->mel_fmax : cfg->sampling_rate / 2.0) : 0.0;
  const StaticKey key(device, cfg->precision, cfg->n_fft, cfg->win_length, n_mels > 0 ? cfg->sampling_rate : 0, n_mels,
                      n_mels > 0 ? cfg->mel_fmin : 0.0, fmax, native1024 ? 1 : 0);
  std::lock_guard<std::mutex> lock(g_static_mu);
  auto it = g_static.find(key);
  if (it != g_static.end()) { *out = it->second; return 0; }
  std::shared_ptr<StaticTables> S(new StaticTables());
  int rc = f64 ? upload_tables<double>(cfg->win_length, native1024, S->tab)
               : upload_tables<float>(cfg->win_length, native1024, S->tab);
  if (!rc && n_mels > 0) {
    make_mel_csr(cfg->sampling_rate, cfg->n_fft, n_mels, cfg->mel_fmin, fmax, S->mel, &S->mel_dense);
    make_mel_padded(S->mel, n_mels, native1024 ? HMAG : FEAT_PLANE_ELEMS, S->melp);
    rc = upload(S->mel.ptr, &S->d_mel_ptr);
    if (!rc) rc = upload(S->mel.k0, &S->d_mel_k0);
    if (!rc && S->melp.ok) rc = upload(S->melp.w, &S->d_melp_w);
    if (!rc) {
      if (f64) { double* d; rc = upload(S->mel.w, &d); S->d_mel_w = d; }
      else { std::vector<float> wf(S->mel.w.begin(), S->mel.w.end()); float* d; rc = upload(wf, &d); S->d_mel_w = d; }
    }
  }
  if (rc) return rc;     // partially filled entry is dropped (its few allocations leak only on a CUDA error)
  g_static[key] = S;
  *out = S;
  return 0;
}

// One device allocation holding several host arrays back to back (256-byte aligned sections).
struct DeviceBlob {
  std::vector<char> host;
  char* dev = nullptr;
  template <typename T> size_t add(const std::vector<T>& v) {
    const size_t off = (host.size() + 255) & ~(size_t)255;
    host.resize(off + v.size() * sizeof(T));
    if (!v.empty()) std::memcpy(host.data() + off, v.data(), v.size() * sizeof(T));
    return off;
  }
  // Stream-ordered allocation from the device's default memory pool (kept warm: the release threshold
  // is raised once) on a private non-blocking stream, so a plan for a new batch shape costs a few
  // microseconds and never waits for the caller's streams -- a cudaMalloc / cudaMemcpy / cudaFree
  // triple maps and unmaps memory and synchronises with the legacy default stream.  A plan must outlive
  // the work that uses it (see sstts.h).
  static cudaStream_t plan_stream(int device) {
    static std::mutex mu;
    static std::map<int, cudaStream_t> streams;
    std::lock_guard<std::mutex> lock(mu);
    auto it = streams.find(device);
    if (it != streams.end()) return it->second;
    cudaStream_t st = nullptr;
    cudaMemPool_t pool;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); st = nullptr; }
    if (st && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    streams[device] = st;
    return st;
  }
  int commit() {
    if (host.empty()) return 0;
    int device = 0;
    CU(cudaGetDevice(&device));
    stream = plan_stream(device);
    if (stream && cudaMallocAsync((void**)&dev, host.size(), stream) == cudaSuccess) {
      pooled = true;
      CU(cudaMemcpyAsync(dev, host.data(), host.size(), cudaMemcpyHostToDevice, stream));
      CU(cudaStreamSynchronize(stream));
    } else {
      (void)cudaGetLastError();
      CU(cudaMalloc((void**)&dev, host.size()));
      CU(cudaMemcpy(dev, host.data(), host.size(), cudaMemcpyHostToDevice));
    }
    std::vector<char>().swap(host);
    return 0;
  }
  template <typename T> T* at(size_t off) const { return reinterpret_cast<T*>(dev + off); }
  void release() {
    if (dev) { if (pooled) cudaFreeAsync(dev, stream); else cudaFree(dev); }
    dev = nullptr;
  }
  bool pooled = false;
  cudaStream_t stream = nullptr;
};

int check_config(const sstts_stft_config* cfg, bool allow_embedded) {
  if (!cfg) return fail(SSTTS_ERR_INVALID, "config is NULL");
  (void)allow_embedded;
  if (cfg->n_fft != 2048 && cfg->n_fft != 1024 && cfg->n_fft != 512)
    return fail(SSTTS_ERR_INVALID, "n_fft must be 2048, 1024 or 512");
  if (cfg->win_length < 2 || cfg->win_length > cfg->n_fft || ((cfg->n_fft - cfg->win_length) & 1))
    return fail(SSTTS_ERR_INVALID, "win_length must be in [2, n_fft] with n_fft - win_length even");
  if (cfg->hop_length < 1) return fail(SSTTS_ERR_INVALID, "hop_length must be >= 1");
  if (cfg->precision != SSTTS_F32 && cfg->precision != SSTTS_F64)
    return fail(SSTTS_ERR_INVALID, "precision must be SSTTS_F32 or SSTTS_F64");
  return 0;
}

bool is_model_geometry(int n_fft, int win, int hop) { return n_fft == 2048 && win == 1102 && hop == 275; }
bool is_stats_geometry(int n_fft, int win, int hop) { return n_fft == 1024 && win == 1024 && hop == 256; }
typedef StaticGeom<1102, 275, 2048> ModelGeom;   // tacotron/params/model.py:13-24
typedef NativeGeom1024<1024, 256> StatsGeom;     // datasets/statistics.py:31-34 (native 512-point complex transform)

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

// The dynamic shared memory limit of a kernel is process-wide state: concurrent callers with different
// batch shapes (different tile spans) would race if each set "its" size, so every caller sets the same
// value, the device's opt-in maximum; the occupancy query below uses the actual size.
int max_optin_smem() {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
  return v;
}

// Per (kernel, device, threads, shared-memory size): the opt-in shared-memory attribute is set once and the
// occupancy is queried once -- both are driver calls of several microseconds, far too slow to repeat on every
// call of the single-utterance path (tacotron/serve.py:69-72 calls spectrogram_to_wav per sentence).
struct KernelConfigKey {
  const void* fn; int device, threads; size_t smem;
  bool operator<(const KernelConfigKey& o) const {
    return std::tie(fn, device, threads, smem) < std::tie(o.fn, o.device, o.threads, o.smem);
  }
};
std::mutex g_kcfg_mu;
std::map<KernelConfigKey, int> g_kcfg;          // -> resident blocks per SM
std::map<std::pair<const void*, int>, bool> g_kattr;   // (kernel, device) -> attribute set

template <typename K>
int configure_kernel(K kernel, int threads, size_t smem, int* blocks_per_sm) {
  int dev = 0;
  CU(cudaGetDevice(&dev));
  const KernelConfigKey key{reinterpret_cast<const void*>(kernel), dev, threads, smem};
  std::lock_guard<std::mutex> lock(g_kcfg_mu);
  auto it = g_kcfg.find(key);
  if (it != g_kcfg.end()) { *blocks_per_sm = it->second; return 0; }
  const int limit = max_optin_smem();
  if (limit <= 0 || smem > (size_t)limit)
    return fail(SSTTS_ERR_CUDA, "kernel does not fit on this device (shared memory)");
  bool& attr_set = g_kattr[std::make_pair(key.fn, dev)];
  if (!attr_set) {
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, limit));
    attr_set = true;
  }
  int occ = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
  if (occ < 1) return fail(SSTTS_ERR_CUDA, "kernel does not fit on this device (shared memory / registers)");
  g_kcfg[key] = occ;
  *blocks_per_sm = occ;
  return 0;
}

}  // namespace

// ----------------------------------------------------------------------------------------------
// plans
// ----------------------------------------------------------------------------------------------
// CUDA graph of the iteration + finalize launches of one small call, cached per plan and buffer set.
struct GLGraphKey {
  const void* mag; const void* ws; const void* wav; const void* mse; int n_iter, variant;
  bool operator==(const GLGraphKey& o) const {
    return mag == o.mag && ws == o.ws && wav == o.wav && mse == o.mse && n_iter == o.n_iter && variant == o.variant;
  }
};
struct GLGraphEntry {
  GLGraphKey key;
  cudaGraphExec_t exec = nullptr;
  int seen = 0;               // calls with this buffer set so far
};

struct sstts_gl_plan {
  sstts_stft_config cfg;
  GLPlanHost host;
  mutable std::mutex graph_mu;
  mutable std::vector<GLGraphEntry> graphs;   // see run_griffin_lim
  std::shared_ptr<StaticTables> st;     // transform tables + window (shared, cached per configuration)
  DeviceBlob blob;                      // offset tables + tile records of this batch shape
  long long* d_frame_off = nullptr;
  long long* d_pad_off = nullptr;
  long long* d_sample_off = nullptr;
  GLTile* d_tiles = nullptr;
  int device = 0;
  int n_sms = 0;
  bool native1024 = false;              // n_fft 1024 transformed natively (two frames per warp, 16-frame tiles)
};

struct sstts_feat_plan {
  sstts_stft_config cfg;
  FeatPlanHost host;
  std::shared_ptr<StaticTables> st;     // transform tables, window, mel filterbank
  DeviceBlob blob;
  long long* d_sample_off = nullptr;
  long long* d_sample_len = nullptr;
  long long* d_frame_off = nullptr;
  long long* d_row_off = nullptr;
  FeatTile* d_tiles = nullptr;
  int device = 0;
  int n_sms = 0;
};

extern "C" {

int sstts_version(void) { return SSTTS_VERSION; }
const char* sstts_last_error(void) { return g_last_error.c_str(); }

int sstts_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  if (n < 1) return fail(SSTTS_ERR_NO_DEVICE, "no CUDA device");
  return n;
}

// ------------------------------------------------------------------------------ Griffin-Lim
int sstts_gl_plan_create(const sstts_stft_config* cfg, int n_utts, const int64_t* frame_off_host,
                         sstts_gl_plan** plan_out) {
  if (plan_out) *plan_out = nullptr;
  int rc = check_config(cfg, false);
  if (rc) return rc;
  if (!plan_out || !frame_off_host || n_utts < 1) return fail(SSTTS_ERR_INVALID, "bad plan arguments");
  sstts_gl_plan* P = new (std::nothrow) sstts_gl_plan();
  if (!P) return fail(SSTTS_ERR_INVALID, "out of host memory");
  P->cfg = *cfg;
  std::string err;
  std::vector<long long> fo(frame_off_host, frame_off_host + n_utts + 1);
  cudaError_t e = cudaGetDevice(&P->device);
  if (e != cudaSuccess) { delete P; return cuda_fail(e, "cudaGetDevice"); }
  const bool f64 = cfg->precision == SSTTS_F64;
  const int warps = f64 ? kWarps : kGlWarps;
  const size_t smem_limit = (size_t)(max_optin_smem() > 0 ? max_optin_smem() : 0);
  P->native1024 = f64 ? gl_native_1024<double>(cfg->n_fft, cfg->win_length, cfg->hop_length, warps, smem_limit)
                      : gl_native_1024<float>(cfg->n_fft, cfg->win_length, cfg->hop_length, warps, smem_limit);
  if (!build_gl_plan(n_utts, fo.data(), cfg->win_length, cfg->hop_length, P->host, err, cfg->n_fft,
                     P->native1024 ? 2 * warps : warps)) {
    delete P;
    return fail(SSTTS_ERR_INVALID, err);
  }
  P->n_sms = sm_count();
  rc = get_static_tables(cfg, P->device, false, P->native1024, &P->st);
  if (!rc) {
    const size_t o0 = P->blob.add(P->host.frame_off), o1 = P->blob.add(P->host.pad_off);
    const size_t o2 = P->blob.add(P->host.sample_off), o3 = P->blob.add(P->host.tiles);
    rc = P->blob.commit();
    if (!rc && P->blob.dev) {
      P->d_frame_off = P->blob.at<long long>(o0); P->d_pad_off = P->blob.at<long long>(o1);
      P->d_sample_off = P->blob.at<long long>(o2); P->d_tiles = P->blob.at<GLTile>(o3);
    }
  }
  if (rc) { sstts_gl_plan_destroy(P); return rc; }
  *plan_out = P;
  return 0;
}

void sstts_gl_plan_destroy(sstts_gl_plan* P) {
  if (!P) return;
  for (auto& g : P->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  P->blob.release();
  delete P;
}

size_t sstts_gl_workspace_bytes(const sstts_gl_plan* P) {
  if (!P) return 0;
  const size_t elem = P->cfg.precision == SSTTS_F64 ? sizeof(double) : sizeof(float);
  return 4 * (size_t)P->host.total_pad * elem + 256;
}
int64_t sstts_gl_total_frames(const sstts_gl_plan* P) { return P ? P->host.total_frames : 0; }
int64_t sstts_gl_total_samples(const sstts_gl_plan* P) { return P ? P->host.total_samples : 0; }
const int64_t* sstts_gl_sample_offsets(const sstts_gl_plan* P) {
  return P ? reinterpret_cast<const int64_t*>(P->host.sample_off.data()) : nullptr;
}

}  // extern "C"

namespace {

// SSTTS_GL_STAGING=bulk selects the iteration kernel whose interior tiles are staged with bulk asynchronous
// copies (see GLSmem in stft_kernels.cuh for the A/B); read once per process.
bool gl_bulk_requested() {
  static const bool v = [] { const char* e = getenv("SSTTS_GL_STAGING"); return e && std::string(e) == "bulk"; }();
  return v;
}

// SSTTS_GL_GRAPH=0 disables the CUDA-graph path of small calls (read once per process).
bool gl_graphs_enabled() {
  static const bool v = [] { const char* e = getenv("SSTTS_GL_GRAPH"); return !(e && std::string(e) == "0"); }();
  return v;
}

// The launches after the synthesis step: n_iter iteration kernels + finalize, enqueued on `st`.
template <typename T, typename G, int W, bool BULK>
int enqueue_iterations(const sstts_gl_plan* P, GLArgs<T> A, T* const (&buf)[4], int n_iter, float* wav_out,
                       double* mse_frame, size_t smem, int grid_i, int grid_m, cudaStream_t st) {
  const GLPlanHost& H = P->host;
  const int n_sms = P->n_sms > 0 ? P->n_sms : 148;
  int cur = 0;
  for (int it = 0; it < n_iter; ++it) {
    A.pin0 = buf[2 * cur]; A.pin1 = buf[2 * cur + 1];
    A.pout0 = buf[2 * (cur ^ 1)]; A.pout1 = buf[2 * (cur ^ 1) + 1];
    if (it == n_iter - 1 && mse_frame) {
      A.mse_frame = mse_frame;
      gl_step_kernel<T, G, W, false, true, BULK><<<grid_m, W * 32, smem, st>>>(A);
    } else {
      gl_step_kernel<T, G, W, false, false, BULK><<<grid_i, W * 32, smem, st>>>(A);
    }
    CU(cudaGetLastError());
    cur ^= 1;
  }
  GLFinalArgs<T> F;
  F.pin0 = buf[2 * cur]; F.pin1 = buf[2 * cur + 1];
  F.frame_off = P->d_frame_off; F.pad_off = P->d_pad_off; F.sample_off = P->d_sample_off;
  F.tiles = P->d_tiles; F.n_tiles = A.n_tiles;
  F.window = A.tab.window;
  F.wav_out = wav_out;
  F.win = H.win; F.hop = H.hop; F.n_fft = H.n_fft;
  int grid_f = n_sms * 8;
  if (grid_f > A.n_tiles) grid_f = A.n_tiles;
  gl_finalize_kernel<T, G, 256><<<grid_f, 256, sizeof(T) * (round_up4(H.win) + round_up4(H.hop)), st>>>(F);
  CU(cudaGetLastError());
  return 0;
}

template <typename T, typename G, int W, bool BULK>
int run_griffin_lim(const sstts_gl_plan* P, const float* mag, const float* phase0, uint64_t seed,
                    int64_t first, int n_iter, void* workspace, float* wav_out, double* mse_frame,
                    cudaStream_t st) {
  const GLPlanHost& H = P->host;
  if (H.tiles.empty()) return 0;
  T* ws = reinterpret_cast<T*>(workspace);
  T* const buf[4] = {ws, ws + H.total_pad, ws + 2 * H.total_pad, ws + 3 * H.total_pad};

  GLArgs<T> A;
  A.mag = mag;
  A.phase0 = reinterpret_cast<const float2*>(phase0);
  A.phase_seed = seed; A.phase_first = first;
  A.frame_off = P->d_frame_off;
  A.pad_off = P->d_pad_off;
  A.tiles = P->d_tiles;
  A.n_tiles = (int)H.tiles.size();
  A.tab = tables_view<T>(P->st->tab);
  A.mse_frame = nullptr;
  A.win = H.win; A.hop = H.hop; A.span_max = H.span_max; A.n_fft = H.n_fft;

  const size_t smem = gl_step_smem_bytes<T>(W, H.win, H.hop, H.span_max, BULK, G::kNative1024);
  int occ_s = 0, occ_i = 0, occ_m = 0, rc;
  if ((rc = configure_kernel(gl_step_kernel<T, G, W, true, false, BULK>, W * 32, smem, &occ_s))) return rc;
  if ((rc = configure_kernel(gl_step_kernel<T, G, W, false, false, BULK>, W * 32, smem, &occ_i))) return rc;
  if (mse_frame && (rc = configure_kernel(gl_step_kernel<T, G, W, false, true, BULK>, W * 32, smem, &occ_m))) return rc;
  const int n_sms = P->n_sms > 0 ? P->n_sms : 148;
  int grid_s = n_sms * occ_s, grid_i = n_sms * occ_i, grid_m = n_sms * (occ_m > 0 ? occ_m : 1);
  if (grid_s > A.n_tiles) grid_s = A.n_tiles;
  if (grid_i > A.n_tiles) grid_i = A.n_tiles;
  if (grid_m > A.n_tiles) grid_m = A.n_tiles;

  // step 0: random phase -> wave_1 (written to buf[0], buf[1]); always a plain launch (its arguments carry the
  // per-call seed / phase pointer)
  A.pin0 = nullptr; A.pin1 = nullptr; A.pout0 = buf[0]; A.pout1 = buf[1];
  gl_step_kernel<T, G, W, true, false, BULK><<<grid_s, W * 32, smem, st>>>(A);
  CU(cudaGetLastError());

  // Calls of up to kGraphTiles tiles (the single-utterance shape of tacotron/serve.py:39-86, and the ~1,250-tile
  // sub-batches of the pipelined batch call) replay the n_iter + 1 remaining launches as ONE CUDA graph per
  // (plan, buffers): a repeated call with the same buffers captures it, later calls launch it -- one driver
  // call instead of 51, which matters when six synthesis threads (tacotron/serve.py:69-72) issue their
  // launches at the same time (p99 18.6 -> 3.8 ms) and takes 1 ms off the 33 ms of a pipelined 256-utterance
  // call.  The caching allocator of the host hands a thread the same buffers call after call, so the steady
  // state is all graph launches.  Larger launches (0.5 ms per kernel) gain nothing and stay plain.
  constexpr int kGraphTiles = 4096;
  static const int tile_limit = [] { const char* e = getenv("SSTTS_GL_GRAPH_TILES"); return e ? atoi(e) : 0; }();
  const bool small = gl_graphs_enabled() && n_iter >= 4 && A.n_tiles <= (tile_limit > 0 ? tile_limit : kGraphTiles);
  if (!small) return enqueue_iterations<T, G, W, BULK>(P, A, buf, n_iter, wav_out, mse_frame, smem, grid_i, grid_m, st);
  const GLGraphKey key{mag, workspace, wav_out, mse_frame, n_iter, (int)sizeof(T) * 2 + (BULK ? 1 : 0)};
  // capture + instantiation cost about as much as five sets of plain launches: a single-wave call (the latency
  // path) captures at its second sighting, a larger one at its fourth
  const int capture_at = A.n_tiles <= 2 * n_sms ? 2 : 4;
  cudaGraphExec_t exec = nullptr;
  bool seen = false;
  bool capture = false;
  {
    std::lock_guard<std::mutex> lock(P->graph_mu);
    for (auto& g : P->graphs)
      if (g.key == key) { seen = true; exec = g.exec; capture = !exec && ++g.seen >= capture_at; break; }
    if (!seen) {
      if (P->graphs.size() >= 16) {                       // oldest entry out
        if (P->graphs.front().exec) cudaGraphExecDestroy(P->graphs.front().exec);
        P->graphs.erase(P->graphs.begin());
      }
      GLGraphEntry e; e.key = key; e.seen = 1;
      P->graphs.push_back(e);
    }
  }
  if (exec) { CU(cudaGraphLaunch(exec, st)); return 0; }
  if (!capture) return enqueue_iterations<T, G, W, BULK>(P, A, buf, n_iter, wav_out, mse_frame, smem, grid_i, grid_m, st);
  // capture (thread-local mode: other threads keep making CUDA calls)
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    (void)cudaGetLastError();
    return enqueue_iterations<T, G, W, BULK>(P, A, buf, n_iter, wav_out, mse_frame, smem, grid_i, grid_m, st);
  }
  rc = enqueue_iterations<T, G, W, BULK>(P, A, buf, n_iter, wav_out, mse_frame, smem, grid_i, grid_m, st);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (ce != cudaSuccess || !graph) return cuda_fail(ce, "cudaStreamEndCapture");
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) return cuda_fail(ie, "cudaGraphInstantiate");
  {
    std::lock_guard<std::mutex> lock(P->graph_mu);
    bool stored = false;
    for (auto& g : P->graphs)
      if (g.key == key && !g.exec) { g.exec = exec; stored = true; break; }
    if (!stored) {                                        // evicted meanwhile, or another thread was faster
      CU(cudaGraphLaunch(exec, st));
      CU(cudaStreamSynchronize(st));
      cudaGraphExecDestroy(exec);
      return 0;
    }
  }
  CU(cudaGraphLaunch(exec, st));
  return 0;
}

// stand-alone generator of the batched API's initial phase (seeded_phasor, stft_kernels.cuh)
__global__ void random_phase_kernel(unsigned long long seed, long long first, long long n, float2* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = seeded_phasor(seed, first + i);
}

// audio/synthesis.py:85: angles = exp(2j * pi * np.random.rand(bins, T)).  The per-item functions keep the
// reference's use of numpy's GLOBAL random stream; the host only draws the uniforms (bin-major, like
// np.random.rand(bins, T)) and this kernel turns them into frame-major unit phasors -- the complex
// exponential and the transpose to the device layout are the expensive part on the host.
__global__ void phase_from_uniform_kernel(const double* __restrict__ u, long long n_frames, int n_bins,
                                          float2* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n_frames * n_bins; i += stride) {
    const long long t = i / n_bins;
    const int k = (int)(i - t * n_bins);
    double s, c;
    sincospi(2.0 * u[(long long)k * n_frames + t], &s, &c);
    out[i] = make_float2((float)c, (float)s);
  }
}

// tacotron/inference.py:94-101,175: normalised model output -> dB (inv_normalize_decibel) ->
// magnitude (decibel_to_magnitude) -> magnitude ** power, in float32 like the reference.
// mag ** power = 10 ** (dB * power / 20).  *flag is set when a dB value is below -100
// (the reference raises AssertionError, audio/conversion.py:47-49).
__global__ void denormalize_magnitude_kernel(const float* __restrict__ x, long long n, float ref_db,
                                             float range_db, float power, float* __restrict__ out,
                                             int* flag) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float c = power * 0.16609640474436813f;   // log2(10) / 20
  bool bad = false;
  for (; i < n; i += stride) {
    const float v = fminf(fmaxf(x[i], 0.0f), 1.0f);
    const float db = (v - 1.0f) * range_db + ref_db;
    bad |= db < -100.0f;
    out[i] = exp2f(db * c);
  }
  if (bad && flag) atomicOr(flag, 1);
}

// Magnitude half of librosa.core.phase_vocoder as audio/effects.py:77-80 uses it (the stretched
// spectrogram's phase is discarded by the np.abs that follows): output frame t sits at input position
// step = t * rate; |D| is interpolated linearly between frames floor(step) and floor(step) + 1 (frames
// past the end count as zero, like the two zero columns librosa pads).
__global__ void stretch_magnitude_kernel(const float2* __restrict__ spec, long long n_frames, int n_bins,
                                         double rate, long long n_out, float* __restrict__ mag_out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n_out * n_bins; i += stride) {
    const long long t = i / n_bins;
    const int k = (int)(i - t * n_bins);
    const double step = (double)t * rate;              // np.arange(0, T, rate)[t]
    const long long c0 = (long long)step;
    const double alpha = step - floor(step);           // np.mod(step, 1.0)
    float m0 = 0.0f, m1 = 0.0f;
    if (c0 < n_frames) { const float2 v = spec[c0 * n_bins + k]; m0 = hypotf(v.x, v.y); }
    if (c0 + 1 < n_frames) { const float2 v = spec[(c0 + 1) * n_bins + k]; m1 = hypotf(v.x, v.y); }
    mag_out[i] = (float)((1.0 - alpha) * (double)m0 + alpha * (double)m1);
  }
}

// librosa.feature.mfcc(S=mel_spec, n_mfcc) as called by audio/features.py:111: the orthonormal DCT-II
// basis of librosa.filters.dct(n_mfcc, n_mels) applied along the mel axis, in float64 like np.dot.
//   basis[0][m] = 1 / sqrt(n),  basis[c][m] = cos(c (2 m + 1) pi / (2 n)) sqrt(2 / n)
__global__ void dct_project_kernel(const double* __restrict__ mel, long long n_frames, int n_mels, int n_mfcc,
                                   double* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const double s0 = 1.0 / sqrt((double)n_mels), s1 = sqrt(2.0 / (double)n_mels);
  for (; i < n_frames * n_mfcc; i += stride) {
    const long long t = i / n_mfcc;
    const int c = (int)(i - t * n_mfcc);
    const double* row = mel + t * n_mels;
    double acc = 0.0;
    if (c == 0) {
      for (int m = 0; m < n_mels; ++m) acc += s0 * row[m];
    } else {
      for (int m = 0; m < n_mels; ++m)
        acc += (cospi((double)c * (double)(2 * m + 1) / (double)(2 * n_mels)) * s1) * row[m];
    }
    out[i] = acc;
  }
}

// Peak normalisation of the reconstructed utterances -- librosa.util.normalize(y, norm=np.inf) as
// save_wav(..., norm=True) applies it (audio/io.py:33-53, tacotron/inference.py:199): y / max|y|, left
// alone when the peak is below float32 tiny.  One CTA per utterance and pass (grid-stride over utterances).
__global__ void peak_normalize_kernel(float* __restrict__ wav, const long long* __restrict__ sample_off, int n_utts) {
  __shared__ float s_red[8];
  __shared__ float s_peak;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int u = blockIdx.x; u < n_utts; u += gridDim.x) {
    float* y = wav + sample_off[u];
    const long long n = sample_off[u + 1] - sample_off[u];
    float m = 0.0f;
    for (long long i = tid; i < n; i += blockDim.x) m = fmaxf(m, fabsf(y[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    if (tid == 0) {
      float p = 0.0f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) p = fmaxf(p, s_red[w]);
      s_peak = p;
    }
    __syncthreads();
    const float peak = s_peak;
    if (peak >= 1.17549435e-38f) {
      for (long long i = tid; i < n; i += blockDim.x) y[i] = y[i] / peak;   // true division, like numpy
    }
    __syncthreads();
  }
}

// 16-bit PCM -> float32 in [-1, 1) exactly as the host decode does (audio/io.py:30 via librosa.load:
// int16 / 32768): clips can be uploaded as they sit in the wav files, at half the bytes.
__global__ void pcm16_to_float_kernel(const short* __restrict__ pcm, long long n, float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (float)pcm[i] * (1.0f / 32768.0f);
}

__global__ void minmax_init_kernel(long long* mm, int n_clips) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_clips * 4) mm[i] = (i & 1) ? encode_ordered(-1e300) : encode_ordered(1e300);
}
__global__ void minmax_decode_kernel(long long* mm, int n_clips) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_clips * 4) {
    long long c = mm[i];
    c = c >= 0 ? c : (c ^ 0x7fffffffffffffffLL);
    reinterpret_cast<double*>(mm)[i] = __longlong_as_double(c);
  }
}

template <typename T, typename G, int W>
int run_features(const sstts_feat_plan* P, const float* wav, const sstts_feat_outputs* O,
                 cudaStream_t st) {
  const FeatPlanHost& H = P->host;
  FeatArgs<T> A;
  A.wav = wav;
  A.sample_off = P->d_sample_off; A.sample_len = P->d_sample_len;
  A.frame_off = P->d_frame_off; A.row_off = P->d_row_off;
  A.tiles = P->d_tiles; A.n_tiles = (int)H.tiles.size();
  const StaticTables& S = *P->st;
  A.tab = tables_view<T>(S.tab);
  A.mel_ptr = S.d_mel_ptr; A.mel_k0 = S.d_mel_k0; A.mel_w = reinterpret_cast<const T*>(S.d_mel_w);
  A.n_mels = P->cfg.n_mels;
  A.mel_nnz = (int)S.mel.w.size();
  A.spec_out = reinterpret_cast<float2*>(O->spec_dev);
  A.lin_out = O->lin_db_dev;
  A.mel_out = O->mel_db_dev;
  A.melraw_out = O->mel_raw_dev;
  A.minmax_out = reinterpret_cast<long long*>(O->minmax_dev);
  // audio/conversion.py:78: the divisor abs(ref) + abs(max) is formed in Python floats.
  A.lin_ref_db = (float)O->lin_ref_db;
  A.lin_range_db = (float)(fabs(O->lin_ref_db) + fabs(O->lin_max_db));
  A.mel_ref_db = O->mel_ref_db;
  A.mel_range_db = fabs(O->mel_ref_db) + fabs(O->mel_max_db);
  A.mel_power = (float)(O->mel_power == 0.0 ? 1.0 : O->mel_power);
  A.normalize = O->normalize;
  A.win = H.win; A.hop = H.hop; A.span_max = H.span_max;
  A.n_fft = P->cfg.n_fft;
  if (A.n_mels < 1) { A.mel_out = nullptr; A.melraw_out = nullptr; }

  // the pre-calculation configuration (n_fft 2048: linear + mel dB) and the statistics configuration (native
  // n_fft 1024: any of per-clip extrema / linear dB / mel dB) run the fused float32-epilogue mode of the kernel
  const bool fused_ok = S.melp.ok && S.d_melp_w && !A.spec_out && !A.melraw_out && A.mel_power == 1.0f && !O->force_generic;
  const bool fast = fused_ok && (feat_native_1024(A.n_fft) ? (A.lin_out || A.mel_out || A.minmax_out)
                                                           : (A.n_fft == NFFT && A.lin_out && A.mel_out && !A.minmax_out));
  A.melp_w = S.d_melp_w;
  A.melp_slots = fast ? S.melp.n_slots : 0;
  A.melp_total = fast ? S.melp.total : 0;
  for (int j = 0; j < 4; ++j) { A.melp_len[j] = S.melp.len[j]; A.melp_woff[j] = S.melp.woff[j]; A.melp_mbase[j] = S.melp.mbase[j]; }

  const size_t smem = stft_feature_smem_bytes<T>(W, H.win, H.span_max, P->cfg.n_mels, (int)S.mel.w.size(),
                                                 A.melp_total, feat_plane_elems<G>());
  auto kernel = fast ? stft_feature_kernel<T, G, W, FeatMode::kDbFeatures>
                     : stft_feature_kernel<T, G, W, FeatMode::kGeneric>;
  int occ = 0, rc;
  if ((rc = configure_kernel(kernel, W * 32, smem, &occ))) return rc;
  const int n_sms = P->n_sms > 0 ? P->n_sms : 148;
  int grid = n_sms * occ;
  if (grid > A.n_tiles) grid = A.n_tiles;
  if (A.minmax_out) {
    const int n = H.n_clips * 4;
    minmax_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(A.minmax_out, H.n_clips);
    CU(cudaGetLastError());
  }
  kernel<<<grid, W * 32, smem, st>>>(A);
  CU(cudaGetLastError());
  if (A.minmax_out) {
    const int n = H.n_clips * 4;
    minmax_decode_kernel<<<(n + 255) / 256, 256, 0, st>>>(A.minmax_out, H.n_clips);
    CU(cudaGetLastError());
  }
  return 0;
}

}  // namespace

extern "C" {

static int griffin_lim_dispatch(const sstts_gl_plan* P, const float* mag_dev, const float* phase0_dev,
                                uint64_t seed, int64_t first, int n_iter, void* workspace_dev,
                                float* wav_out_dev, double* mse_frame_dev, void* stream) {
  if (!P || !mag_dev || !workspace_dev || n_iter < 0) return fail(SSTTS_ERR_INVALID, "bad griffin_lim arguments");
  if (P->host.total_samples > 0 && !wav_out_dev) return fail(SSTTS_ERR_INVALID, "wav_out_dev is NULL");
  if (mse_frame_dev && n_iter < 1) return fail(SSTTS_ERR_INVALID, "mse needs n_iter >= 1");
  if (reinterpret_cast<uintptr_t>(workspace_dev) & 15)
    return fail(SSTTS_ERR_INVALID, "workspace_dev must be 16-byte aligned (bulk copies read it in 16-byte units)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool model = is_model_geometry(P->cfg.n_fft, P->host.win, P->host.hop);
#define GL_CALL(T, G, W, B) run_griffin_lim<T, G, W, B>(P, mag_dev, phase0_dev, seed, first, n_iter, workspace_dev, \
                                                        wav_out_dev, mse_frame_dev, st)
  if (P->native1024) {
    const bool stats = is_stats_geometry(P->cfg.n_fft, P->host.win, P->host.hop);   // also audio/effects.py:71-86
    if (P->cfg.precision == SSTTS_F64)
      return stats ? GL_CALL(double, StatsGeom, kWarps, false) : GL_CALL(double, DynGeom1024, kWarps, false);
    return stats ? GL_CALL(float, StatsGeom, kGlWarps, false) : GL_CALL(float, DynGeom1024, kGlWarps, false);
  }
  if (P->cfg.precision == SSTTS_F64)
    return model ? GL_CALL(double, ModelGeom, kWarps, false) : GL_CALL(double, DynGeom, kWarps, false);
  if (gl_bulk_requested()) return model ? GL_CALL(float, ModelGeom, kGlWarps, true) : GL_CALL(float, DynGeom, kGlWarps, true);
  return model ? GL_CALL(float, ModelGeom, kGlWarps, false) : GL_CALL(float, DynGeom, kGlWarps, false);
#undef GL_CALL
}

int sstts_griffin_lim(const sstts_gl_plan* P, const float* mag_dev, const float* phase0_dev,
                      int n_iter, void* workspace_dev, float* wav_out_dev, double* mse_frame_dev,
                      void* stream) {
  if (!phase0_dev) return fail(SSTTS_ERR_INVALID, "phase0_dev is NULL (use sstts_griffin_lim_seeded)");
  return griffin_lim_dispatch(P, mag_dev, phase0_dev, 0, 0, n_iter, workspace_dev, wav_out_dev, mse_frame_dev, stream);
}

int sstts_griffin_lim_seeded(const sstts_gl_plan* P, const float* mag_dev, uint64_t seed, int64_t first_element,
                             int n_iter, void* workspace_dev, float* wav_out_dev, double* mse_frame_dev,
                             void* stream) {
  if (first_element < 0) return fail(SSTTS_ERR_INVALID, "first_element must be >= 0");
  return griffin_lim_dispatch(P, mag_dev, nullptr, seed, first_element, n_iter, workspace_dev, wav_out_dev,
                              mse_frame_dev, stream);
}

int sstts_random_phase_at(uint64_t seed, int64_t first, int64_t n, float* phase_dev, void* stream) {
  if (n < 0 || first < 0 || (n > 0 && !phase_dev)) return fail(SSTTS_ERR_INVALID, "bad random_phase arguments");
  if (n == 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  random_phase_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      seed, first, n, reinterpret_cast<float2*>(phase_dev));
  CU(cudaGetLastError());
  return 0;
}

int sstts_phase_from_uniform(const double* uniform_dev, int64_t n_frames, int n_bins, float* phase_dev, void* stream) {
  if (n_frames < 0 || n_bins < 1 || (n_frames > 0 && (!uniform_dev || !phase_dev)))
    return fail(SSTTS_ERR_INVALID, "bad phase_from_uniform arguments");
  if (n_frames == 0) return 0;
  long long blocks = (n_frames * n_bins + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  phase_from_uniform_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      uniform_dev, n_frames, n_bins, reinterpret_cast<float2*>(phase_dev));
  CU(cudaGetLastError());
  return 0;
}

int sstts_random_phase(uint64_t seed, int64_t n, float* phase_dev, void* stream) {
  return sstts_random_phase_at(seed, 0, n, phase_dev, stream);
}

int sstts_peak_normalize(const sstts_gl_plan* P, float* wav_dev, void* stream) {
  if (!P || (P->host.total_samples > 0 && !wav_dev)) return fail(SSTTS_ERR_INVALID, "bad peak_normalize arguments");
  if (P->host.total_samples == 0) return 0;
  int grid = P->host.n_utts < 148 * 4 ? P->host.n_utts : 148 * 4;
  peak_normalize_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(wav_dev, P->d_sample_off,
                                                                                   P->host.n_utts);
  CU(cudaGetLastError());
  return 0;
}

int sstts_pcm16_to_float(const int16_t* pcm_dev, int64_t n, float* out_dev, void* stream) {
  if (n < 0 || (n > 0 && (!pcm_dev || !out_dev))) return fail(SSTTS_ERR_INVALID, "bad pcm16_to_float arguments");
  if (n == 0) return 0;
  long long blocks = (n + 1023) / 1024;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pcm16_to_float_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const short*>(pcm_dev), n, out_dev);
  CU(cudaGetLastError());
  return 0;
}

int sstts_dct_project(const double* mel_dev, int64_t n_frames, int n_mels, int n_mfcc, double* out_dev,
                      void* stream) {
  if (!mel_dev || !out_dev || n_frames < 1 || n_mels < 1 || n_mfcc < 1 || n_mfcc > n_mels)
    return fail(SSTTS_ERR_INVALID, "bad dct_project arguments (need 1 <= n_mfcc <= n_mels)");
  long long blocks = (n_frames * n_mfcc + 127) / 128;
  if (blocks > 148 * 32) blocks = 148 * 32;
  dct_project_kernel<<<(int)blocks, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(mel_dev, n_frames, n_mels,
                                                                                      n_mfcc, out_dev);
  CU(cudaGetLastError());
  return 0;
}

int64_t sstts_stretch_frames(int64_t n_frames, double rate) {
  if (n_frames < 1 || !(rate > 0.0)) return 0;
  int64_t n = (int64_t)ceil((double)n_frames / rate);   // len(np.arange(0, n_frames, rate))
  while (n > 0 && (double)(n - 1) * rate >= (double)n_frames) --n;
  while ((double)n * rate < (double)n_frames) ++n;
  return n;
}

int sstts_stretch_magnitude(const float* spec_dev, int64_t n_frames, int n_bins, double rate,
                            float* mag_out_dev, void* stream) {
  if (!spec_dev || !mag_out_dev || n_frames < 1 || n_bins < 1 || !(rate > 0.0))
    return fail(SSTTS_ERR_INVALID, "bad stretch_magnitude arguments");
  const long long n_out = sstts_stretch_frames(n_frames, rate);
  long long blocks = (n_out * n_bins + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) return 0;
  stretch_magnitude_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(spec_dev), n_frames, n_bins, rate, n_out, mag_out_dev);
  CU(cudaGetLastError());
  return 0;
}

int sstts_denormalize_magnitude(const float* norm_dev, int64_t n, double ref_db, double max_db,
                                double power, float* mag_out_dev, int* flag_dev, void* stream) {
  if (n < 0 || (n > 0 && (!norm_dev || !mag_out_dev))) return fail(SSTTS_ERR_INVALID, "bad denormalize arguments");
  if (n == 0) return 0;
  long long blocks = (n + 1023) / 1024;
  if (blocks > 148 * 16) blocks = 148 * 16;
  denormalize_magnitude_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      norm_dev, n, (float)ref_db, (float)(fabs(ref_db) + fabs(max_db)), (float)power, mag_out_dev, flag_dev);
  CU(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------ features
int sstts_feat_plan_create_ranges(const sstts_stft_config* cfg, int n_clips, const int64_t* clip_start_host,
                                  const int64_t* clip_len_host, int reduction, sstts_feat_plan** plan_out) {
  if (plan_out) *plan_out = nullptr;
  int rc = check_config(cfg, true);
  if (rc) return rc;
  if (!plan_out || !clip_start_host || !clip_len_host || n_clips < 1) return fail(SSTTS_ERR_INVALID, "bad plan arguments");
  if (cfg->n_mels < 0 || cfg->n_mels > 1024) return fail(SSTTS_ERR_INVALID, "n_mels out of range");
  sstts_feat_plan* P = new (std::nothrow) sstts_feat_plan();
  if (!P) return fail(SSTTS_ERR_INVALID, "out of host memory");
  P->cfg = *cfg;
  std::string err;
  std::vector<long long> cs(clip_start_host, clip_start_host + n_clips), cl(clip_len_host, clip_len_host + n_clips);
  if (!build_feat_plan(n_clips, cs.data(), cl.data(), cfg->n_fft, cfg->win_length, cfg->hop_length, reduction,
                       P->host, err, cfg->precision == SSTTS_F64)) {
    delete P;
    return fail(SSTTS_ERR_INVALID, err);
  }
  cudaError_t e = cudaGetDevice(&P->device);
  if (e != cudaSuccess) { delete P; return cuda_fail(e, "cudaGetDevice"); }
  P->n_sms = sm_count();
  if (cfg->n_mels > 0 && cfg->sampling_rate < 1) { sstts_feat_plan_destroy(P); return fail(SSTTS_ERR_INVALID, "sampling_rate must be > 0"); }
  rc = get_static_tables(cfg, P->device, cfg->n_mels > 0, feat_native_1024(cfg->n_fft), &P->st);
  if (!rc) {
    const size_t o0 = P->blob.add(P->host.sample_off), o1 = P->blob.add(P->host.sample_len);
    const size_t o2 = P->blob.add(P->host.frame_off), o3 = P->blob.add(P->host.row_off);
    const size_t o4 = P->blob.add(P->host.tiles);
    rc = P->blob.commit();
    if (!rc && P->blob.dev) {
      P->d_sample_off = P->blob.at<long long>(o0); P->d_sample_len = P->blob.at<long long>(o1);
      P->d_frame_off = P->blob.at<long long>(o2); P->d_row_off = P->blob.at<long long>(o3);
      P->d_tiles = P->blob.at<FeatTile>(o4);
    }
  }
  if (rc) { sstts_feat_plan_destroy(P); return rc; }
  *plan_out = P;
  return 0;
}

int sstts_feat_plan_create(const sstts_stft_config* cfg, int n_clips, const int64_t* sample_off_host,
                           int reduction, sstts_feat_plan** plan_out) {
  if (plan_out) *plan_out = nullptr;
  if (!sample_off_host || n_clips < 1) return fail(SSTTS_ERR_INVALID, "bad plan arguments");
  std::vector<int64_t> cs(n_clips), cl(n_clips);
  for (int c = 0; c < n_clips; ++c) { cs[c] = sample_off_host[c]; cl[c] = sample_off_host[c + 1] - sample_off_host[c]; }
  return sstts_feat_plan_create_ranges(cfg, n_clips, cs.data(), cl.data(), reduction, plan_out);
}

int sstts_trim_bounds(const float* wav_dev, int n_clips, const int64_t* clip_start_dev,
                      const int64_t* clip_len_dev, double top_db, int frame_length, int hop_length,
                      int64_t* bounds_dev, void* stream) {
  if (!wav_dev || !clip_start_dev || !clip_len_dev || !bounds_dev || n_clips < 1 || frame_length < 2 ||
      hop_length < 1)
    return fail(SSTTS_ERR_INVALID, "bad trim arguments");
  int grid = n_clips < 148 * 8 ? n_clips : 148 * 8;
  trim_bounds_kernel<256><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      wav_dev, reinterpret_cast<const long long*>(clip_start_dev), reinterpret_cast<const long long*>(clip_len_dev),
      n_clips, frame_length, hop_length, top_db, reinterpret_cast<long long*>(bounds_dev));
  CU(cudaGetLastError());
  return 0;
}

void sstts_feat_plan_destroy(sstts_feat_plan* P) {
  if (!P) return;
  P->blob.release();
  delete P;
}

int64_t sstts_feat_total_frames(const sstts_feat_plan* P) { return P ? P->host.total_frames : 0; }
int64_t sstts_feat_total_rows(const sstts_feat_plan* P) { return P ? P->host.total_rows : 0; }
const int64_t* sstts_feat_frame_offsets(const sstts_feat_plan* P) {
  return P ? reinterpret_cast<const int64_t*>(P->host.frame_off.data()) : nullptr;
}
const int64_t* sstts_feat_row_offsets(const sstts_feat_plan* P) {
  return P ? reinterpret_cast<const int64_t*>(P->host.row_off.data()) : nullptr;
}
const double* sstts_feat_mel_basis(const sstts_feat_plan* P) {
  return (P && P->st && !P->st->mel_dense.empty()) ? P->st->mel_dense.data() : nullptr;
}

int sstts_stft_features(const sstts_feat_plan* P, const float* wav_dev, const sstts_feat_outputs* out,
                        void* stream) {
  if (!P || !wav_dev || !out) return fail(SSTTS_ERR_INVALID, "bad stft_features arguments");
  if ((out->mel_db_dev || out->mel_raw_dev) && P->cfg.n_mels < 1)
    return fail(SSTTS_ERR_INVALID, "mel outputs requested but the plan has no filterbank");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool model = is_model_geometry(P->cfg.n_fft, P->host.win, P->host.hop);
  const bool stats = is_stats_geometry(P->cfg.n_fft, P->host.win, P->host.hop);
  const bool native = feat_native_1024(P->cfg.n_fft);
  if (P->cfg.precision == SSTTS_F64) {
    if (model) return run_features<double, ModelGeom, kFeatWarpsF64>(P, wav_dev, out, st);
    if (stats) return run_features<double, StatsGeom, kFeatWarpsF64Native>(P, wav_dev, out, st);
    if (native) return run_features<double, DynGeom1024, kFeatWarpsF64Native>(P, wav_dev, out, st);
    return run_features<double, DynGeom, kFeatWarpsF64>(P, wav_dev, out, st);
  }
  if (model) return run_features<float, ModelGeom, kWarps>(P, wav_dev, out, st);
  if (stats) return run_features<float, StatsGeom, kWarps>(P, wav_dev, out, st);
  if (native) return run_features<float, DynGeom1024, kWarps>(P, wav_dev, out, st);
  return run_features<float, DynGeom, kWarps>(P, wav_dev, out, st);
}

}  // extern "C"
