// AddressSanitizer run of the emulated kernels (compute-sanitizer is unavailable on the GPU pool):
//   g++ -O1 -g -std=c++20 -fsanitize=address -pthread -I. -I../../single_speaker_tts_b200/csrc \
//       asan_main.cpp -o asan_main && ./asan_main
// Every shared-memory carve-up and every global buffer is an exactly-sized heap block, so an
// out-of-bounds access of a kernel aborts here.
#include "emu_driver.cpp"

int main() {
  std::vector<long long> frames = {1, 2, 5, 9, 13, 30, 8, 16};
  std::vector<long long> fo(1, 0);
  for (long long t : frames) fo.push_back(fo.back() + t);
  const long long T = fo.back();
  std::vector<float> mag((size_t)T * 1025), phase((size_t)T * 1025 * 2);
  unsigned s = 1;
  for (auto& v : mag) { s = s * 1664525u + 1013904223u; v = (s >> 8) * (1.0f / 16777216.0f); }
  for (size_t i = 0; i < phase.size(); i += 2) {
    s = s * 1664525u + 1013904223u;
    const float a = 6.2831853f * (s >> 8) * (1.0f / 16777216.0f);
    phase[i] = cosf(a); phase[i + 1] = sinf(a);
  }
  long long n_out = 0;
  for (long long t : frames) n_out += 275 * (t - 1);
  std::vector<float> wav((size_t)n_out);
  std::vector<double> mse((size_t)T);
  for (int prec = 0; prec < 2; ++prec) {
    if (emu_griffin_lim(1102, 275, prec, (int)frames.size(), fo.data(), mag.data(), phase.data(), 2, wav.data(),
                        mse.data(), 3, 2048)) return 1;
    if (emu_griffin_lim(1024, 256, prec, (int)frames.size(), fo.data(), mag.data(), phase.data(), 1, wav.data(),
                        nullptr, 2, 2048)) return 1;
  }
  // the bulk-copy staging variant of the float32 iteration kernel (cp.async.bulk emulated as memcpy: the 16-byte
  // rounded source ranges must stay inside the workspace, the destinations inside the shared-memory carve-up)
  setenv("SSTTS_GL_STAGING", "bulk", 1);
  if (emu_griffin_lim(1102, 275, 0, (int)frames.size(), fo.data(), mag.data(), phase.data(), 3, wav.data(),
                      mse.data(), 3, 2048)) return 1;
  if (emu_griffin_lim(1024, 256, 0, (int)frames.size(), fo.data(), mag.data(), phase.data(), 2, wav.data(),
                      nullptr, 2, 2048)) return 1;
  unsetenv("SSTTS_GL_STAGING");
  // shorter transforms: n_fft 1024 natively (two frames per warp; compile-time and run-time geometry) and, with
  // SSTTS_GL_NATIVE1024=0, embedded in the 2048-point kernels like n_fft 512: exactly-sized (sum T, n_fft/2 + 1) inputs
  for (int cfgi = 0; cfgi < 4; ++cfgi) {
    if (cfgi == 3) setenv("SSTTS_GL_NATIVE1024", "0", 1);
    const int n_fft = cfgi == 1 ? 512 : 1024, win = cfgi == 1 ? 400 : (cfgi == 2 ? 800 : 1024);
    const int hop = cfgi == 1 ? 100 : (cfgi == 2 ? 200 : 256), nb = n_fft / 2 + 1;
    std::vector<float> mag_s((size_t)T * nb), phase_s((size_t)T * nb * 2);
    for (size_t i = 0; i < mag_s.size(); ++i) mag_s[i] = mag[i];
    for (size_t i = 0; i < phase_s.size(); ++i) phase_s[i] = phase[i];
    long long n_s = 0;
    for (long long t : frames) n_s += hop * (t - 1);
    std::vector<float> wav_s((size_t)n_s);
    for (int prec = 0; prec < 2; ++prec)
      if (emu_griffin_lim(win, hop, prec, (int)frames.size(), fo.data(), mag_s.data(), phase_s.data(), 2, wav_s.data(),
                          mse.data(), 3, n_fft)) return 1;
  }
  unsetenv("SSTTS_GL_NATIVE1024");
  std::vector<long long> lens = {1, 2, 274, 275, 276, 1500, 5000, 9000, 12345, 7001};
  std::vector<long long> so(1, 0);
  for (long long n : lens) so.push_back(so.back() + n);
  std::vector<float> x((size_t)so.back());
  for (auto& v : x) { s = s * 1664525u + 1013904223u; v = (s >> 8) * (1.0f / 16777216.0f) - 0.5f; }
  for (int prec = 0; prec < 2; ++prec) {
    for (int cfgi = 0; cfgi < 2; ++cfgi) {
      const int n_fft = cfgi ? 1024 : 2048, win = cfgi ? 1024 : 1102, hop = cfgi ? 256 : 275, r = cfgi ? 1 : 5;
      const int nb = n_fft / 2 + 1;
      long long rows = 0;
      for (long long n : lens) { long long t = 1 + n / hop; rows += (t + r - 1) / r * r; }
      std::vector<float> spec((size_t)rows * nb * 2), lin((size_t)rows * nb), mel((size_t)rows * 80);
      std::vector<double> raw((size_t)rows * 80), mm(lens.size() * 4);
      if (emu_stft_features(n_fft, win, hop, prec, 22050, 80, 0.0, cfgi ? 11025.0 : 8000.0, (int)lens.size(), so.data(), r,
                            x.data(), spec.data(), lin.data(), mel.data(), raw.data(), mm.data(), 1, 35.66, 100.0, 6.02,
                            99.89, 1.0, 3, 0)) return 1;
      // fused dB-feature mode (lin + mel only): 16-byte staging next to the ends of the exactly-sized
      // wav buffer, padded mel table reads inside the per-warp plane
      if (!cfgi && emu_stft_features(n_fft, win, hop, prec, 22050, 80, 0.0, 8000.0, (int)lens.size(), so.data(), r,
                                     x.data(), nullptr, lin.data(), mel.data(), nullptr, nullptr, 1, 35.66, 100.0, 6.02,
                                     99.89, 1.0, 3, 1)) return 1;
      if (!cfgi && emu_stft_features(n_fft, win, hop, prec, 22050, 80, 0.0, 11025.0, (int)lens.size(), so.data(), r,
                                     x.data(), nullptr, lin.data(), mel.data(), nullptr, nullptr, 0, 0.0, 0.0, 0.0,
                                     0.0, 1.0, 2, 1)) return 1;
    }
  }
  // native n_fft 1024 path: fused statistics mode (extrema only, and with dB outputs), run-time geometry
  for (int prec = 0; prec < 2; ++prec) {
    long long rows = 0, rows2 = 0;
    for (long long n : lens) { rows += 1 + n / 256; long long t = 1 + n / 200; rows2 += (t + 4) / 5 * 5; }
    std::vector<float> lin((size_t)rows * 513), mel((size_t)rows * 80), lin2((size_t)rows2 * 513), mel2((size_t)rows2 * 80);
    std::vector<float> spec2((size_t)rows2 * 513 * 2);
    std::vector<double> mm(lens.size() * 4), raw2((size_t)rows2 * 80);
    if (emu_stft_features(1024, 1024, 256, prec, 22050, 80, 0.0, 11025.0, (int)lens.size(), so.data(), 1, x.data(),
                          nullptr, nullptr, nullptr, nullptr, mm.data(), 0, 0.0, 0.0, 0.0, 0.0, 1.0, 3, 1)) return 1;
    if (emu_stft_features(1024, 1024, 256, prec, 22050, 80, 0.0, 11025.0, (int)lens.size(), so.data(), 1, x.data(),
                          nullptr, lin.data(), mel.data(), nullptr, mm.data(), 1, 35.66, 100.0, 6.02, 99.89, 1.0, 2, 1)) return 1;
    if (emu_stft_features(1024, 800, 200, prec, 22050, 80, 0.0, 8000.0, (int)lens.size(), so.data(), 5, x.data(),
                          spec2.data(), lin2.data(), mel2.data(), raw2.data(), mm.data(), 1, 35.66, 100.0, 6.02, 99.89, 2.0, 3, 0)) return 1;
    if (emu_stft_features(1024, 800, 200, prec, 22050, 80, 0.0, 8000.0, (int)lens.size(), so.data(), 5, x.data(),
                          nullptr, lin2.data(), mel2.data(), nullptr, nullptr, 1, 35.66, 100.0, 6.02, 99.89, 1.0, 3, 1)) return 1;
  }
  std::vector<long long> cs(so.begin(), so.end() - 1), bounds(lens.size() * 2);
  emu_trim_bounds(x.data(), (int)lens.size(), cs.data(), lens.data(), 60.0, 2048, 512, bounds.data());
  printf("asan run ok\n");
  return 0;
}
