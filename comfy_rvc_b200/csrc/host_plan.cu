// Host-side (CPU) piece of the song-level driver: the quiet-point search of VC.pipeline
// (/root/reference/vc_infer_pipeline.py:127-135).  The reference accumulates `audio_sum += audio_pad[i : i - window]`
// for i in range(window) over the WHOLE song (160 float64 passes over 9.6 M samples for a 10 min song: 1.8 s of numpy)
// and then looks at 2 * t_query samples around each centre.  Every element's sum is independent of the others, so only
// the examined elements are computed here, each with the reference's own accumulation order (left to right in double,
// starting from +0.0), several elements at a time in registers and the ranges split over threads: bit-identical sums,
// same first-minimum index.
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/rvcb200.h"

namespace {

struct Best {
  double v;
  int64_t j;
};

// first j in [j0, j1) minimising |sum_{i < window} a[j + i]| (sum accumulated left to right from +0.0)
Best scan_range(const double* a, int64_t j0, int64_t j1, int window) {
  Best best{INFINITY, -1};
  constexpr int U = 8;                                   // elements in flight: independent accumulators -> SIMD
  int64_t j = j0;
  for (; j + U <= j1; j += U) {
    double acc[U];
    for (int u = 0; u < U; ++u) acc[u] = 0.0;
    const double* p = a + j;
    for (int i = 0; i < window; ++i)
      for (int u = 0; u < U; ++u) acc[u] += p[i + u];
    for (int u = 0; u < U; ++u) {
      const double v = std::fabs(acc[u]);
      if (v < best.v) { best.v = v; best.j = j + u; }
    }
  }
  for (; j < j1; ++j) {
    double acc = 0.0;
    for (int i = 0; i < window; ++i) acc += a[j + i];
    const double v = std::fabs(acc);
    if (v < best.v) { best.v = v; best.j = j; }
  }
  return best;
}

}  // namespace

extern "C" int64_t rvcb200_host_quiet_point(const double* audio_pad, int64_t lo, int64_t hi, int32_t window, int32_t n_threads) {
  if (!audio_pad || hi <= lo || window < 1) return -1;
  const int64_t n = hi - lo;
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  if (n < 4096 * (int64_t)nt) nt = (int)(n / 4096 > 0 ? n / 4096 : 1);
  std::vector<Best> res((size_t)nt, Best{INFINITY, -1});
  std::vector<std::thread> th;
  const int64_t step = (n + nt - 1) / nt;
  for (int t = 1; t < nt; ++t) {
    const int64_t a0 = lo + t * step, a1 = a0 + step < hi ? a0 + step : hi;
    if (a0 >= hi) break;
    th.emplace_back([&res, audio_pad, a0, a1, window, t] { res[(size_t)t] = scan_range(audio_pad, a0, a1, window); });
  }
  res[0] = scan_range(audio_pad, lo, lo + step < hi ? lo + step : hi, window);
  for (auto& x : th) x.join();
  Best best{INFINITY, -1};
  for (const Best& b : res)                              // ranges are in ascending order: strict '<' keeps the first minimum
    if (b.j >= 0 && b.v < best.v) best = b;
  return best.j < 0 ? -1 : best.j - lo;
}
