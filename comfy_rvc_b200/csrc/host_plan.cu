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

// ---------------------------------------------------------------------------------------------------------------------
// Zero-phase high-pass of VC.pipeline: `audio = signal.filtfilt(bh, ah, audio)` (/root/reference/vc_infer_pipeline.py:122)
// with scipy's defaults (method "pad", padtype "odd", padlen = 3 * max(len(a), len(b))), i.e.
//     ext = odd_ext(x, padlen);  y = lfilter(b, a, ext, zi = zi * ext[0]);  y = lfilter(b, a, y[::-1], zi = zi * y[-1])[::-1]
// followed by the reflect padding of :141 (`np.pad(audio, (t_pad, t_pad), mode="reflect")`), written straight into the
// caller's (pinned) staging buffer.  lfilter is scipy's transposed direct form II, one sample at a time in double
// (scipy/signal/_lfilter.c.in, DOUBLE_filt): the same expressions in the same order, compiled without FMA contraction, so
// the result is bit-identical to scipy (tests/test_pipeline_host.py).  The two passes stay sequential: the order-5
// direct form with its poles clustered at |z| = 0.98-0.994 amplifies rounding noise to ~1e-8 absolute, so a
// chunk-parallel run (chunks started early from a zero state; tried: max |diff| 2.8e-8 on a 10 min song) converges to the
// sequential one only down to that noise floor, never bit for bit -- and 1e-8 is one float32 ulp of the samples HuBERT
// consumes.  What is gained is the constant: order known at compile time, state in registers, odd extension and reflect
// padding fused in (0.11 s instead of 0.25 s for scipy + np.pad on a 10 min song).
namespace {

constexpr int kMaxOrder = 16;

struct Df2t {
  int order;                       // len(b) - 1
  double b[kMaxOrder + 1], a[kMaxOrder + 1];
};

// lfilter(b, a, x, zi = z) over i in [0, N): x is read through `get` (forward: ext[i]; backward: y1[N-1-i]) and written
// through `put`.  M > 0: order known at compile time (fully unrolled; same expressions, same rounding), M = 0: any order.
template <int M, typename Get, typename Put>
void run_filter(const Df2t& f, Get get, Put put, int64_t N, double* z) {
  if constexpr (M > 0) {
    double zz[M], bb[M + 1], aa[M + 1];
    for (int k = 0; k < M; ++k) zz[k] = z[k];
    for (int k = 0; k <= M; ++k) { bb[k] = f.b[k]; aa[k] = f.a[k]; }
    for (int64_t i = 0; i < N; ++i) {
      const double xn = get(i);
      const double yn = zz[0] + bb[0] * xn;
#pragma GCC unroll 16
      for (int k = 0; k < M - 1; ++k) zz[k] = zz[k + 1] + xn * bb[k + 1] - yn * aa[k + 1];
      zz[M - 1] = xn * bb[M] - yn * aa[M];
      put(i, yn);
    }
  } else {
    const int m = f.order;
    for (int64_t i = 0; i < N; ++i) {
      const double xn = get(i);
      const double yn = z[0] + f.b[0] * xn;
      for (int k = 0; k < m - 1; ++k) z[k] = z[k + 1] + xn * f.b[k + 1] - yn * f.a[k + 1];
      z[m - 1] = xn * f.b[m] - yn * f.a[m];
      put(i, yn);
    }
  }
}

template <typename Get, typename Put>
void run_pass(const Df2t& f, Get get, Put put, int64_t N, const double* zi, double x0) {
  double z[kMaxOrder];
  for (int k = 0; k < f.order; ++k) z[k] = zi[k] * x0;            // lfilter(..., zi = zi * x[0])
  if (f.order == 5) run_filter<5>(f, get, put, N, z);              // the pipeline's Butterworth high-pass
  else run_filter<0>(f, get, put, N, z);
}

}  // namespace

extern "C" int rvcb200_host_filtfilt_pad(const double* x, int64_t n, const double* b, const double* a, const double* zi,
                                         int32_t order, int64_t pad, double* out, double* scratch) {
  if (!x || !b || !a || !zi || !out || !scratch || order < 1 || order > kMaxOrder || pad < 0) return 1;
  const int64_t padlen = 3 * (int64_t)(order + 1);
  if (n <= padlen || n <= pad) return 1;                           // scipy raises for n <= padlen; np.pad "reflect" needs n > pad
  Df2t f;
  f.order = order;
  for (int k = 0; k <= order; ++k) { f.b[k] = b[k] / a[0]; f.a[k] = a[k] / a[0]; }
  const int64_t N = n + 2 * padlen;
  // odd extension, read on the fly: ext[i] = 2 x[0] - x[padlen - i] | x[i - padlen] | 2 x[n-1] - x[n - 2 - (i - padlen - n)]
  auto ext = [x, n, padlen](int64_t i) -> double {
    if (i < padlen) return 2 * x[0] - x[padlen - i];
    if (i < padlen + n) return x[i - padlen];
    return 2 * x[n - 1] - x[n - 2 - (i - padlen - n)];
  };
  double* y1 = scratch;                                            // [N] forward result
  run_pass(f, ext, [y1](int64_t i, double v) { y1[i] = v; }, N, zi, ext(0));
  // backward pass over y1 reversed; its output index i is position N-1-i of the final signal: keep [padlen, padlen + n)
  double* dst = out + pad;
  run_pass(f, [y1, N](int64_t i) { return y1[N - 1 - i]; },
           [dst, N, padlen, n](int64_t i, double v) {
             const int64_t p = N - 1 - i - padlen;
             if (p >= 0 && p < n) dst[p] = v;
           },
           N, zi, y1[N - 1]);
  for (int64_t k = 0; k < pad; ++k) {                              // np.pad(..., mode="reflect")
    out[pad - 1 - k] = dst[k + 1];
    out[pad + n + k] = dst[n - 2 - k];
  }
  return 0;
}
