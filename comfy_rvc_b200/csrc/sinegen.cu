// NSF harmonic source: SineGen + SourceModuleHnNSF (/root/reference/lib/infer_pack/models.py:361-411,
// 455-467) as scan kernels.  harmonic_num = 0 (models.py:489-491) so there is one channel.
//
// The reference runs this in fp32 on the CPU where torch.cumsum accumulates in fp64 and rounds
// each element to fp32.  The kernels therefore carry fp64 partial sums
// through a warp-shuffle / block / grid scan and round once per element:
//
//   pass A (one block per item):  rad[t] = fmod(f0[t]/sr, 1);  c[t] = (float)(sum_{<=t} rad) * upp
//   pass B1 (grid):  per 2048-sample chunk, fp64 sum of v[i] = rad_up[i] - wrap[i]
//                    where y[i] = lerp(c, align_corners=True), wrap[i] = fmod(y[i],1) < fmod(y[i-1],1)
//   pass B2 (one block per item):  exclusive fp64 scan of the chunk sums
//   pass B3 (grid):  in-chunk fp64 scan + offset -> phase -> sin -> uv/noise mix -> tanh(w x + b)
#include <math_constants.h>

#include "common.cuh"

namespace rvc {
namespace {

constexpr int kChunkThreads = 256;
constexpr int kPerThread = 8;
constexpr int kChunk = kChunkThreads * kPerThread;  // 2048 samples

__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  return v;
}

// ---- pass A ----------------------------------------------------------------------------------
__global__ void sine_frame_scan_kernel(const float* __restrict__ f0, float* __restrict__ rad, float* __restrict__ cum,
                                       int T, int upp, int sr) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthr = blockDim.x;
  const int per = (T + nthr - 1) / nthr;
  const int t0 = tid * per, t1 = min(T, t0 + per);
  const float* fb = f0 + (long long)b * T;
  float* rb = rad + (long long)b * T;
  float* cb = cum + (long long)b * T;
  __shared__ double wsum[32];
  double s = 0.0;
  for (int t = t0; t < t1; ++t) {
    float r = fmodf(__fdiv_rn(fb[t], (float)sr), 1.0f);   // models.py:377
    rb[t] = r;
    s += (double)r;
  }
  double incl = warp_incl_scan(s, lane);
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    double w = (lane < (nthr >> 5)) ? wsum[lane] : 0.0;
    double wi = warp_incl_scan(w, lane);
    wsum[lane] = wi - w;  // exclusive
  }
  __syncthreads();
  double run = wsum[warp] + (incl - s);
  for (int t = t0; t < t1; ++t) {
    run += (double)rb[t];
    cb[t] = __fmul_rn((float)run, (float)upp);            // models.py:383-384
  }
}

struct SineGeom {
  int T, upp;
  long long L;
  float lin_scale;   // (float)(T-1) / (L-1), align_corners=True
  float near_scale;  // (float)(1.0 / upp), nearest with scale_factor
};

__device__ __forceinline__ float interp_c(const float* __restrict__ cb, const SineGeom& g, long long i) {
  // F.interpolate(mode="linear", align_corners=True): models.py:385-390, App. C step 3
  const float src = __fmul_rn(g.lin_scale, (float)i);
  int i0 = (int)floorf(src);
  if (i0 > g.T - 1) i0 = g.T - 1;
  const int i1 = i0 + (i0 < g.T - 1 ? 1 : 0);
  float l1 = __fsub_rn(src, (float)i0);
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  const float l0 = __fsub_rn(1.f, l1);
  const float y = __fmaf_rn(l0, cb[i0], __fmul_rn(l1, cb[i1]));
  return fmodf(y, 1.0f);                                   // models.py:396
}

__device__ __forceinline__ int nearest_idx(const SineGeom& g, long long i) {
  int k = (int)floorf(__fmul_rn((float)i, g.near_scale));
  return k > g.T - 1 ? g.T - 1 : k;
}

// v[i] = (float)(rad_up[i] + shift[i]) for the PER consecutive samples starting at i0
__device__ __forceinline__ void chunk_values(const float* __restrict__ rb, const float* __restrict__ cb,
                                             const SineGeom& g, long long i0, float* v) {
  float prev = (i0 > 0 && i0 - 1 < g.L) ? interp_c(cb, g, i0 - 1) : 0.f;
#pragma unroll
  for (int q = 0; q < kPerThread; ++q) {
    const long long i = i0 + q;
    if (i < g.L) {
      const float cur = interp_c(cb, g, i);
      const float shift = (i > 0 && __fsub_rn(cur, prev) < 0.f) ? -1.0f : 0.0f;   // models.py:397-399
      v[q] = __fadd_rn(rb[nearest_idx(g, i)], shift);
      prev = cur;
    } else {
      v[q] = 0.f;
    }
  }
}

// ---- pass B1 ---------------------------------------------------------------------------------
__global__ void sine_chunk_sum_kernel(const float* __restrict__ rad, const float* __restrict__ cum,
                                      double* __restrict__ chunk_sum, SineGeom g, int nchunks) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* rb = rad + (long long)b * g.T;
  const float* cb = cum + (long long)b * g.T;
  float v[kPerThread];
  chunk_values(rb, cb, g, (long long)chunk * kChunk + (long long)tid * kPerThread, v);
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < kPerThread; ++q) s += (double)v[q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double ws[kChunkThreads / 32];
  if (lane == 0) ws[warp] = s;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kChunkThreads / 32; ++w) t += ws[w];
    chunk_sum[(long long)b * nchunks + chunk] = t;
  }
}

// ---- pass B2 ---------------------------------------------------------------------------------
__global__ void sine_chunk_scan_kernel(double* __restrict__ chunk_sum, int nchunks) {
  const int b = blockIdx.x;
  double* cs = chunk_sum + (long long)b * nchunks;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int per = (nchunks + nthr - 1) / nthr;
  const int c0 = tid * per, c1 = min(nchunks, c0 + per);
  __shared__ double wsum[32];
  double s = 0.0;
  for (int c = c0; c < c1; ++c) s += cs[c];
  double incl = warp_incl_scan(s, lane);
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    double w = (lane < (nthr >> 5)) ? wsum[lane] : 0.0;
    double wi = warp_incl_scan(w, lane);
    wsum[lane] = wi - w;
  }
  __syncthreads();
  double run = wsum[warp] + (incl - s);
  for (int c = c0; c < c1; ++c) {
    double x = cs[c];
    cs[c] = run;  // exclusive prefix
    run += x;
  }
}

// ---- pass B3 ---------------------------------------------------------------------------------
__global__ void sine_emit_kernel(const float* __restrict__ f0, const float* __restrict__ rad,
                                 const float* __restrict__ cum, const double* __restrict__ chunk_off,
                                 const float* __restrict__ noise, float* __restrict__ har, SineGeom g, int nchunks,
                                 float lin_w, float lin_b) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* rb = rad + (long long)b * g.T;
  const float* cb = cum + (long long)b * g.T;
  const float* fb = f0 + (long long)b * g.T;
  const long long i0 = (long long)chunk * kChunk + (long long)tid * kPerThread;
  float v[kPerThread];
  chunk_values(rb, cb, g, i0, v);
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < kPerThread; ++q) s += (double)v[q];
  double incl = warp_incl_scan(s, lane);
  __shared__ double ws[kChunkThreads / 32];
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  double base = chunk_off[(long long)b * nchunks + chunk];
  for (int w = 0; w < warp; ++w) base += ws[w];
  double run = base + (incl - s);
  const float two_pi_step = (float)CUDART_PI;  // (x * 2) * np.pi in fp32: models.py:400-402
  const float amp_uv = 0.003f;                 // noise_std
  const float amp_un = __fdiv_rn(0.1f, 3.0f);  // sine_amp / 3   (models.py:408)
#pragma unroll
  for (int q = 0; q < kPerThread; ++q) {
    const long long i = i0 + q;
    if (i >= g.L) break;
    run += (double)v[q];
    const float ph = (float)run;
    const float sn = __fmul_rn(sinf(__fmul_rn(__fmul_rn(ph, 2.0f), two_pi_step)), 0.1f);
    const float uv = fb[nearest_idx(g, i)] > 0.f ? 1.f : 0.f;    // models.py:353-359,404-407
    const float namp = uv > 0.f ? amp_uv : amp_un;
    const float nz = __fmul_rn(namp, noise[(long long)b * g.L + i]);
    const float sw = __fadd_rn(__fmul_rn(sn, uv), nz);            // models.py:410
    har[(long long)b * g.L + i] = tanhf(__fmaf_rn(sw, lin_w, lin_b));  // models.py:466
  }
}

}  // namespace

size_t sine_scratch_bytes(int B, int T, int upp) {
  const long long L = (long long)T * upp;
  const long long nchunks = (L + kChunk - 1) / kChunk;
  size_t bytes = 0;
  bytes += sizeof(double) * (size_t)B * nchunks;           // chunk sums / offsets (first: 8B aligned)
  bytes += sizeof(float) * (size_t)B * T * 2;              // rad, cum
  return (bytes + 255) & ~(size_t)255;
}

cudaError_t launch_sine_source(const float* f0, const float* noise, float* har, int B, int T, int upp, int sr,
                               float lin_w, float lin_b, void* scratch, cudaStream_t st) {
  if (B <= 0 || T <= 0 || upp <= 0) return cudaErrorInvalidValue;
  SineGeom g;
  g.T = T; g.upp = upp; g.L = (long long)T * upp;
  g.lin_scale = g.L > 1 ? (float)(T - 1) / (float)(g.L - 1) : 0.f;
  g.near_scale = (float)(1.0 / (double)upp);
  const int nchunks = (int)((g.L + kChunk - 1) / kChunk);
  double* chunk_sum = reinterpret_cast<double*>(scratch);
  float* rad = reinterpret_cast<float*>(chunk_sum + (size_t)B * nchunks);
  float* cum = rad + (size_t)B * T;
  sine_frame_scan_kernel<<<B, 1024, 0, st>>>(f0, rad, cum, T, upp, sr);
  sine_chunk_sum_kernel<<<dim3(nchunks, B), kChunkThreads, 0, st>>>(rad, cum, chunk_sum, g, nchunks);
  sine_chunk_scan_kernel<<<B, 1024, 0, st>>>(chunk_sum, nchunks);
  sine_emit_kernel<<<dim3(nchunks, B), kChunkThreads, 0, st>>>(f0, rad, cum, chunk_sum, noise, har, g, nchunks, lin_w,
                                                               lin_b);
  launch_counter().n += 4;
  return cudaGetLastError();
}

}  // namespace rvc
