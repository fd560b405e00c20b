// Device-side pre/post steps of the segmented conversion driver (VC.vc / VC.pipeline):
//   prepare_feats : x2 nearest interpolation of the HuBERT frames + "protect" blend   (vc_infer_pipeline.py:77-95)
//   absmax        : max |x| over the concatenated song                                 (vc_infer_pipeline.py:188)
//   to_int16      : x * 32768 / (max / 0.99), C truncation                             (vc_infer_pipeline.py:188-189)
// All three are HBM-bound streaming kernels (one pass, 16-byte accesses); they exist so that a whole song is
// converted without a host round trip per segment (the reference syncs and gc's after every segment, :102-112).
#include <cuda_fp16.h>

#include "common.cuh"

namespace rvc {
namespace {

// out[t][c] = f[t/2][c] * p + f0[t/2][c] * (1 - p),  p = pitchf[t] < 1 ? protect : 1   (fp32, no contraction:
// the reference evaluates mul, mul, add as separate torch ops)
template <typename TIn>
__global__ void prepare_feats_kernel(const TIn* __restrict__ f, const TIn* __restrict__ f0, const float* __restrict__ pitchf,
                                     float* __restrict__ out, int T, int C, float protect, int use_protect) {
  const long long total = (long long)T * (C / 4);
  const int c4n = C / 4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(idx / c4n);
    const int c = (int)(idx - (long long)t * c4n) * 4;
    const long long src = (long long)(t >> 1) * C + c;
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = (float)f[src + i];
    if (use_protect) {
      const float pf = pitchf[t];
      // pitchff[pitchf > 0] = 1; pitchff[pitchf < 1] = protect  (second assignment wins for 0 < f0 < 1)
      float p = pf;
      if (pf > 0.f) p = 1.f;
      if (pf < 1.f) p = protect;
      const float q = __fsub_rn(1.f, p);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        b[i] = (float)f0[src + i];
        a[i] = __fadd_rn(__fmul_rn(a[i], p), __fmul_rn(b[i], q));
      }
    }
    *reinterpret_cast<float4*>(out + (long long)t * C + c) = make_float4(a[0], a[1], a[2], a[3]);
  }
}

__global__ void absmax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ out_bits) {
  float m = 0.f;
  const long long n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x4[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order like their bit patterns
  }
}

__device__ __forceinline__ short cvt_i16(float v, float audio_max) {
  // numpy: (x * 32768 / audio_max).astype(int16) -- two float32 roundings, then truncation toward zero
  return (short)__float2int_rz(__fdiv_rn(__fmul_rn(v, 32768.f), audio_max));
}

__global__ void to_int16_kernel(const float* __restrict__ x, long long n, const float* __restrict__ absmax, short* __restrict__ out) {
  const float audio_max = __fdiv_rn(*absmax, 0.99f);
  const long long n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x4[i];
    short4 o;
    o.x = cvt_i16(v.x, audio_max); o.y = cvt_i16(v.y, audio_max); o.z = cvt_i16(v.z, audio_max); o.w = cvt_i16(v.w, audio_max);
    *reinterpret_cast<short4*>(out + i * 4) = o;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = cvt_i16(x[i], audio_max);
}

// Quiet-point search of VC.pipeline (vc_infer_pipeline.py:127-135) for ONE centre: candidates j in [lo, hi), value
// |audio_pad[j] + ... + audio_pad[j + window - 1]| with the sum accumulated left to right in double from +0.0 -- the
// reference's `audio_sum += audio_pad[i : i - window]` order per element, so the sums are bit-identical to numpy's.
// Block b scans its contiguous share of the range and writes its first minimum; the (few) block results are reduced by
// the caller in ascending block order with a strict '<' (first minimum overall, like np.where(seg == seg.min())[0][0]).
__global__ void quiet_point_kernel(const double* __restrict__ a, long long lo, long long hi, int window, double* __restrict__ best_v,
                                   long long* __restrict__ best_j) {
  const long long per = (hi - lo + gridDim.x - 1) / gridDim.x;
  const long long j0 = lo + (long long)blockIdx.x * per;
  const long long j1 = j0 + per < hi ? j0 + per : hi;
  double bv = INFINITY;
  long long bj = -1;
  for (long long j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
    double acc = 0.0;
    for (int i = 0; i < window; ++i) acc = __dadd_rn(acc, a[j + i]);
    const double v = fabs(acc);
    if (v < bv) { bv = v; bj = j; }
  }
  __shared__ double sv[256];
  __shared__ long long sj[256];
  sv[threadIdx.x] = bv; sj[threadIdx.x] = bj;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const double v2 = sv[threadIdx.x + o];
      const long long j2 = sj[threadIdx.x + o];
      const double v1 = sv[threadIdx.x];
      const long long jj = sj[threadIdx.x];
      if (j2 >= 0 && (jj < 0 || v2 < v1 || (v2 == v1 && j2 < jj))) { sv[threadIdx.x] = v2; sj[threadIdx.x] = j2; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { best_v[blockIdx.x] = sv[0]; best_j[blockIdx.x] = sj[0]; }
}

inline unsigned stream_grid(long long work, int threads) {
  long long blocks = (work + threads - 1) / threads;
  const long long cap = 148 * 8;
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

cudaError_t launch_prepare_feats(const void* f, const void* f0, int dtype, const float* pitchf, float* out, int F, int T, int C,
                                 float protect, int use_protect, cudaStream_t st) {
  if (T <= 0 || C % 4 != 0 || T > 2 * F || (use_protect && (!f0 || !pitchf))) return cudaErrorInvalidValue;
  const unsigned grid = stream_grid((long long)T * (C / 4), 256);
  if (dtype == 0)
    prepare_feats_kernel<float><<<grid, 256, 0, st>>>((const float*)f, (const float*)f0, pitchf, out, T, C, protect, use_protect);
  else if (dtype == 1)
    prepare_feats_kernel<__half><<<grid, 256, 0, st>>>((const __half*)f, (const __half*)f0, pitchf, out, T, C, protect, use_protect);
  else
    return cudaErrorInvalidValue;
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_absmax(const float* x, long long n, float* out, int reset, cudaStream_t st) {
  if (n < 0 || (reinterpret_cast<uintptr_t>(x) & 15)) return cudaErrorInvalidValue;
  if (reset) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), st);
    if (e != cudaSuccess) return e;
  }
  if (n == 0) return cudaSuccess;
  absmax_kernel<<<stream_grid(n / 4 + 1, 256), 256, 0, st>>>(x, n, reinterpret_cast<unsigned*>(out));
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_quiet_point(const double* audio_pad, long long lo, long long hi, int window, double* best_v, long long* best_j,
                               int n_blocks, cudaStream_t st) {
  if (!audio_pad || !best_v || !best_j || hi <= lo || window < 1 || n_blocks < 1 || n_blocks > 65535) return cudaErrorInvalidValue;
  quiet_point_kernel<<<n_blocks, 256, 0, st>>>(audio_pad, lo, hi, window, best_v, best_j);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_to_int16(const float* x, long long n, const float* absmax, short* out, cudaStream_t st) {
  if (n < 0 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 7)) return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  to_int16_kernel<<<stream_grid(n / 4 + 1, 256), 256, 0, st>>>(x, n, absmax, out);
  launch_counter().n++;
  return cudaGetLastError();
}

}  // namespace rvc
