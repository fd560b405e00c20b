// First layer of the HuBERT / ContentVec feature encoder: Conv1d(1 -> C, k = 10, stride 5, no bias)
// -> GroupNorm(C groups over C channels = per-channel normalisation over time, affine) -> GELU, as HuggingFace
// `HubertGroupNormConvLayer` computes it (transformers/models/hubert/modeling_hubert.py; reached from
// /root/reference/lib/infer_pack/loaders.py:52-61).  With one input channel the "convolution" is 10 FMAs per output, so it
// is recomputed instead of stored: pass 1 accumulates per-channel sum / sum of squares over time (float per 256-frame
// chunk, double across chunks), pass 2 recomputes, normalises, applies GELU and writes the fp16 channels-last operand
// [L0][C] of the tensor-core conv stack (conv_tc.cu).  The other front-end layers are ordinary contractions on the
// generic tcgen05 kernel (comfy_rvc_b200/hubert.py).
#include <cuda_fp16.h>

#include "common.cuh"

namespace rvc {
namespace {

constexpr int kFrames = 256;      // output frames per block
constexpr int kMaxK = 16;

template <bool APPLY>
__global__ void hubert_conv0_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ gn_w,
                                    const float* __restrict__ gn_b, double* __restrict__ stats, __half* __restrict__ y,
                                    long long n, long long L0, int C, int K, int S, float eps, long long y_bstride) {
  extern __shared__ float xs[];                       // (kFrames - 1) * S + K input samples of this chunk
  const int b = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * kFrames;
  const int nt = (int)min((long long)kFrames, L0 - t0);
  const int nx = (nt - 1) * S + K;
  const float* xb = x + (long long)b * n + t0 * S;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = xb[i];
  __syncthreads();
  double* st = stats + (long long)b * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float wk[kMaxK];
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) wk[k] = k < K ? w[c * K + k] : 0.f;
    float mean = 0.f, scale = 0.f, shift = 0.f;
    if (APPLY) {
      const double m = st[c] / (double)L0;
      const double var = st[C + c] / (double)L0 - m * m;          // biased, like torch.nn.GroupNorm
      mean = (float)m;
      scale = (float)(1.0 / sqrt((var > 0.0 ? var : 0.0) + (double)eps)) * gn_w[c];
      shift = gn_b[c];
    }
    float s = 0.f, q = 0.f;
    for (int t = 0; t < nt; ++t) {
      const float* xp = xs + t * S;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxK; ++k)
        if (k < K) acc = fmaf(wk[k], xp[k], acc);
      if (APPLY) {
        const float v = (acc - mean) * scale + shift;
        y[(long long)b * y_bstride + (t0 + t) * C + c] = __float2half_rn(0.5f * v * (1.f + erff(v * 0.70710678118654752f)));
      } else {
        s += acc; q = fmaf(acc, acc, q);
      }
    }
    if (!APPLY) {
      atomicAdd(&st[c], (double)s);
      atomicAdd(&st[C + c], (double)q);
    }
  }
}

}  // namespace

cudaError_t launch_hubert_conv0(const float* x, const float* w, const float* gn_w, const float* gn_b, double* stats, void* y16,
                                int B, long long n, int C, int K, int S, float eps, long long y_bstride, cudaStream_t st) {
  if (!x || !w || !gn_w || !gn_b || !stats || !y16 || B <= 0 || K < 1 || K > kMaxK || S < 1 || n < K || C < 1) return cudaErrorInvalidValue;
  const long long L0 = (n - K) / S + 1;
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C * B, st);
  if (e != cudaSuccess) return e;
  const dim3 grid((unsigned)((L0 + kFrames - 1) / kFrames), (unsigned)B);
  const size_t smem = sizeof(float) * ((size_t)(kFrames - 1) * S + K);
  hubert_conv0_kernel<false><<<grid, 256, smem, st>>>(x, w, gn_w, gn_b, stats, reinterpret_cast<__half*>(y16), n, L0, C, K, S, eps, y_bstride);
  hubert_conv0_kernel<true><<<grid, 256, smem, st>>>(x, w, gn_w, gn_b, stats, reinterpret_cast<__half*>(y16), n, L0, C, K, S, eps, y_bstride);
  launch_counter().n += 2;
  return cudaGetLastError();
}

}  // namespace rvc
