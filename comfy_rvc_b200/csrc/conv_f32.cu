// fp32 CUDA-core implicit-GEMM 1-D convolution, channels-last (the int16 +-1 LSB parity path).
//
//   y[b][j*os+g][co] = epi( bias[co] + sum_{tap<ntaps} sum_{ci<Cin}
//                           lrelu_s(x[b][j + g_off[g] + tap*dil][ci]) * w[g][tap][ci][co] )
//
// One kernel covers every dense contraction of the path: Conv1d with any kernel/dilation
// (/root/reference/lib/infer_pack/modules.py:295-308 ResBlock1, attentions.py:387-395 FFN,
// modules.py:192 WN in_layer, models.py:545 conv_pre), all 1x1 convs / Linear
// (attentions.py:213-219, models.py:92-104, modules.py:438-440,202), and ConvTranspose1d
// (models.py:551) as `G = stride` phase groups of ceil(k/stride)-tap convolutions whose
// outputs interleave.
//
// Tiling: time is the GEMM M axis, C_out the N axis, (tap, ci) the K axis.  A CTA of 128
// threads owns BM x BN outputs (128x64, 256x32 or 512x16), each thread an 8x8 register tile whose
// rows are interleaved (row = i*TMT + tm) so that the per-tap shifted reads of the staged
// input are conflict-free scalar LDS and the weight reads are broadcast LDS.128.  The input
// rows (BM + halo, 8 channels per chunk) are staged ONCE per channel chunk and reused by
// every tap; weights for the chunk arrive by cp.async; both are double-buffered.
#include "common.cuh"

namespace rvc {

namespace {

constexpr int kThreads = 128;
constexpr int BK = 8;          // input channels per chunk
constexpr int kMaxHalo = 50;   // (ntaps-1)*dil <= 50  (k=11, d=5)

template <int BN>
struct Tile {
  static constexpr int TN = BN / 8;            // threads along N
  static constexpr int TMT = kThreads / TN;    // threads along M
  static constexpr int BM = TMT * 8;           // rows per CTA
  static constexpr int NA = ((BM + kMaxHalo) * 2 + kThreads - 1) / kThreads;  // float4 prefetch regs
};

__host__ __device__ inline int lda_for(int nrows) { return ((nrows + 7) / 8) * 8 + 4; }  // == 4 (mod 8)

template <int BN>
__global__ void __launch_bounds__(kThreads) conv_f32_kernel(const ConvDesc p) {
  using TL = Tile<BN>;
  constexpr int TN = TL::TN, TMT = TL::TMT, BM = TL::BM, NA = TL::NA;

  const int tid = threadIdx.x;
  const int tn = tid % TN, tm = tid / TN;
  const int b = blockIdx.z / p.G, g = blockIdx.z % p.G;
  const int j0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int halo = (p.ntaps - 1) * p.dil;
  const int nrows = BM + halo;
  const int lda = lda_for(nrows);

  extern __shared__ __align__(16) float smem[];
  float* A_s = smem;                                   // [2][BK][lda]
  float* B_s = smem + 2 * BK * lda;                    // [2][ntaps][BK][BN]
  const int bstage = p.ntaps * BK * BN;

  const float* xb = p.x + (long long)b * p.x_bstride;
  const int row0 = j0 + p.g_off[g];
  int lin = p.L_in;
  if (p.in_len) lin = min(lin, p.in_len[b]);
  const float* wg = p.w + (long long)g * p.ntaps * p.Cin * p.Cout + n0;
  const float slope = p.in_slope;

  float4 areg[NA];
  const int nvec = nrows * 2;

  auto load_a = [&](int ci0) {
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      int idx = tid + q * kThreads;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < nvec) {
        int r = idx >> 1, h = idx & 1;
        int row = row0 + r;
        if (row >= 0 && row < lin) {
          v = __ldg(reinterpret_cast<const float4*>(xb + (long long)row * p.ldx + ci0 + h * 4));
          v.x = lrelu(v.x, slope); v.y = lrelu(v.y, slope); v.z = lrelu(v.z, slope); v.w = lrelu(v.w, slope);
        }
      }
      areg[q] = v;
    }
  };
  auto store_a = [&](int buf) {
    float* As = A_s + buf * BK * lda;
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      int idx = tid + q * kThreads;
      if (idx < nvec) {
        int r = idx >> 1, h = idx & 1;
        float* d = As + (h * 4) * lda + r;
        d[0] = areg[q].x; d[lda] = areg[q].y; d[2 * lda] = areg[q].z; d[3 * lda] = areg[q].w;
      }
    }
  };
  auto load_b = [&](int ci0, int buf) {
    float* Bs = B_s + buf * bstage;
    const int nv = p.ntaps * BK * (BN / 4);
    for (int idx = tid; idx < nv; idx += kThreads) {
      int c4 = idx % (BN / 4);
      int rc = idx / (BN / 4);           // tap*BK + c
      int tap = rc / BK, c = rc % BK;
      const float* src = wg + ((long long)tap * p.Cin + ci0 + c) * p.Cout + c4 * 4;
      cp_async16(Bs + rc * BN + c4 * 4, src);
    }
    cp_async_commit();
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nch = p.Cin / BK;
  load_a(0);
  load_b(0, 0);
  store_a(0);
  cp_async_wait_all();
  __syncthreads();

  const int co_a = tn * 4, co_b = BN / 2 + tn * 4;
  int buf = 0;
  for (int ch = 0; ch < nch; ++ch) {
    if (ch + 1 < nch) {
      load_a((ch + 1) * BK);
      load_b((ch + 1) * BK, buf ^ 1);
    }
    const float* As = A_s + buf * BK * lda + tm;
    const float* Bs = B_s + buf * bstage;
    for (int tap = 0; tap < p.ntaps; ++tap) {
      const float* At = As + tap * p.dil;
      const float* Bt = Bs + tap * BK * BN;
#pragma unroll
      for (int c = 0; c < BK; ++c) {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = At[c * lda + i * TMT];
        const float4 b0 = *reinterpret_cast<const float4*>(Bt + c * BN + co_a);
        const float4 b1 = *reinterpret_cast<const float4*>(Bt + c * BN + co_b);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i][0] = fmaf(a[i], b0.x, acc[i][0]); acc[i][1] = fmaf(a[i], b0.y, acc[i][1]);
          acc[i][2] = fmaf(a[i], b0.z, acc[i][2]); acc[i][3] = fmaf(a[i], b0.w, acc[i][3]);
          acc[i][4] = fmaf(a[i], b1.x, acc[i][4]); acc[i][5] = fmaf(a[i], b1.y, acc[i][5]);
          acc[i][6] = fmaf(a[i], b1.z, acc[i][6]); acc[i][7] = fmaf(a[i], b1.w, acc[i][7]);
        }
      }
    }
    if (ch + 1 < nch) {
      store_a(buf ^ 1);
      cp_async_wait_all();
    }
    __syncthreads();
    buf ^= 1;
  }

  // ---------------------------------- epilogue ----------------------------------------------
  const int olen = p.out_len ? p.out_len[b] : 0x7fffffff;
  float* yb = p.y + (long long)b * p.y_bstride;
  const float* rb = p.res ? p.res + (long long)b * p.res_bstride : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = j0 + i * TMT + tm;
    if (j >= p.Lj) continue;
    const long long orow = (long long)j * p.out_stride + g;
    const bool valid = orow < olen;
#pragma unroll
    for (int hsel = 0; hsel < 2; ++hsel) {
      const int co = n0 + (hsel ? co_b : co_a);
      float v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = acc[i][hsel * 4 + c];
      if (p.bias) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + co));
        v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
      }
      if (p.cond) {
        const float4 cv = __ldg(reinterpret_cast<const float4*>(p.cond + (long long)b * p.cond_bstride + co));
        v[0] += cv.x; v[1] += cv.y; v[2] += cv.z; v[3] += cv.w;
      }
      if (p.gather) {
        const long long gi = p.gidx[(long long)b * p.gidx_bstride + orow];
        const float4 gv = __ldg(reinterpret_cast<const float4*>(p.gather + gi * p.Cout + co));
        v[0] += gv.x; v[1] += gv.y; v[2] += gv.z; v[3] += gv.w;
      }
      if (p.alpha != 1.f) {
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] *= p.alpha;
      }
      int nout = 4, oc = co;
      if (p.gate) {  // interleaved (tanh, sigmoid) pairs: commons.py:211-218
        v[0] = tanhf(v[0]) * sigmoidf_(v[1]);
        v[1] = tanhf(v[2]) * sigmoidf_(v[3]);
        nout = 2; oc = co >> 1;
      }
      if (p.mask_pre && !valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = 0.f;
      }
      if (rb) {
        const float* rp = rb + orow * p.ldr + oc;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nout) { float r = rp[c]; v[c] = (p.res_mode == 2) ? (r - v[c]) : (v[c] + r); }
      }
      if (p.relu) {
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = fmaxf(v[c], 0.f);
      } else if (p.out_slope != 1.f) {
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = lrelu(v[c], p.out_slope);
      }
      float* yp = yb + orow * p.ldy + oc;
      if (p.accum) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nout) v[c] = yp[c] + v[c];
      }
      if (p.div != 1.f) {
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = v[c] / p.div;
      }
      if (p.mask_post && !valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = 0.f;
      }
      if (nout == 4) {
        *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        *reinterpret_cast<float2*>(yp) = make_float2(v[0], v[1]);
      }
    }
  }
}

template <int BN>
cudaError_t launch_t(const ConvDesc& d, int B, cudaStream_t st) {
  using TL = Tile<BN>;
  const int halo = (d.ntaps - 1) * d.dil;
  const int lda = lda_for(TL::BM + halo);
  const size_t smem = sizeof(float) * (2 * BK * lda + 2 * (size_t)d.ntaps * BK * BN);
  static SmemOptIn opt;
  if (cudaError_t e = opt_in_smem(conv_f32_kernel<BN>, smem, opt)) return e;
  dim3 grid((d.Lj + TL::BM - 1) / TL::BM, d.Cout / BN, B * d.G);
  conv_f32_kernel<BN><<<grid, kThreads, smem, st>>>(d);
  launch_counter().n++;
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_conv_f32(const ConvDesc& d, int B, cudaStream_t st) {
  if (d.Cin % BK != 0 || d.Cout % 16 != 0 || d.ntaps < 1 || (d.ntaps - 1) * d.dil > kMaxHalo || d.G < 1 || d.G > 16 ||
      d.ldx % 4 != 0 || d.ldy % 2 != 0 || d.Lj <= 0 || B <= 0)
    return cudaErrorInvalidValue;
  if (d.Cout % 64 == 0) return launch_t<64>(d, B, st);
  if (d.Cout % 32 == 0) return launch_t<32>(d, B, st);
  return launch_t<16>(d, B, st);   // last stage of the 5-stage v1 ladders (32k, 48k): 32 -> 16 channels
}

}  // namespace rvc
