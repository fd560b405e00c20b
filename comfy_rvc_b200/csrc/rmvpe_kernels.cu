// RMVPE f0 estimator (/root/reference/lib/rmvpe.py): everything that is not a contraction.
// The DeepUnet's 3 x 3 convolutions, the transposed convolutions (as 2 x 2-tap phase GEMMs), the GRU input projection and the
// output Linear run on the generic tcgen05 kernel (conv_tc.cu, 2-D taps); this file holds
//   * the log-mel front end (`MelSpectrogram.forward`, rmvpe.py:489-556): reflect padding, periodic Hann window, a 1024-point
//     FFT per frame in shared memory (the reference convolves with the windowed DFT matrix, rmvpe.py:86-152 -- same numbers,
//     N log N instead of N^2), magnitude, the 128 triangular HTK filters (sparse: each covers a few dozen bins), log(clamp),
//     and the model's input BatchNorm (rmvpe.py:299) -> fp16 image [T][129][8] (pixel 128 of a line = zero pad, channel 0 = value);
//     frames past the end reproduce `mel2hidden`'s reflect padding of the frame axis (rmvpe.py:594-595);
//   * AvgPool2d(2, 2) (rmvpe.py:318) fp32 image -> fp16 operand of the next level;
//   * the pixel shuffle behind a transposed convolution run as a GEMM over (phase, channel) columns (rmvpe.py:355-366);
//   * the bidirectional GRU recurrence (rmvpe.py:217-229): one 8-CTA cluster per direction, W_hh resident in registers,
//     the hidden state exchanged through distributed shared memory (st.async + mbarrier complete_tx);
//   * sigmoid + `to_local_average_cents` + `decode` (rmvpe.py:610-615, 658-684) in float64 like the reference's numpy.
#include <cuda_fp16.h>

#include "common.cuh"

namespace rvc {
namespace {

constexpr int kNfft = 1024, kHop = 160, kMels = 128, kBins = 513;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- log-mel -----------------------------------------------
__global__ void __launch_bounds__(256)
rmvpe_logmel_kernel(const float* __restrict__ audio, long long n, const float* __restrict__ window, const float2* __restrict__ twiddle,
                    const float* __restrict__ mel_basis, const int2* __restrict__ mel_range, float bn_scale, float bn_shift,
                    float* __restrict__ mel_out, __half* __restrict__ img, int n_frames, int img_pitch, int img_c) {
  __shared__ float re[kNfft], im[kNfft];
  __shared__ float2 tw[kNfft / 2];
  __shared__ float mag[kBins + 3];
  const int tid = threadIdx.x;
  const int j = blockIdx.x;                                          // frame index on the (reflect-padded) frame axis
  const int src = j < n_frames ? j : 2 * n_frames - 2 - j;           // F.pad(mel, (0, padding), "reflect")
  for (int i = tid; i < kNfft / 2; i += 256) tw[i] = twiddle[i];
  for (int i = tid; i < kNfft; i += 256) {
    long long s = (long long)src * kHop - kNfft / 2 + i;             // F.pad(audio, 512, "reflect"), frame at hop 160
    if (s < 0) s = -s;
    if (s >= n) s = 2 * (n - 1) - s;
    const int r = (int)(__brev((unsigned)i) >> 22);                  // bit-reversed position (10 bits)
    re[r] = audio[s] * window[i];
    im[r] = 0.f;
  }
  __syncthreads();
#pragma unroll 1
  for (int lg = 1; lg <= 10; ++lg) {                                 // radix-2 decimation in time
    const int half = 1 << (lg - 1), step = kNfft >> lg;
    for (int b = tid; b < kNfft / 2; b += 256) {
      const int k = b & (half - 1);
      const int i0 = ((b >> (lg - 1)) << lg) + k, i1 = i0 + half;
      const float2 w = tw[k * step];                                 // exp(-2 pi i k / len)
      const float xr = re[i1] * w.x - im[i1] * w.y, xi = re[i1] * w.y + im[i1] * w.x;
      const float ar = re[i0], ai = im[i0];
      re[i1] = ar - xr; im[i1] = ai - xi;
      re[i0] = ar + xr; im[i0] = ai + xi;
    }
    __syncthreads();
  }
  for (int k = tid; k < kBins; k += 256) mag[k] = sqrtf(re[k] * re[k] + im[k] * im[k]);        // rmvpe.py:147
  __syncthreads();
  if (tid < kMels) {
    const int2 rg = mel_range[tid];
    const float* bw = mel_basis + (size_t)tid * kBins;
    float acc = 0.f;
    for (int k = rg.x; k < rg.y; ++k) acc = fmaf(bw[k], mag[k], acc);                           // rmvpe.py:550
    const float lm = logf(fmaxf(acc, 1e-5f));                                                   // rmvpe.py:553
    if (mel_out && j < n_frames) mel_out[(size_t)tid * n_frames + j] = lm;
    if (img) {
      const __half h = __float2half_rn(fmaf(lm, bn_scale, bn_shift));                           // Encoder.bn, rmvpe.py:299
      img[((size_t)j * img_pitch + tid) * img_c] = h;                                          // channel 0; the rest stays zero
    }
  }
}

// `mel2hidden` called with a caller-supplied log-mel (rmvpe.py:591-608): mel fp32 [128][n_frames] -> the same image as above
__global__ void rmvpe_mel_to_img_kernel(const float* __restrict__ mel, __half* __restrict__ img, int n_frames, int frames_out,
                                        float bn_scale, float bn_shift, int img_pitch, int img_c) {
  const long long total = (long long)frames_out * kMels;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx % kMels);
    const int j = (int)(idx / kMels);
    const int src = j < n_frames ? j : 2 * n_frames - 2 - j;
    const __half h = __float2half_rn(fmaf(mel[(size_t)m * n_frames + src], bn_scale, bn_shift));
    img[((size_t)j * img_pitch + m) * img_c] = h;
  }
}

// ---------------------------------------------------------------- pooling / shuffle / pack ------------------------------
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// x fp32 [2 H2][p_in pixels][C] (ldx floats per pixel) -> y fp16 [H2][p_out pixels][C]; pixels [W2, p_out) of a line of y (its
// zero padding) are written as zero
__global__ void rmvpe_pool_kernel(const float* __restrict__ x, int ldx, __half* __restrict__ y, int H2, int W2, int C, int p_in,
                                  int p_out) {
  const int c8n = C / 8;
  const long long total = (long long)H2 * p_out * c8n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % c8n);
    const long long px = idx / c8n;
    const int ox = (int)(px % p_out);
    const long long oy = px / p_out;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (ox < W2) {
      const float* p00 = x + ((size_t)(2 * oy) * p_in + 2 * ox) * ldx + c8 * 8;
      const float* p10 = p00 + (size_t)p_in * ldx;
      float v[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 a = *reinterpret_cast<const float4*>(p00 + h * 4), b = *reinterpret_cast<const float4*>(p00 + ldx + h * 4);
        const float4 c = *reinterpret_cast<const float4*>(p10 + h * 4), d = *reinterpret_cast<const float4*>(p10 + ldx + h * 4);
        v[h * 4 + 0] = (((a.x + b.x) + c.x) + d.x) * 0.25f;
        v[h * 4 + 1] = (((a.y + b.y) + c.y) + d.y) * 0.25f;
        v[h * 4 + 2] = (((a.z + b.z) + c.z) + d.z) * 0.25f;
        v[h * 4 + 3] = (((a.w + b.w) + c.w) + d.w) * 0.25f;
      }
      o = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
    }
    *reinterpret_cast<uint4*>(y + (size_t)px * C + c8 * 8) = o;
  }
}

// g fp16 [H][p_in pixels][4][Co] (phase = py * 2 + px) -> the up-sampled half of the decoder's concat buffer: output pixel
// (2y + py, 2x + px) lives in row (2y + py) * fp_out + (2x + px) / pack of `out` (row stride ld), columns ((2x + px) % pack) * Co ..
// (pack = pixels per 64-channel row of the wide levels, 1 elsewhere)
__global__ void rmvpe_shuffle_kernel(const __half* __restrict__ g, __half* __restrict__ out, int H, int W, int Co, int ld, int p_in,
                                     int fp_out, int pack) {
  const int c8n = Co / 8;
  const long long total = (long long)H * W * 4 * c8n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % c8n);
    long long r = idx / c8n;
    const int ph = (int)(r & 3); r >>= 2;
    const int x = (int)(r % W);
    const long long y = r / W;
    const uint4 v = *reinterpret_cast<const uint4*>(g + (((size_t)y * p_in + x) * 4 + ph) * Co + c8 * 8);
    const int ox = 2 * x + (ph & 1);
    const size_t orow = (size_t)(2 * y + (ph >> 1)) * fp_out + ox / pack;
    *reinterpret_cast<uint4*>(out + orow * ld + (size_t)(ox % pack) * Co + c8 * 8) = v;
  }
}

// cnn output fp32 [T][pitch pixels][ldc] (channels 0..2) -> GRU operand fp16 [T][3 W], column c * W + w (rmvpe.py:468 transpose + flatten)
__global__ void rmvpe_gru_pack_kernel(const float* __restrict__ y, int ldc, __half* __restrict__ x16, long long T, int W, int pitch) {
  const long long total = T * 3 * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(idx % W);
    const int c = (int)((idx / W) % 3);
    const long long t = idx / (3 * W);
    x16[idx] = __float2half_rn(y[((size_t)t * pitch + w) * ldc + c]);
  }
}

// ---------------------------------------------------------------- GRU recurrence ----------------------------------------
// torch.nn.GRU(384, 256, bidirectional), gates r | z | n (oracle/rmvpe_oracle.py `bigru`).  gi = W_ih x + b_ih comes from a
// tensor-core GEMM: [T][2 * 768] fp32, column dir * 768 + gate * 256 + unit.  Grid = 2 clusters of 8 CTAs (cluster = direction);
// CTA `rank` owns hidden units [32 rank, 32 rank + 32): thread (unit, slice s of 8) keeps the 3 x 32 weights of its unit's rows
// for columns {4 (8 i + s) .. + 3, i < 8} in registers.  Per step: 96 FMAs against the hidden state in shared memory, an 8-lane
// butterfly, the gate arithmetic (every lane of the 8 redundantly), then lane s stores the unit's new value into CTA s's copy of
// the state through distributed shared memory (see the exchange note below); both directions run concurrently.
constexpr int kGruH = 256, kGruCl = 8, kGruUnits = kGruH / kGruCl;

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// (plain try_wait, as for TMA / multicast data: with `.acquire.cluster` every step paid a CCTL.IVALL -- an L1 invalidate, 11 % of
//  the kernel's stall samples in profiles/r2_ncu_rmvpe.md; the mbarrier's complete_tx is what orders the st.async data)
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "GRU_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra GRU_WAIT_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// sigmoid / tanh from the hardware exponential (ex2.approx, 2 ulp): ~1e-7 absolute on the gates
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

// Exchange of the hidden state (second version; the first used one barrier.cluster per step and took 1.02 us per step,
// profiles/r2_rmvpe_bench_v1.jsonl): every new value travels as ONE `st.async` to each CTA of the cluster that both writes the
// float into that CTA's state buffer and counts 4 bytes on that CTA's mbarrier (complete_tx), so data and "it is there" are the
// same message; a CTA waits on its own mbarrier for the 1024 bytes of a step.  The state is double-buffered and the data
// dependency orders the reuse of a buffer: a CTA can only send step it + 1 values after it has received every step it value,
// which every sender emitted after its own reads of the buffer those values now overwrite.
__global__ void __cluster_dims__(kGruCl, 1, 1) __launch_bounds__(256, 1)
rmvpe_gru_kernel(const float* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                 __half* __restrict__ out16, float* __restrict__ out32, int T) {
  __shared__ __align__(16) float hbuf[2][kGruH];
  __shared__ __align__(8) unsigned long long bars[2];
  const int dir = blockIdx.x / kGruCl;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x, s = tid & 7, u = (int)rank * kGruUnits + (tid >> 3);
  float4 w[3][8];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      w[g][i] = *reinterpret_cast<const float4*>(w_hh + ((size_t)(dir * 3 + g) * kGruH + u) * kGruH + 4 * (8 * i + s));
  const float bh_r = b_hh[dir * 3 * kGruH + u], bh_z = b_hh[dir * 3 * kGruH + kGruH + u], bh_n = b_hh[dir * 3 * kGruH + 2 * kGruH + u];
  hbuf[0][tid] = 0.f;
  hbuf[1][tid] = 0.f;
  const uint32_t bar0 = smem_addr(&bars[0]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // buffer 1 receives the results of step 0, buffer 0 those of step 1
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8), "r"(kGruH * 4) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0), "r"(kGruH * 4) : "memory");
  }
  uint32_t remote_h, remote_bar;                                      // state buffer and barriers of cluster CTA `s`
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_h) : "r"(smem_addr(&hbuf[0][0])), "r"((uint32_t)s));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_bar) : "r"(bar0), "r"((uint32_t)s));
  __syncthreads();
  cluster_sync_all();                                                 // every CTA runs, has zeroed its state and armed its barriers
  const float* gp = gi + dir * 3 * kGruH + u;
  const size_t ldg = 2 * 3 * kGruH;
  const int dt = dir ? -1 : 1;
  int t = dir ? T - 1 : 0;
  // input gates of this step and the next one (two steps of load latency hidden)
  float gr = gp[(size_t)t * ldg], gz = gp[(size_t)t * ldg + kGruH], gn = gp[(size_t)t * ldg + 2 * kGruH];
  float gr1 = 0.f, gz1 = 0.f, gn1 = 0.f;
  if (T > 1) { const size_t o = (size_t)(t + dt) * ldg; gr1 = gp[o]; gz1 = gp[o + kGruH]; gn1 = gp[o + 2 * kGruH]; }
#pragma unroll 1
  for (int it = 0; it < T; ++it) {
    const int cur = it & 1;
    if (it > 0) {
      // step it - 1's 256 values have landed in hbuf[cur]: use number (it - 1) / 2 of bars[cur] -> parity ((it - 1) >> 1) & 1
      mbar_wait_parity(bar0 + 8 * cur, (uint32_t)(((it - 1) >> 1) & 1));
      if (tid == 0 && it + 2 < T)                                     // arm it for its next use (the results of step it + 1)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * cur), "r"(kGruH * 4) : "memory");
    }
    const float* hcur = hbuf[cur];
    const float4* hc = reinterpret_cast<const float4*>(hcur);
    float ar = 0.f, az = 0.f, an = 0.f, br = 0.f, bz = 0.f, bn = 0.f;  // two chains per gate
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const float4 h4 = hc[8 * i + s], k4 = hc[8 * (i + 1) + s];
      ar = fmaf(w[0][i].x, h4.x, ar); ar = fmaf(w[0][i].y, h4.y, ar); ar = fmaf(w[0][i].z, h4.z, ar); ar = fmaf(w[0][i].w, h4.w, ar);
      az = fmaf(w[1][i].x, h4.x, az); az = fmaf(w[1][i].y, h4.y, az); az = fmaf(w[1][i].z, h4.z, az); az = fmaf(w[1][i].w, h4.w, az);
      an = fmaf(w[2][i].x, h4.x, an); an = fmaf(w[2][i].y, h4.y, an); an = fmaf(w[2][i].z, h4.z, an); an = fmaf(w[2][i].w, h4.w, an);
      br = fmaf(w[0][i + 1].x, k4.x, br); br = fmaf(w[0][i + 1].y, k4.y, br); br = fmaf(w[0][i + 1].z, k4.z, br); br = fmaf(w[0][i + 1].w, k4.w, br);
      bz = fmaf(w[1][i + 1].x, k4.x, bz); bz = fmaf(w[1][i + 1].y, k4.y, bz); bz = fmaf(w[1][i + 1].z, k4.z, bz); bz = fmaf(w[1][i + 1].w, k4.w, bz);
      bn = fmaf(w[2][i + 1].x, k4.x, bn); bn = fmaf(w[2][i + 1].y, k4.y, bn); bn = fmaf(w[2][i + 1].z, k4.z, bn); bn = fmaf(w[2][i + 1].w, k4.w, bn);
    }
    ar += br; az += bz; an += bn;
    const float hp = hcur[u];
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      ar += __shfl_xor_sync(0xffffffffu, ar, o);
      az += __shfl_xor_sync(0xffffffffu, az, o);
      an += __shfl_xor_sync(0xffffffffu, an, o);
    }
    const float r = fast_sigmoid(gr + (ar + bh_r));
    const float z = fast_sigmoid(gz + (az + bh_z));
    const float nn = fast_tanh(gn + r * (an + bh_n));
    const float hn = (hp - nn) * z + nn;                              // ATen's gru cell: (h - n) * z + n
    if (it + 1 < T) {                                                 // nobody reads the last step's state: no traffic after exit
      const uint32_t dst = remote_h + (uint32_t)(((cur ^ 1) * kGruH + u) * 4);
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst), "r"(__float_as_uint(hn)),
                   "r"(remote_bar + 8 * (cur ^ 1))
                   : "memory");
    }
    if (s == 0) {
      out16[(size_t)t * (2 * kGruH) + dir * kGruH + u] = __float2half_rn(hn);
      if (out32) out32[(size_t)t * (2 * kGruH) + dir * kGruH + u] = hn;
    }
    t += dt;
    gr = gr1; gz = gz1; gn = gn1;
    if (it + 2 < T) { const size_t o = (size_t)(t + dt) * ldg; gr1 = gp[o]; gz1 = gp[o + kGruH]; gn1 = gp[o + 2 * kGruH]; }
  }
}

// ---------------------------------------------------------------- salience + decode -------------------------------------
constexpr int kClasses = 360, kDecWarps = 8;

// One warp per frame.  from_hidden = 0: in = logits [T][ld] -> hidden = sigmoid (rmvpe.py:455) and f0; 1: in = hidden [T][ld].
__global__ void __launch_bounds__(32 * kDecWarps)
rmvpe_decode_kernel(const float* __restrict__ in, int ld, int from_hidden, float* __restrict__ hidden, double* __restrict__ f0,
                    double* __restrict__ cents_out, int T, float thred) {
  __shared__ float sal[kDecWarps][kClasses + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * kDecWarps + warp;
  if (t >= T) return;
  float best = -1.f;
  int bi = 0;
  for (int c = lane; c < kClasses; c += 32) {
    const float x = in[(size_t)t * ld + c];
    const float v = from_hidden ? x : 1.f / (1.f + expf(-x));
    sal[warp][c] = v;
    if (hidden) hidden[(size_t)t * kClasses + c] = v;
    if (v > best) { best = v; bi = c; }                               // ascending c per lane: keeps the first maximum
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }  // np.argmax: first occurrence
  }
  __syncwarp();
  if (lane == 0) {
    // rmvpe.py:658-684: 9 bins around the arg-max of the zero-padded salience.  The products (float32 salience x float64 cents)
    // and their sum are float64, the weight sum is a FLOAT32 sum of the float32 salience (np.sum keeps the dtype); numpy sums 9
    // contiguous values as ((a0+a1)+(a2+a3))+((a4+a5)+(a6+a7)) + a8
    double p[9];
    float q[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int c = bi - 4 + k;
      const bool in_range = c >= 0 && c < kClasses;
      const float sv = in_range ? sal[warp][c] : 0.f;
      const double cents = in_range ? 20.0 * (double)c + 1997.3794084376191 : 0.0;
      p[k] = (double)sv * cents;
      q[k] = sv;
    }
    const double ps = (((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]))) + p[8];
    const float ws = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(q[0], q[1]), __fadd_rn(q[2], q[3])),
                                         __fadd_rn(__fadd_rn(q[4], q[5]), __fadd_rn(q[6], q[7]))), q[8]);
    double cents_pred = ps / (double)ws;
    if (best <= thred) cents_pred = 0.0;
    if (cents_out) cents_out[t] = cents_pred;
    double f = 10.0 * exp2(cents_pred / 1200.0);                      // rmvpe.py:612
    if (f == 10.0) f = 0.0;                                           // rmvpe.py:613
    if (f0) f0[t] = f;
  }
}

inline unsigned blocks_for(long long total, int threads, int cap = 148 * 16) {
  long long b = (total + threads - 1) / threads;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

cudaError_t launch_rmvpe_logmel(const float* audio, long long n, const float* window, const void* twiddle, const float* mel_basis,
                                const void* mel_range, float bn_scale, float bn_shift, float* mel_out, void* img, int n_frames,
                                int frames_out, int img_pitch, int img_c, cudaStream_t st) {
  if (!audio || !window || !twiddle || !mel_basis || !mel_range || n <= kNfft / 2 || n_frames != (int)(n / kHop) + 1 ||
      frames_out < n_frames || frames_out > 2 * n_frames - 1 || (!mel_out && !img) || (img && (img_pitch < kMels || img_c < 1)))
    return cudaErrorInvalidValue;
  rmvpe_logmel_kernel<<<frames_out, 256, 0, st>>>(audio, n, window, reinterpret_cast<const float2*>(twiddle), mel_basis,
                                                  reinterpret_cast<const int2*>(mel_range), bn_scale, bn_shift, mel_out,
                                                  reinterpret_cast<__half*>(img), n_frames, img_pitch, img_c);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_rmvpe_pool(const float* x, int ldx, void* y16, int H2, int W2, int C, int p_in, int p_out, cudaStream_t st) {
  if (!x || !y16 || H2 < 1 || W2 < 1 || C % 8 != 0 || ldx % 4 != 0 || ldx < C || p_in < 2 * W2 || p_out < W2) return cudaErrorInvalidValue;
  const long long total = (long long)H2 * p_out * (C / 8);
  rmvpe_pool_kernel<<<blocks_for(total, 256), 256, 0, st>>>(x, ldx, reinterpret_cast<__half*>(y16), H2, W2, C, p_in, p_out);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_rmvpe_shuffle(const void* g16, void* out16, int H, int W, int Co, int ld, int p_in, int fp_out, int pack,
                                 cudaStream_t st) {
  if (!g16 || !out16 || H < 1 || W < 1 || Co % 8 != 0 || ld % 8 != 0 || pack < 1 || ld < pack * Co || p_in < W || (2 * W) % pack != 0 ||
      fp_out < 2 * W / pack)
    return cudaErrorInvalidValue;
  const long long total = (long long)H * W * 4 * (Co / 8);
  rmvpe_shuffle_kernel<<<blocks_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __half*>(g16), reinterpret_cast<__half*>(out16),
                                                               H, W, Co, ld, p_in, fp_out, pack);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_rmvpe_gru_pack(const float* y, int ldc, void* x16, long long T, int W, int pitch, cudaStream_t st) {
  if (!y || !x16 || T < 1 || W < 1 || ldc < 3 || pitch < W) return cudaErrorInvalidValue;
  rmvpe_gru_pack_kernel<<<blocks_for(T * 3 * W, 256), 256, 0, st>>>(y, ldc, reinterpret_cast<__half*>(x16), T, W, pitch);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_rmvpe_gru(const float* gi, const float* w_hh, const float* b_hh, void* out16, float* out32, int T,
                             cudaStream_t st) {
  if (!gi || !w_hh || !b_hh || !out16 || T < 1) return cudaErrorInvalidValue;
  rmvpe_gru_kernel<<<2 * kGruCl, 256, 0, st>>>(gi, w_hh, b_hh, reinterpret_cast<__half*>(out16), out32, T);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_rmvpe_mel_to_img(const float* mel, void* img, int n_frames, int frames_out, float bn_scale, float bn_shift,
                                    int img_pitch, int img_c, cudaStream_t st) {
  if (!mel || !img || n_frames < 1 || frames_out < n_frames || frames_out > 2 * n_frames - 1 || img_pitch < kMels || img_c < 1)
    return cudaErrorInvalidValue;
  rmvpe_mel_to_img_kernel<<<blocks_for((long long)frames_out * kMels, 256), 256, 0, st>>>(mel, reinterpret_cast<__half*>(img), n_frames,
                                                                                         frames_out, bn_scale, bn_shift, img_pitch, img_c);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_rmvpe_decode(const float* in, int ld, int from_hidden, float* hidden, double* f0, double* cents, int T, float thred,
                                cudaStream_t st) {
  if (!in || (!f0 && !cents) || T < 1 || ld < kClasses) return cudaErrorInvalidValue;
  rmvpe_decode_kernel<<<(T + kDecWarps - 1) / kDecWarps, 32 * kDecWarps, 0, st>>>(in, ld, from_hidden, hidden, f0, cents, T, thred);
  launch_counter().n++;
  return cudaGetLastError();
}

}  // namespace rvc
