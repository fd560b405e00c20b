// Bandwidth-bound glue kernels of the synthesis path (channels-last fp32).
#include <cuda_fp16.h>

#include <stdlib.h>

#include "common.cuh"

namespace rvc {

LaunchCounter& launch_counter() {
  static LaunchCounter c;
  return c;
}

int current_num_sms() {
  static std::atomic<int> cache[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  const bool cached = dev >= 0 && dev < kMaxDevices;
  int n = cached ? cache[dev].load(std::memory_order_acquire) : 0;
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (cached) cache[dev].store(n, std::memory_order_release);
  return n;
}

// RVCB200_PDL=0/1 forces programmatic dependent launch off/on; unset, the engine turns it on per infer() call for
// launch-bound sizes (pdl_set_auto): measured 5-8 % on 1-10 s segments (the step is ~165 kernels of 6-30 us, so the
// prologue of kernel n+1 under the tail of kernel n is visible), nothing on a 60 s segment.
namespace {
thread_local bool g_pdl_auto = false;
}
void pdl_set_auto(bool on) { g_pdl_auto = on; }
bool pdl_enabled() {
  static const int env = [] { const char* e = getenv("RVCB200_PDL"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
  return env >= 0 ? env != 0 : g_pdl_auto;
}

namespace {

// ---- LayerNorm over contiguous channels: one warp per row (modules.py:25-28) ---------------
template <int NV>   // NV = ceil(C / 32) values per lane: 8 (C <= 256, the synthesizer) or 32 (C <= 1024, the HuBERT front end)
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float* __restrict__ y, long long rows, int C,
                                 float eps, __half* __restrict__ y16, const int* __restrict__ len, int T) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (threadIdx.x == 0) pdl_trigger();
  pdl_wait();
  if (row >= rows) return;
  const float* xr = x + row * C;
  float v[NV];
  float s = 0.f;
  int n = 0;
  for (int c = lane; c < C; c += 32) { v[n] = xr[c]; s += v[n]; ++n; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
  for (int i = 0; i < n; ++i) { float d = v[i] - mean; q += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + eps);
  float* yr = y + row * C;
  n = 0;
  // optional fp16 copy (the MMA operand of the next contraction), zeroed for rows >= len (x * x_mask)
  const bool keep = !len || (int)(row % T) < len[row / T];
  for (int c = lane; c < C; c += 32) {
    const float o = (v[n] - mean) * rstd * gamma[c] + beta[c];
    yr[c] = o;
    if (y16) y16[row * C + c] = __float2half_rn(keep ? o : 0.f);
    ++n;
  }
}

__global__ void len_to_i32_kernel(const long long* len64, int* len32, int B, int T) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    long long l = len64[b];
    len32[b] = (int)(l < 0 ? 0 : (l > T ? T : l));
  }
}

// ---- speaker conditioning: every cond 1x1 conv on g = emb_g[sid] in one launch -------------
// (models.py:683/799 emb_g; models.py:546-547 dec.cond; modules.py:189 WN cond_layer)
__global__ void cond_gemv_kernel(const float* __restrict__ emb_g, const long long* __restrict__ sid,
                                 const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ out,
                                 int gin, int n_out, int n_spk) {
  const int b = blockIdx.y;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_out) return;
  long long s = sid[b];
  if (s < 0) s = 0;
  if (s >= n_spk) s = n_spk - 1;
  const float* gvec = emb_g + s * gin;
  const float* w = W + (long long)warp * gin;
  float acc = 0.f;
  for (int i = lane; i < gin; i += 32) acc = fmaf(w[i], gvec[i], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[(long long)b * n_out + warp] = acc + bias[warp];
}

// ---- prior sample: z_p = (m + exp(logs)*eps*0.66666)*mask, noise read channels-first --------
__global__ void zp_sample_kernel(const float* __restrict__ stats, const float* __restrict__ noise,
                                 const int* __restrict__ len, float* __restrict__ zp, int T, int C,
                                 float* __restrict__ z, __half* __restrict__ z16) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const float* nb = noise + (long long)b * C * T;
  for (int r = ty; r < 32; r += 8) {
    int c = c0 + r, t = t0 + tx;
    tile[r][tx] = (c < C && t < T) ? nb[(long long)c * T + t] : 0.f;
  }
  __syncthreads();
  const int L = len[b];
  for (int r = ty; r < 32; r += 8) {
    int t = t0 + r, c = c0 + tx;
    if (t < T && c < C) {
      const float* sr = stats + ((long long)b * T + t) * (2 * C);
      float m = sr[c], lg = sr[C + c];
      float v = m + expf(lg) * tile[tx][r] * 0.66666f;
      const float o = (t < L) ? v : 0.f;
      const long long oi = ((long long)b * T + t) * C + c;
      zp[oi] = o;
      if (z) z[oi] = o;
      if (z16) z16[oi] = __float2half_rn(o);
    }
  }
}

// ---- noise_convs[i](har_source) added in place (models.py:552-553) -------------------------
template <int VEC>
__global__ void noise_conv_add_kernel(const float* __restrict__ har, const float* __restrict__ wn,
                                      const float* __restrict__ nb, float* __restrict__ y, long long L_har,
                                      long long L_out, int C, int k, int s, int pad) {
  extern __shared__ float sw[];  // [k][C] + bias[C]
  for (int i = threadIdx.x; i < k * C; i += blockDim.x) sw[i] = wn[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[k * C + i] = nb[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int cv = C / VEC;  // vectors per row
  const long long total = L_out * cv;
  const float* hb = har + (long long)b * L_har;
  float* yb = y + (long long)b * L_out * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / cv;
    const int c = (int)(idx % cv) * VEC;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    const long long h0 = t * s - pad;
    for (int kk = 0; kk < k; ++kk) {
      const long long h = h0 + kk;
      if (h < 0 || h >= L_har) continue;
      const float hv = __ldg(hb + h);
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = fmaf(hv, sw[kk * C + c + v], acc[v]);
    }
    float4* yp = reinterpret_cast<float4*>(yb + t * C + c);
    float4 o = *yp;
    o.x += acc[0] + sw[k * C + c]; o.y += acc[1] + sw[k * C + c + 1];
    o.z += acc[2] + sw[k * C + c + 2]; o.w += acc[3] + sw[k * C + c + 3];
    *yp = o;
  }
}

// ---- lrelu -> conv_post (C -> 1, k taps, no bias) -> tanh (models.py:561-563) ---------------
// One warp per 32 consecutive output samples: lane = channel, loop over taps with shuffles.
__global__ void conv_post_tanh_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out,
                                      long long L, int C, int k, float slope) {
  // block handles TB output samples; stage lrelu(x) rows [t0-pad, t0+TB+pad) in smem
  extern __shared__ float sx[];  // [(TB + k - 1)][C+1]
  const int TB = blockDim.x;
  const int pad = (k - 1) / 2;
  const int b = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * TB;
  const float* xb = x + (long long)b * L * C;
  const int rows = TB + k - 1;
  const int ldc = C + 1;
  for (int idx = threadIdx.x; idx < rows * C; idx += blockDim.x) {
    int r = idx / C, c = idx % C;
    long long t = t0 - pad + r;
    float v = 0.f;
    if (t >= 0 && t < L) v = lrelu(xb[t * C + c], slope);
    sx[r * ldc + c] = v;
  }
  float* swt = sx + rows * ldc;  // [k][C]
  for (int idx = threadIdx.x; idx < k * C; idx += blockDim.x) swt[idx] = w[idx];
  __syncthreads();
  const long long t = t0 + threadIdx.x;
  if (t < L) {
    float acc = 0.f;
    for (int kk = 0; kk < k; ++kk) {
      const float* xr = sx + (threadIdx.x + kk) * ldc;
      const float* wr = swt + kk * C;
      for (int c = 0; c < C; ++c) acc = fmaf(xr[c], wr[c], acc);
    }
    out[(long long)b * L + t] = tanhf(acc);
  }
}

__global__ void copy_rows_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd,
                                 long long rows, int C) {
  const int cv = C / 4;
  const long long total = rows * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / cv;
    int c = (int)(idx % cv) * 4;
    *reinterpret_cast<float4*>(dst + r * ldd + c) = *reinterpret_cast<const float4*>(src + r * lds + c);
  }
}

}  // namespace

cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, float* y, long long rows, int C,
                             float eps, cudaStream_t st, void* y16, const int* len, int T) {
  if (C > 1024 || rows <= 0) return cudaErrorInvalidValue;
  const int wpb = 8;
  const dim3 grid((unsigned)((rows + wpb - 1) / wpb)), block(wpb * 32);
  if (C <= 256)
    launch_pdl(layernorm_kernel<8>, grid, block, 0, st, x, gamma, beta, y, rows, C, eps, reinterpret_cast<__half*>(y16), len,
               T > 0 ? T : 1);
  else
    launch_pdl(layernorm_kernel<32>, grid, block, 0, st, x, gamma, beta, y, rows, C, eps, reinterpret_cast<__half*>(y16), len,
               T > 0 ? T : 1);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_len_to_i32(const long long* len64, int* len32, int B, int T, cudaStream_t st) {
  len_to_i32_kernel<<<(B + 127) / 128, 128, 0, st>>>(len64, len32, B, T);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_cond_gemv(const float* emb_g, const long long* sid, const float* W, const float* bias, float* out,
                             int B, int gin, int n_out, int n_spk, cudaStream_t st) {
  dim3 grid((n_out * 32 + 255) / 256, B);
  cond_gemv_kernel<<<grid, 256, 0, st>>>(emb_g, sid, W, bias, out, gin, n_out, n_spk);
  launch_counter().n++;
  return cudaGetLastError();
}

// Flow, tensor path: the [x0 | mask | zeros] block in front of the gate outputs of the flow's activation buffer
// (weights.py `flow.%d.inf.%d.w`): x0 = channels [in_off, in_off + half) of z16, the mask channel is 1 on frames < len.
__global__ void flow_x0_init_kernel(const __half* __restrict__ z16, const int* __restrict__ len, __half* __restrict__ abuf, int T,
                                    int C, int in_off, int half, int xb, int aw) {
  const long long row = (long long)blockIdx.x * blockDim.y + threadIdx.y;   // b * T + t
  const int b = blockIdx.y;
  const long long t = row;
  if (t >= T) return;
  const bool valid = t < len[b];
  const __half* src = z16 + ((long long)b * T + t) * C + in_off;
  __half* dst = abuf + ((long long)b * T + t) * aw;
  for (int c = threadIdx.x; c < xb; c += blockDim.x)
    dst[c] = c < half ? src[c] : (c == half && valid ? __float2half(1.f) : __float2half(0.f));
}

cudaError_t launch_flow_x0_init(const void* z16, const int* len, void* abuf, int B, int T, int C, int in_off, int half, int xb, int aw,
                                cudaStream_t st) {
  dim3 block(32, 8), grid((T + 7) / 8, B);
  flow_x0_init_kernel<<<grid, block, 0, st>>>(reinterpret_cast<const __half*>(z16), len, reinterpret_cast<__half*>(abuf), T, C, in_off,
                                              half, xb, aw);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_zp_sample(const float* stats, const float* noise_cf, const int* len, float* zp, int B, int T, int C,
                             cudaStream_t st, float* z, void* z16) {
  dim3 grid((T + 31) / 32, (C + 31) / 32, B);
  zp_sample_kernel<<<grid, dim3(32, 8), 0, st>>>(stats, noise_cf, len, zp, T, C, z, reinterpret_cast<__half*>(z16));
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_noise_conv_add(const float* har, const float* wn, const float* nb, float* y, int B, long long L_har,
                                  long long L_out, int C, int k, int s, int pad, cudaStream_t st) {
  if (C % 4 != 0) return cudaErrorInvalidValue;
  const size_t smem = sizeof(float) * ((size_t)k * C + C);
  static SmemOptIn opt;
  if (cudaError_t e = opt_in_smem(noise_conv_add_kernel<4>, smem, opt)) return e;
  const long long total = L_out * (C / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  dim3 grid((unsigned)blocks, B);
  noise_conv_add_kernel<4><<<grid, 256, smem, st>>>(har, wn, nb, y, L_har, L_out, C, k, s, pad);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_conv_post_tanh(const float* x, const float* w, float* out, int B, long long L, int C, int k,
                                  float slope, cudaStream_t st) {
  const int TB = 256;
  const size_t smem = sizeof(float) * ((size_t)(TB + k - 1) * (C + 1) + (size_t)k * C);
  static SmemOptIn opt;
  if (cudaError_t e = opt_in_smem(conv_post_tanh_kernel, smem, opt)) return e;
  dim3 grid((unsigned)((L + TB - 1) / TB), B);
  conv_post_tanh_kernel<<<grid, TB, smem, st>>>(x, w, out, L, C, k, slope);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_copy_rows(const float* src, int lds, float* dst, int ldd, long long rows, int C, cudaStream_t st) {
  if (C % 4 || lds % 4 || ldd % 4) return cudaErrorInvalidValue;
  long long total = rows * (C / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  copy_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, lds, dst, ldd, rows, C);
  launch_counter().n++;
  return cudaGetLastError();
}

}  // namespace rvc
