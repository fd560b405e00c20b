// Windowed relative-position multi-head self-attention, fp32, flash-style (never materialises
// the T x T score matrix nor the reference's zero-padded [T, 2T-1] relative tensors).
//
// Restates /root/reference/lib/infer_pack/attentions.py:222-270 in the banded form of
// SURVEY.md App. D:
//     q~ = q / sqrt(dk)
//     S[i][j] = q~_i . k_j + [|j-i| <= w] q~_i . Ek[j-i+w]         (keys j >= len masked)
//     P = softmax_j(S);   O_i = sum_j P[i][j] v_j + sum_{|j-i|<=w} P[i][j] Ev[j-i+w]
// The reference fills masked scores with -1e4, which underflows to an exact 0 weight for
// every valid query, i.e. masked keys are skipped; rows i >= len are written as 0 (the
// reference produces finite garbage there that every consumer multiplies by x_mask).
//
// One CTA = 64 queries of one (batch, head); 128 threads; keys/values stream through smem in
// tiles of 64 with an online softmax.  dk = 96, window = 10 (heads share Ek/Ev).
#include "common.cuh"

namespace rvc {
namespace {

constexpr int BQ = 64, BKV = 64, DK = 96, NTHR = 128;
constexpr int LDT = 65;   // leading dim of Ps[i][j] (scalar, conflict-free column reads)
constexpr int LDK = 68;   // leading dim of the transposed tiles Qs[d][i], Ks[d][j] (float4-aligned rows)
constexpr int MAXR = 21;  // 2*window+1 <= 21

struct AttnSmem {
  float Qs[DK * LDK];
  float Ks[DK * LDK];
  float Vs[BKV * DK];
  float Ps[BQ * LDT];
  float Ek[MAXR * DK];
  float Ev[MAXR * DK];
  float R[BQ * MAXR];
};

__global__ void __launch_bounds__(NTHR) attention_f32_kernel(const float* __restrict__ qkv,
                                                             const float* __restrict__ rel_k,
                                                             const float* __restrict__ rel_v,
                                                             const int* __restrict__ len, float* __restrict__ out,
                                                             int T, int n_heads, int window) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int tj = tid & 7, ti = tid >> 3;  // 8 column groups x 16 row groups
  const int b = blockIdx.z, h = blockIdx.y;
  const int i0 = blockIdx.x * BQ;
  const int H = n_heads * DK;
  const int ld = 3 * H;
  const int L = min(len ? len[b] : T, T);
  const int nrel = 2 * window + 1;
  const float* base = qkv + (long long)b * T * ld;
  float* ob = out + (long long)b * T * H + h * DK;

  if (i0 >= L) {  // whole query tile is padding: zeros
    for (int idx = tid; idx < BQ * (DK / 4); idx += NTHR) {
      int i = idx / (DK / 4), q = idx % (DK / 4);
      if (i0 + i < T) *reinterpret_cast<float4*>(ob + (long long)(i0 + i) * H + q * 4) = make_float4(0, 0, 0, 0);
    }
    return;
  }

  const float inv_sqrt = sqrtf((float)DK);
  // ---- stage Q (scaled), Ek, Ev -------------------------------------------------------------
  for (int idx = tid; idx < BQ * (DK / 4); idx += NTHR) {
    int i = idx % BQ, q = idx / BQ;   // lanes over rows: conflict-free transposing stores
    float4 v = make_float4(0, 0, 0, 0);
    if (i0 + i < T) v = *reinterpret_cast<const float4*>(base + (long long)(i0 + i) * ld + h * DK + q * 4);
    sm.Qs[(q * 4 + 0) * LDK + i] = __fdiv_rn(v.x, inv_sqrt);
    sm.Qs[(q * 4 + 1) * LDK + i] = __fdiv_rn(v.y, inv_sqrt);
    sm.Qs[(q * 4 + 2) * LDK + i] = __fdiv_rn(v.z, inv_sqrt);
    sm.Qs[(q * 4 + 3) * LDK + i] = __fdiv_rn(v.w, inv_sqrt);
  }
  for (int idx = tid; idx < nrel * DK; idx += NTHR) {
    sm.Ek[idx] = rel_k[idx];
    sm.Ev[idx] = rel_v[idx];
  }
  __syncthreads();
  // R[i][r] = q~_i . Ek[r]
  for (int idx = tid; idx < BQ * nrel; idx += NTHR) {
    int i = idx % BQ, r = idx / BQ;
    float acc = 0.f;
    for (int d = 0; d < DK; ++d) acc = fmaf(sm.Qs[d * LDK + i], sm.Ek[r * DK + d], acc);
    sm.R[i * MAXR + r] = acc;
  }

  float o[4][12];
  float m_run[4], l_run[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m_run[a] = -INFINITY;
    l_run[a] = 0.f;
#pragma unroll
    for (int c = 0; c < 12; ++c) o[a][c] = 0.f;
  }

  const int ntiles = (L + BKV - 1) / BKV;
  for (int kt = 0; kt < ntiles; ++kt) {
    const int j0 = kt * BKV;
    __syncthreads();  // previous tile fully consumed (also orders R writes before first use)
    for (int idx = tid; idx < BKV * (DK / 4); idx += NTHR) {
      int j = idx % BKV, q = idx / BKV;   // K: lanes over keys (transposing store)
      float4 kv = make_float4(0, 0, 0, 0);
      if (j0 + j < L) kv = *reinterpret_cast<const float4*>(base + (long long)(j0 + j) * ld + H + h * DK + q * 4);
      sm.Ks[(q * 4 + 0) * LDK + j] = kv.x;
      sm.Ks[(q * 4 + 1) * LDK + j] = kv.y;
      sm.Ks[(q * 4 + 2) * LDK + j] = kv.z;
      sm.Ks[(q * 4 + 3) * LDK + j] = kv.w;
    }
    for (int idx = tid; idx < BKV * (DK / 4); idx += NTHR) {
      int j = idx / (DK / 4), q = idx % (DK / 4);   // V: lanes over channels (row-major store)
      float4 vv = make_float4(0, 0, 0, 0);
      if (j0 + j < L) vv = *reinterpret_cast<const float4*>(base + (long long)(j0 + j) * ld + 2 * H + h * DK + q * 4);
      *reinterpret_cast<float4*>(&sm.Vs[j * DK + q * 4]) = vv;
    }
    __syncthreads();

    // ---- S = Q K^T : thread owns rows ti+16a (a<4), cols tj*4+c and 32+tj*4+c ---------------
    float s[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 8; ++c) s[a][c] = 0.f;
#pragma unroll 4
    for (int d = 0; d < DK; ++d) {
      float qa[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) qa[a] = sm.Qs[d * LDK + ti + 16 * a];
      float kb[8];
      {
        const float4 k0 = *reinterpret_cast<const float4*>(&sm.Ks[d * LDK + tj * 4]);
        const float4 k1 = *reinterpret_cast<const float4*>(&sm.Ks[d * LDK + 32 + tj * 4]);
        kb[0] = k0.x; kb[1] = k0.y; kb[2] = k0.z; kb[3] = k0.w;
        kb[4] = k1.x; kb[5] = k1.y; kb[6] = k1.z; kb[7] = k1.w;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) s[a][c] = fmaf(qa[a], kb[c], s[a][c]);
    }
    // ---- relative-key band, key mask, online softmax -----------------------------------------
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int il = ti + 16 * a;
      const int i = i0 + il;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int j = j0 + (c < 4 ? tj * 4 + c : 32 + tj * 4 + (c - 4));
        const int r = j - i + window;
        if (r >= 0 && r < nrel) s[a][c] += sm.R[il * MAXR + r];
        if (j >= L) s[a][c] = -INFINITY;
        mx = fmaxf(mx, s[a][c]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      const float m_new = fmaxf(m_run[a], mx);   // finite: key 0 of tile 0 is always valid
      const float scale = expf(m_run[a] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float pv = expf(s[a][c] - m_new);  // exp(-inf) = 0 for masked keys
        s[a][c] = pv;
        rs += pv;
      }
      rs += __shfl_xor_sync(0xffffffffu, rs, 1);
      rs += __shfl_xor_sync(0xffffffffu, rs, 2);
      rs += __shfl_xor_sync(0xffffffffu, rs, 4);
      l_run[a] = l_run[a] * scale + rs;
      m_run[a] = m_new;
#pragma unroll
      for (int c = 0; c < 12; ++c) o[a][c] *= scale;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        sm.Ps[il * LDT + tj * 4 + c] = s[a][c];
        sm.Ps[il * LDT + 32 + tj * 4 + c] = s[a][4 + c];
      }
    }
    __syncthreads();
    // ---- O += P V : rows ti+16a, cols tj*4+c + 32m (m<3) --------------------------------------
#pragma unroll 2
    for (int j = 0; j < BKV; ++j) {
      float pa[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) pa[a] = sm.Ps[(ti + 16 * a) * LDT + j];
      float vb[12];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const float4 v4 = *reinterpret_cast<const float4*>(&sm.Vs[j * DK + m * 32 + tj * 4]);
        vb[m * 4 + 0] = v4.x; vb[m * 4 + 1] = v4.y; vb[m * 4 + 2] = v4.z; vb[m * 4 + 3] = v4.w;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 12; ++c) o[a][c] = fmaf(pa[a], vb[c], o[a][c]);
    }
    // ---- relative-value band (only tiles that intersect [i-w, i+w]) ----------------------------
    if (j0 <= i0 + BQ - 1 + window && j0 + BKV - 1 >= i0 - window) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int il = ti + 16 * a;
        const int i = i0 + il;
        for (int r = 0; r < nrel; ++r) {
          const int jl = i + r - window - j0;
          if (jl < 0 || jl >= BKV) continue;
          const float pv = sm.Ps[il * LDT + jl];
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            const float4 e4 = *reinterpret_cast<const float4*>(&sm.Ev[r * DK + m * 32 + tj * 4]);
            o[a][m * 4 + 0] = fmaf(pv, e4.x, o[a][m * 4 + 0]);
            o[a][m * 4 + 1] = fmaf(pv, e4.y, o[a][m * 4 + 1]);
            o[a][m * 4 + 2] = fmaf(pv, e4.z, o[a][m * 4 + 2]);
            o[a][m * 4 + 3] = fmaf(pv, e4.w, o[a][m * 4 + 3]);
          }
        }
      }
    }
  }
  // ---- normalise and store ---------------------------------------------------------------------
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ti + 16 * a;
    if (i >= T) continue;
    const bool valid = i < L;
    const float inv = valid ? 1.f / l_run[a] : 0.f;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      float4 v;
      v.x = valid ? o[a][m * 4 + 0] * inv : 0.f;
      v.y = valid ? o[a][m * 4 + 1] * inv : 0.f;
      v.z = valid ? o[a][m * 4 + 2] * inv : 0.f;
      v.w = valid ? o[a][m * 4 + 3] * inv : 0.f;
      *reinterpret_cast<float4*>(ob + (long long)i * H + m * 32 + tj * 4) = v;
    }
  }
}

}  // namespace

cudaError_t launch_attention_f32(const float* qkv, const float* rel_k, const float* rel_v, const int* len, float* out,
                                 int B, int T, int n_heads, int dk, int window, cudaStream_t st) {
  if (dk != DK || 2 * window + 1 > MAXR || B <= 0 || T <= 0) return cudaErrorInvalidValue;
  static SmemOptIn opt;
  if (cudaError_t e = opt_in_smem(attention_f32_kernel, sizeof(AttnSmem), opt)) return e;
  dim3 grid((T + BQ - 1) / BQ, n_heads, B);
  attention_f32_kernel<<<grid, NTHR, sizeof(AttnSmem), st>>>(qkv, rel_k, rel_v, len, out, T, n_heads, window);
  launch_counter().n++;
  return cudaGetLastError();
}

}  // namespace rvc
