// tcgen05 / TMEM implicit-GEMM 1-D convolution for sm_100a (fp16 or bf16 operands, fp32 accumulate).
//
//   D[128 time rows][N = C_out tile] (TMEM, fp32) += sum_{k-block} sum_{tap} A_tap[128][KB] * W_tap[N][KB]^T
//
// Data layout ("planar-vector", PV): activations live as [B][C/8][Lp][8] 16-bit (and
// [B][C/4][Lp][4] fp32 for the residual stream), i.e. one 16-byte vector per (row, channel group),
// rows contiguous inside a channel-group plane, PADF zero rows in front and >= 96 behind.  This is
// exactly the tcgen05 no-swizzle K-major canonical layout ((8,m),(8,2)):((16B,128B),(2B,LBO))
// (cute/atom/mma_traits_sm100.hpp "LayoutType::INTERLEAVE"), with LBO = plane pitch, so
//   * one bulk copy (cp.async.bulk, UBLKCP) per channel group stages 128+halo rows ONCE per
//     64-channel k-block and every tap of a dilated kernel is just a descriptor start-address
//     offset of tap*dil*16 bytes into that slab (no im2col, no re-fetch per tap);
//   * zero padding at sequence ends is physical (the pad rows are never written);
//   * the epilogue (thread = TMEM lane = time row) reads residuals and writes outputs as 512-byte
//     coalesced warp transactions.
// Weights are pre-packed on the host into the smem image [tap][k-block][8 groups][N][8] so one
// bulk copy per (k-block, tap) feeds the B operand.
//
// Warp roles (192 threads): warp 0 = bulk-copy producer (one lane), warp 1 = TMEM allocator + MMA
// issuer (one lane), warps 2..5 = epilogue (TMEM lane quadrant = warp_id % 4).
// Replaces the cuDNN calls behind modules.py:295-308 (ResBlock1), models.py:545-551 (conv_pre, ups).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_tc.cuh"

namespace rvc {
namespace {

constexpr int kEpiWarps = 8;
constexpr int kThreadsTC = 64 + 32 * kEpiWarps;   // producer warp + MMA warp + epilogue warps
constexpr int BM = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major, no swizzle: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t pack2(bool bf16, float a, float b) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

// One epilogue pass over CH accumulator columns of this thread's row: TMEM -> regs, + bias/cond/residual,
// accumulate, /div, fp32 store, lrelu + 16-bit store.  All loads of the chunk are issued before any use.
template <int CH>
__device__ __forceinline__ void epilogue_chunk(const TcConvDesc& p, uint32_t taddr, bool row_ok, int co, size_t pitch_o,
                                               size_t orow16, unsigned char* y32, unsigned char* y16,
                                               const unsigned char* r32, const float* cond) {
  uint32_t r[CH];
  if (CH == 32) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
        "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16 % CH]),
          "=r"(r[17 % CH]), "=r"(r[18 % CH]), "=r"(r[19 % CH]), "=r"(r[20 % CH]), "=r"(r[21 % CH]), "=r"(r[22 % CH]),
          "=r"(r[23 % CH]), "=r"(r[24 % CH]), "=r"(r[25 % CH]), "=r"(r[26 % CH]), "=r"(r[27 % CH]), "=r"(r[28 % CH]),
          "=r"(r[29 % CH]), "=r"(r[30 % CH]), "=r"(r[31 % CH])
        : "r"(taddr));
  } else {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
  }
  // residual / accumulator loads overlap the TMEM read
  float4 rr[CH / 4], aa[CH / 4];
  if (row_ok) {
    if (r32) {
#pragma unroll
      for (int k4 = 0; k4 < CH / 4; ++k4)
        rr[k4] = *reinterpret_cast<const float4*>(r32 + (size_t)(co / 4 + k4) * pitch_o + orow16);
    }
    if (p.accum) {
#pragma unroll
      for (int k4 = 0; k4 < CH / 4; ++k4)
        aa[k4] = *reinterpret_cast<const float4*>(y32 + (size_t)(co / 4 + k4) * pitch_o + orow16);
    }
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (!row_ok) return;
  float v[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(r[i]) + __ldg(p.bias + co + i);
  if (cond) {
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] += __ldg(cond + co + i);
  }
  if (r32) {
#pragma unroll
    for (int k4 = 0; k4 < CH / 4; ++k4) {
      v[k4 * 4 + 0] += rr[k4].x; v[k4 * 4 + 1] += rr[k4].y; v[k4 * 4 + 2] += rr[k4].z; v[k4 * 4 + 3] += rr[k4].w;
    }
  }
  if (p.accum) {
#pragma unroll
    for (int k4 = 0; k4 < CH / 4; ++k4) {
      v[k4 * 4 + 0] += aa[k4].x; v[k4 * 4 + 1] += aa[k4].y; v[k4 * 4 + 2] += aa[k4].z; v[k4 * 4 + 3] += aa[k4].w;
    }
  }
  if (p.div != 1.f) {
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = v[i] / p.div;
  }
  if (y32) {
#pragma unroll
    for (int k4 = 0; k4 < CH / 4; ++k4)
      *reinterpret_cast<float4*>(y32 + (size_t)(co / 4 + k4) * pitch_o + orow16) =
          make_float4(v[k4 * 4 + 0], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]);
  }
  if (y16) {
    const bool obf = p.out_bf16 != 0;
#pragma unroll
    for (int k8 = 0; k8 < CH / 8; ++k8) {
      uint4 o;
      o.x = pack2(obf, lrelu(v[k8 * 8 + 0], p.out_slope), lrelu(v[k8 * 8 + 1], p.out_slope));
      o.y = pack2(obf, lrelu(v[k8 * 8 + 2], p.out_slope), lrelu(v[k8 * 8 + 3], p.out_slope));
      o.z = pack2(obf, lrelu(v[k8 * 8 + 4], p.out_slope), lrelu(v[k8 * 8 + 5], p.out_slope));
      o.w = pack2(obf, lrelu(v[k8 * 8 + 6], p.out_slope), lrelu(v[k8 * 8 + 7], p.out_slope));
      *reinterpret_cast<uint4*>(y16 + (size_t)(co / 8 + k8) * pitch_o + orow16) = o;
    }
  }
}

// Persistent, warp-specialised: each CTA walks tiles t = blockIdx.x, +gridDim.x, ...; the accumulator is
// double-buffered in TMEM so the epilogue of tile i overlaps the loads and MMAs of tile i+1.
__global__ void __launch_bounds__(kThreadsTC, 1) conv_tc_kernel(const TcConvDesc p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int halo = (p.ntaps - 1) * p.dil;
  const int R = (BM + halo + 7) & ~7;              // slab rows
  const int ngrp = p.KB / 8;                       // 16-byte channel groups per k-block
  const int nkb = p.Cin / p.KB;
  const uint32_t a_bytes = (uint32_t)ngrp * R * 16;
  const uint32_t b_bytes = (uint32_t)p.N * p.KB * 2;
  const int NA = p.na_stages, NB = p.nb_stages;
  const uint32_t a_stride = (a_bytes + 127) & ~127u, b_stride = (b_bytes + 127) & ~127u;

  unsigned char* slabA = smem;                              // [NA][ngrp][R][16]
  unsigned char* slabB = smem + (size_t)NA * a_stride;      // [NB][ngrp][N][16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(slabB + (size_t)NB * b_stride);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + NA;
  uint64_t* b_full = a_empty + NA;
  uint64_t* b_empty = b_full + NB;
  uint64_t* acc_full = b_empty + NB;     // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int n_mt = (p.Lj + BM - 1) / BM;
  const int n_nt = p.Cout_total / p.N;
  const long long total_tiles = (long long)n_mt * n_nt * p.G * p.batch;
  const int my_tiles = (int)((total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: two accumulator buffers of N columns (power of two >= 32 in total)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](long long t, int& mt, int& nt, int& g, int& b) {
    mt = (int)(t % n_mt); t /= n_mt;
    nt = (int)(t % n_nt); t /= n_nt;
    g = (int)(t % p.G);
    b = (int)(t / p.G);
  };

  if (warp == 0) {
    // =========================== producer: bulk copies global -> smem ===========================
    if (lane == 0) {
      const int Q = my_tiles * nkb;                   // flattened (tile, k-block) sequence
      auto issue_a = [&](int q) {
        int mt, nt, g, b;
        decode((long long)blockIdx.x + (long long)(q / nkb) * gridDim.x, mt, nt, g, b);
        const int kb = q % nkb;
        const unsigned char* xa = reinterpret_cast<const unsigned char*>(p.x16) +
                                  ((size_t)b * (p.Cin / 8)) * p.Lp_in * 16 + (size_t)(mt * BM + p.g_off[g] + p.padf) * 16;
        const int sa = q % NA;
        mbar_wait(&a_empty[sa], ((q / NA) & 1) ^ 1);
        mbar_expect_tx(&a_full[sa], a_bytes);
        for (int c = 0; c < ngrp; ++c)
          bulk_g2s(slabA + sa * a_stride + (size_t)c * R * 16, xa + (size_t)(kb * ngrp + c) * p.Lp_in * 16, R * 16,
                   &a_full[sa]);
      };
      const int ahead = NA - 1;                       // slabs issued ahead of the one being consumed
      if (p.b_stationary) {
        // all (k-block, tap) weight tiles fit in smem: load them once, they serve every tile of this CTA
        const unsigned char* wb = reinterpret_cast<const unsigned char*>(p.w16);
        for (int kb = 0; kb < nkb; ++kb)
          for (int tap = 0; tap < p.ntaps; ++tap) {
            const int sb = kb * p.ntaps + tap;
            mbar_expect_tx(&b_full[sb], b_bytes);
            bulk_g2s(slabB + sb * b_stride, wb + ((size_t)tap * nkb + kb) * b_bytes, b_bytes, &b_full[sb]);
          }
        for (int q = 0; q < Q; ++q) issue_a(q);       // issue_a blocks on a_empty: NA slabs in flight
      } else {
        for (int q = 0; q < ahead && q < Q; ++q) issue_a(q);
        int itb = 0;
        for (int q = 0; q < Q; ++q) {
          int mt, nt, g, b;
          decode((long long)blockIdx.x + (long long)(q / nkb) * gridDim.x, mt, nt, g, b);
          const int kb = q % nkb;
          const unsigned char* wb = reinterpret_cast<const unsigned char*>(p.w16) +
                                    ((size_t)(g * n_nt + nt) * p.ntaps * nkb) * b_bytes;
          for (int tap = 0; tap < p.ntaps; ++tap, ++itb) {
            const int sb = itb % NB;
            mbar_wait(&b_empty[sb], ((itb / NB) & 1) ^ 1);
            mbar_expect_tx(&b_full[sb], b_bytes);
            bulk_g2s(slabB + sb * b_stride, wb + ((size_t)tap * nkb + kb) * b_bytes, b_bytes, &b_full[sb]);
          }
          if (q + ahead < Q) issue_a(q + ahead);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (single thread) =====================================
    if (lane == 0) {
      const uint32_t fmt = p.in_bf16 ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t lbo_a = (uint32_t)R * 16, lbo_b = (uint32_t)p.N * 16;
      const int ksteps = p.KB / 16;
      const uint64_t desc_a0 = make_desc(0, lbo_a, 128), desc_b0 = make_desc(0, lbo_b, 128);
      const bool stat = p.b_stationary != 0;
      int itb = 0, q = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int buf = t & 1;
        mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.N);
        uint32_t accum = 0;
        for (int kb = 0; kb < nkb; ++kb, ++q) {
          const int sa = q % NA;
          mbar_wait(&a_full[sa], (q / NA) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(slabA + sa * a_stride);
          for (int tap = 0; tap < p.ntaps; ++tap, ++itb) {
            int sb;
            if (stat) {
              sb = kb * p.ntaps + tap;
              if (t == 0) { mbar_wait(&b_full[sb], 0); tc_fence_after(); }
            } else {
              sb = itb % NB;
              mbar_wait(&b_full[sb], (itb / NB) & 1);
              tc_fence_after();
            }
            const uint64_t ad0 = desc_a0 + (uint64_t)((a_base + (uint32_t)(tap * p.dil) * 16) >> 4);
            const uint64_t bd0 = desc_b0 + (uint64_t)(smem_u32(slabB + sb * b_stride) >> 4);
            for (int ks = 0; ks < ksteps; ++ks) {
              tc_mma_f16(d_tmem, ad0 + (uint64_t)((2 * ks * lbo_a) >> 4), bd0 + (uint64_t)((2 * ks * lbo_b) >> 4), idesc, accum);
              accum = 1;
            }
            if (!stat) tc_commit(&b_empty[sb]);
          }
          tc_commit(&a_empty[sa]);
        }
        tc_commit(&acc_full[buf]);
      }
    }
  } else {
    // =========================== epilogue: 8 warps, TMEM -> registers -> global ==================
    const int ew = warp - 2;                        // 0..7
    const int qd = warp & 3;                        // TMEM lane quadrant this warp may access
    const int half = ew >> 2;                       // which half of the N columns
    const int wc = p.N >= 32 ? p.N / 2 : p.N;       // columns per warp
    const bool active = p.N >= 32 || half == 0;
    const size_t pitch_o = (size_t)p.Lp_out * 16;
    for (int t = 0; t < my_tiles; ++t) {
      const int buf = t & 1;
      int mt, nt, g, b;
      decode((long long)blockIdx.x + (long long)t * gridDim.x, mt, nt, g, b);
      mbar_wait(&acc_full[buf], (t >> 1) & 1);
      tc_fence_after();
      if (active) {
        const int row = mt * BM + qd * 32 + lane;
        const bool row_ok = row < p.Lj;
        const size_t orow16 = (size_t)((long long)row * p.out_stride + g + p.padf) * 16;
        unsigned char* y32 = p.y32 ? reinterpret_cast<unsigned char*>(p.y32) + (size_t)b * (p.Cout_total / 4) * pitch_o : nullptr;
        unsigned char* y16 = p.y16 ? reinterpret_cast<unsigned char*>(p.y16) + (size_t)b * (p.Cout_total / 8) * pitch_o : nullptr;
        const unsigned char* r32 =
            p.res32 ? reinterpret_cast<const unsigned char*>(p.res32) + (size_t)b * (p.Cout_total / 4) * pitch_o : nullptr;
        const float* cond = p.cond ? p.cond + (size_t)b * p.cond_bstride : nullptr;
        const int c_begin = half * wc;
        const uint32_t tbase = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * p.N);
        if (wc % 32 == 0) {
          for (int c0 = c_begin; c0 < c_begin + wc; c0 += 32)
            epilogue_chunk<32>(p, tbase + (uint32_t)c0, row_ok, nt * p.N + c0, pitch_o, orow16, y32, y16, r32, cond);
        } else {
          for (int c0 = c_begin; c0 < c_begin + wc; c0 += 16)
            epilogue_chunk<16>(p, tbase + (uint32_t)c0, row_ok, nt * p.N + c0, pitch_o, orow16, y32, y16, r32, cond);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[buf])) : "memory");
    }
  }
  // ------------------------------------ teardown -------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
}

size_t tc_smem_bytes(const TcConvDesc& d) {
  const int halo = (d.ntaps - 1) * d.dil;
  const int R = (BM + halo + 7) & ~7;
  const size_t a = (((size_t)(d.KB / 8) * R * 16) + 127) & ~(size_t)127;
  const size_t bb = (((size_t)d.N * d.KB * 2) + 127) & ~(size_t)127;
  return d.na_stages * a + d.nb_stages * bb + 8 * (2 * d.na_stages + 2 * d.nb_stages + 4) + 16 + 128;
}

// ---- PV-layout glue kernels ----------------------------------------------------------------------
__global__ void zero_pads_kernel(unsigned char* base, long long planes, int Lp, int padf, long long L) {
  // zero rows [0, padf) and [padf+L, Lp) of every 16-byte-vector plane
  const long long tail0 = padf + L;
  const int npad = padf + (int)(Lp - tail0);
  const long long total = planes * npad;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long pl = idx / npad;
    const int r = (int)(idx % npad);
    const long long row = r < padf ? r : tail0 + (r - padf);
    *reinterpret_cast<uint4*>(base + (pl * Lp + row) * 16) = make_uint4(0, 0, 0, 0);
  }
}

__global__ void cl_to_pv16_kernel(const float* __restrict__ x, int ldx, long long L, int C, unsigned char* __restrict__ y16,
                                  int Lp, int padf, float slope, bool BF16) {
  // x [B][L][ldx] channels-last fp32 -> PV16 [B][C/8][Lp][8]
  const int b = blockIdx.y;
  const int ng = C / 8;
  const long long total = L * ng;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int gch = (int)(idx / L);
    const long long t = idx % L;
    const float* src = x + ((long long)b * L + t) * ldx + gch * 8;
    const float4 a = *reinterpret_cast<const float4*>(src), c = *reinterpret_cast<const float4*>(src + 4);
    uint4 o;
    o.x = pack2(BF16, lrelu(a.x, slope), lrelu(a.y, slope));
    o.y = pack2(BF16, lrelu(a.z, slope), lrelu(a.w, slope));
    o.z = pack2(BF16, lrelu(c.x, slope), lrelu(c.y, slope));
    o.w = pack2(BF16, lrelu(c.z, slope), lrelu(c.w, slope));
    *reinterpret_cast<uint4*>(y16 + (((long long)b * ng + gch) * Lp + padf + t) * 16) = o;
  }
}

__global__ void noise_add_pv_kernel(const float* __restrict__ har, const float* __restrict__ wn, const float* __restrict__ nb,
                                    unsigned char* __restrict__ x32, unsigned char* __restrict__ x16, long long L_har,
                                    long long L, int C, int k, int s, int pad, int Lp, int padf, float slope, bool BF16) {
  // x32 += noise_conv(har);  x16 = cvt(lrelu(x32))   (models.py:552-553 + the lrelu of modules.py:297)
  extern __shared__ float sw[];  // [k][C] + [C]
  for (int i = threadIdx.x; i < k * C; i += blockDim.x) sw[i] = wn[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[k * C + i] = nb[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int ng = C / 8;
  const long long total = L * ng;
  const float* hb = har + (long long)b * L_har;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int gch = (int)(idx / L);
    const long long t = idx % L;
    const int c = gch * 8;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = sw[k * C + c + i];
    const long long h0 = t * s - pad;
    for (int kk = 0; kk < k; ++kk) {
      const long long h = h0 + kk;
      if (h < 0 || h >= L_har) continue;
      const float hv = __ldg(hb + h);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(hv, sw[kk * C + c + i], acc[i]);
    }
    float4* p0 = reinterpret_cast<float4*>(x32 + (((long long)b * (C / 4) + gch * 2) * Lp + padf + t) * 16);
    float4* p1 = reinterpret_cast<float4*>(x32 + (((long long)b * (C / 4) + gch * 2 + 1) * Lp + padf + t) * 16);
    float4 a = *p0, d = *p1;
    a.x += acc[0]; a.y += acc[1]; a.z += acc[2]; a.w += acc[3];
    d.x += acc[4]; d.y += acc[5]; d.z += acc[6]; d.w += acc[7];
    *p0 = a; *p1 = d;
    uint4 o;
    o.x = pack2(BF16, lrelu(a.x, slope), lrelu(a.y, slope));
    o.y = pack2(BF16, lrelu(a.z, slope), lrelu(a.w, slope));
    o.z = pack2(BF16, lrelu(d.x, slope), lrelu(d.y, slope));
    o.w = pack2(BF16, lrelu(d.z, slope), lrelu(d.w, slope));
    *reinterpret_cast<uint4*>(x16 + (((long long)b * ng + gch) * Lp + padf + t) * 16) = o;
  }
}

__global__ void conv_post_pv_kernel(const unsigned char* __restrict__ x32, const float* __restrict__ w, float* __restrict__ out,
                                    long long L, int C, int k, int Lp, int padf, float slope) {
  // out[b][t] = tanh(sum_{kk,c} lrelu(x[c][t+kk-pad]) * w[kk][c]);  x is PV32 with zero pads
  extern __shared__ float swp[];  // [k][C]
  for (int i = threadIdx.x; i < k * C; i += blockDim.x) swp[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int pad = (k - 1) / 2;
  const int n4 = C / 4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < L; t += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int g4 = 0; g4 < n4; ++g4) {
      const unsigned char* pl = x32 + (((long long)b * n4 + g4) * Lp + padf + t - pad) * 16;
      for (int kk = 0; kk < k; ++kk) {
        const float4 v = *reinterpret_cast<const float4*>(pl + (long long)kk * 16);
        const float* wr = swp + kk * C + g4 * 4;
        acc = fmaf(lrelu(v.x, slope), wr[0], acc);
        acc = fmaf(lrelu(v.y, slope), wr[1], acc);
        acc = fmaf(lrelu(v.z, slope), wr[2], acc);
        acc = fmaf(lrelu(v.w, slope), wr[3], acc);
      }
    }
    out[(long long)b * L + t] = tanhf(acc);
  }
}

__global__ void pv32_to_cl_kernel(const unsigned char* __restrict__ x32, float* __restrict__ y, long long L, int C, int Lp, int padf) {
  // debug/tap helper: PV32 -> channels-last [B][L][C]
  const int b = blockIdx.y;
  const int n4 = C / 4;
  const long long total = L * n4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g4 = (int)(idx / L);
    const long long t = idx % L;
    const float4 v = *reinterpret_cast<const float4*>(x32 + (((long long)b * n4 + g4) * Lp + padf + t) * 16);
    *reinterpret_cast<float4*>(y + ((long long)b * L + t) * C + g4 * 4) = v;
  }
}

__global__ void pv16_to_cl_kernel(const unsigned char* __restrict__ x16, float* __restrict__ y, long long L, int C, int Lp, int padf,
                                  bool BF16) {
  const int b = blockIdx.y;
  const int n8 = C / 8;
  const long long total = L * n8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g8 = (int)(idx / L);
    const long long t = idx % L;
    const uint4 v = *reinterpret_cast<const uint4*>(x16 + (((long long)b * n8 + g8) * Lp + padf + t) * 16);
    const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
    float* dst = y + ((long long)b * L + t) * C + g8 * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f;
      if (BF16) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wds[i]));
      else f = __half22float2(*reinterpret_cast<const __half2*>(&wds[i]));
      dst[2 * i] = f.x; dst[2 * i + 1] = f.y;
    }
  }
}

inline unsigned grid_for(long long total, int threads) {
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace

cudaError_t launch_conv_tc(const TcConvDesc& d_in, int B, cudaStream_t st) {
  TcConvDesc d = d_in;
  if (d.N < 16 || d.N > 256 || d.N % 16 != 0 || d.KB % 16 != 0 || d.KB > 64 || d.Cin % d.KB != 0 || d.Cout_total % d.N != 0 ||
      d.G < 1 || d.G > 16 || d.Lj <= 0 || (d.ntaps - 1) * d.dil > 56 || (d.accum && !d.y32) || B <= 0)
    return cudaErrorInvalidValue;
  d.batch = B;
  int cols = 32;
  while (cols < 2 * d.N) cols <<= 1;          // two accumulator buffers
  d.tmem_cols = cols;
  const int nkb = d.Cin / d.KB;
  {
    // smem policy: weights stationary (loaded once per persistent CTA) when every (k-block, tap) tile
    // fits beside >= 3 activation slabs; otherwise a weight ring with >= 128 KB in flight.
    const int halo = (d.ntaps - 1) * d.dil;
    const int R = (BM + halo + 7) & ~7;
    const size_t a = (((size_t)(d.KB / 8) * R * 16) + 127) & ~(size_t)127;
    const size_t bb = (((size_t)d.N * d.KB * 2) + 127) & ~(size_t)127;
    const size_t budget = 212 * 1024;
    const int nw = nkb * d.ntaps;
    const long long tiles_ = (long long)((d.Lj + BM - 1) / BM) * (d.Cout_total / d.N) * d.G * B;
    d.b_stationary = 0;
    if (d.G == 1 && d.Cout_total == d.N && nw <= 48 && (size_t)nw * bb + 3 * a <= budget && tiles_ > 2 * 148) {
      d.b_stationary = 1;
      d.nb_stages = nw;
      long long na = (long long)(budget - (size_t)nw * bb) / (long long)a;
      d.na_stages = (int)(na > 8 ? 8 : na);
    } else {
      d.na_stages = nkb >= 2 ? 3 : 2;
      long long nb = (long long)(budget - (size_t)d.na_stages * a) / (long long)bb;
      d.nb_stages = (int)(nb > 12 ? 12 : (nb < 2 ? 2 : nb));
    }
  }
  size_t smem = tc_smem_bytes(d);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  static size_t cfgd = 0;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  }
  if (smem > cfgd) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cfgd = smem;
  }
  const long long tiles = (long long)((d.Lj + BM - 1) / BM) * (d.Cout_total / d.N) * d.G * B;
  const unsigned grid = (unsigned)(tiles < num_sms ? tiles : num_sms);      // persistent: one CTA per SM
  conv_tc_kernel<<<grid, kThreadsTC, smem, st>>>(d);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_zero_pads(void* base, long long planes, int Lp, int padf, long long L, cudaStream_t st) {
  const long long total = planes * (Lp - L);
  zero_pads_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<unsigned char*>(base), planes, Lp, padf, L);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_cl_to_pv16(const float* x, int ldx, int B, long long L, int C, void* y16, int Lp, int padf, float slope,
                              bool bf16, cudaStream_t st) {
  dim3 grid(grid_for(L * (C / 8), 256), B);
  cl_to_pv16_kernel<<<grid, 256, 0, st>>>(x, ldx, L, C, reinterpret_cast<unsigned char*>(y16), Lp, padf, slope, bf16);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_noise_add_pv(const float* har, const float* wn, const float* nb, void* x32, void* x16, int B,
                                long long L_har, long long L, int C, int k, int s, int pad, int Lp, int padf, float slope,
                                bool bf16, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)k * C + C);
  static size_t cfg = 48 * 1024;
  if (smem > cfg) {
    cudaError_t e = cudaFuncSetAttribute(noise_add_pv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cfg = smem;
  }
  dim3 grid(grid_for(L * (C / 8), 256), B);
  noise_add_pv_kernel<<<grid, 256, smem, st>>>(har, wn, nb, reinterpret_cast<unsigned char*>(x32),
                                               reinterpret_cast<unsigned char*>(x16), L_har, L, C, k, s, pad, Lp, padf, slope, bf16);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_conv_post_pv(const void* x32, const float* w, float* out, int B, long long L, int C, int k, int Lp,
                                int padf, float slope, cudaStream_t st) {
  dim3 grid(grid_for(L, 256), B);
  conv_post_pv_kernel<<<grid, 256, sizeof(float) * k * C, st>>>(reinterpret_cast<const unsigned char*>(x32), w, out, L, C, k, Lp,
                                                                padf, slope);
  launch_counter().n++;
  return cudaGetLastError();
}

cudaError_t launch_pv_to_cl(const void* src, bool is16, bool bf16, float* y, int B, long long L, int C, int Lp, int padf,
                            cudaStream_t st) {
  if (!is16) {
    dim3 grid(grid_for(L * (C / 4), 256), B);
    pv32_to_cl_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(src), y, L, C, Lp, padf);
  } else {
    dim3 grid(grid_for(L * (C / 8), 256), B);
    pv16_to_cl_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(src), y, L, C, Lp, padf, bf16);
  }
  launch_counter().n++;
  return cudaGetLastError();
}

}  // namespace rvc
