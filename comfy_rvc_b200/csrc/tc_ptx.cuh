// PTX wrappers (mbarrier, TMA, tcgen05) and the lean decoder epilogue shared by the tcgen05 convolution kernels
// (conv_tc.cu: generic persistent implicit-GEMM conv; rbconv_tc.cu: compile-time specialised resblock conv).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_tc.cuh"

namespace rvc {
namespace tc {

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
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {   // one lane of a converged warp (elect.sync)
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// descriptors are passed as (lo, hi) 32-bit halves: only the low word (start address) changes per MMA
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
// Same, executed by every lane of a converged warp but issued only where `leader` is set: the operands
// are computed in warp-uniform control flow so they can live in uniform registers.
__device__ __forceinline__ void tc_mma_f16_pred(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t acc, uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.ne.b32 q, %7, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc), "r"(leader) : "memory");
}
__device__ __forceinline__ void tc_commit_pred(uint64_t* bar, uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(leader) : "memory");
}
// K-major SWIZZLE_128B shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B |
// version 1 [46,48) | base_offset [49,52) | layout_type SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t base_offset) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(base_offset & 7u) << 49) | (2ull << 61);
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
// accumulate, /div, fp32 PV store, lrelu + 16-bit channels-last store.  Loads are issued before use.
template <int CH>
__device__ __forceinline__ void epilogue_chunk(const TcConvDesc& p, uint32_t taddr, bool row_ok, int co, size_t pitch_o,
                                               size_t orow16, unsigned char* y32, unsigned char* y16row,
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
  if (y16row) {
    const bool obf = p.out_bf16 != 0;
#pragma unroll
    for (int k8 = 0; k8 < CH / 8; ++k8) {
      uint4 o;
      o.x = pack2(obf, lrelu(v[k8 * 8 + 0], p.out_slope), lrelu(v[k8 * 8 + 1], p.out_slope));
      o.y = pack2(obf, lrelu(v[k8 * 8 + 2], p.out_slope), lrelu(v[k8 * 8 + 3], p.out_slope));
      o.z = pack2(obf, lrelu(v[k8 * 8 + 4], p.out_slope), lrelu(v[k8 * 8 + 5], p.out_slope));
      o.w = pack2(obf, lrelu(v[k8 * 8 + 6], p.out_slope), lrelu(v[k8 * 8 + 7], p.out_slope));
      *reinterpret_cast<uint4*>(y16row + (size_t)(co + k8 * 8) * 2) = o;
    }
  }
}

}  // namespace tc
}  // namespace rvc
