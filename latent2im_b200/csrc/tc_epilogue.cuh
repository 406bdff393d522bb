// Fused epilogue of the tcgen05 convs, one 32-column accumulator chunk of one pixel row at a time.
//
//   EPI_RAW      : o = acc * demod                          (up-conv output before the blur; data-gradient convs)
//   EPI_ACT_RGB  : y = lrelu(acc*demod + noise + bias)*sqrt2;  o = y * s_next;  rgb[c] += wr[c][co] * y   (plain layers)
//   EPI_ACT      : same without the ToRGB dot products      (composite up-conv)
// The sqrt(2) gain is folded into the staged demod / bias vectors and the noise term.  All per-channel vectors
// are read from statically declared shared memory (LDS, freely scheduled by the compiler around tcgen05.wait::ld).
#pragma once
#include "tc_ptx.cuh"

namespace l2i {
namespace tc {

constexpr int EPI_RAW = 0, EPI_ACT_RGB = 1, EPI_ACT = 2;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// TMA tensor store of a staged shared-memory tile (issued by one lane), bulk-group completion
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// same without closing the bulk group: several stores of one staging buffer form one group
__device__ __forceinline__ void tma_store_4d_nocommit(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 p = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// s_d / s_b / s_n / s_w*: shared-memory vectors already offset to this chunk's first channel.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk32(const uint32_t (&v)[32], const float* s_d, const float* s_b, const float* s_n,
                                                 const float* s_w0, const float* s_w1, const float* s_w2, float nz,
                                                 bool raw_fp16, float& rgb0, float& rgb1, float& rgb2,
                                                 __nv_bfloat16* __restrict__ outc, __nv_bfloat16* __restrict__ yc, int out_swz = 0, int piece_base = 0) {
  uint32_t packed[16];
  if (EPI == EPI_RAW) {
#pragma unroll
    for (int j4 = 0; j4 < 32; j4 += 4) {
      const float4 d4 = *reinterpret_cast<const float4*>(s_d + j4);
      const float o0 = __uint_as_float(v[j4]) * d4.x, o1 = __uint_as_float(v[j4 + 1]) * d4.y;
      const float o2 = __uint_as_float(v[j4 + 2]) * d4.z, o3 = __uint_as_float(v[j4 + 3]) * d4.w;
      packed[j4 >> 1] = raw_fp16 ? pack_f16(o0, o1) : pack_bf16(o0, o1);
      packed[(j4 >> 1) + 1] = raw_fp16 ? pack_f16(o2, o3) : pack_bf16(o2, o3);
    }
  } else {
    uint32_t ypacked[16];
#pragma unroll
    for (int j4 = 0; j4 < 32; j4 += 4) {
      const float4 d4 = *reinterpret_cast<const float4*>(s_d + j4);
      const float4 b4 = *reinterpret_cast<const float4*>(s_b + j4);
      const float4 n4 = *reinterpret_cast<const float4*>(s_n + j4);
      const float dd[4] = {d4.x, d4.y, d4.z, d4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w}, nn[4] = {n4.x, n4.y, n4.z, n4.w};
      float x[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        x[h] = fmaf(__uint_as_float(v[j4 + h]), dd[h], bb[h] + nz);
        x[h] = fmaxf(x[h], 0.2f * x[h]);  // leaky relu
      }
      if (EPI == EPI_ACT_RGB) {
        const float4 w0 = *reinterpret_cast<const float4*>(s_w0 + j4);
        const float4 w1 = *reinterpret_cast<const float4*>(s_w1 + j4);
        const float4 w2 = *reinterpret_cast<const float4*>(s_w2 + j4);
        rgb0 = fmaf(w0.x, x[0], rgb0); rgb0 = fmaf(w0.y, x[1], rgb0); rgb0 = fmaf(w0.z, x[2], rgb0); rgb0 = fmaf(w0.w, x[3], rgb0);
        rgb1 = fmaf(w1.x, x[0], rgb1); rgb1 = fmaf(w1.y, x[1], rgb1); rgb1 = fmaf(w1.z, x[2], rgb1); rgb1 = fmaf(w1.w, x[3], rgb1);
        rgb2 = fmaf(w2.x, x[0], rgb2); rgb2 = fmaf(w2.y, x[1], rgb2); rgb2 = fmaf(w2.z, x[2], rgb2); rgb2 = fmaf(w2.w, x[3], rgb2);
      }
      packed[j4 >> 1] = pack_bf16(x[0] * nn[0], x[1] * nn[1]);
      packed[(j4 >> 1) + 1] = pack_bf16(x[2] * nn[2], x[3] * nn[3]);
      if (yc != nullptr) {  // training: the unscaled activation is kept for the backward pass
        ypacked[j4 >> 1] = pack_bf16(x[0], x[1]);
        ypacked[(j4 >> 1) + 1] = pack_bf16(x[2], x[3]);
      }
    }
    if (yc != nullptr) {
      uint4* dst = reinterpret_cast<uint4*>(yc);
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[k] = make_uint4(ypacked[4 * k], ypacked[4 * k + 1], ypacked[4 * k + 2], ypacked[4 * k + 3]);
    }
  }
  if (outc != nullptr) {   // staging mode (TMA store source): outc = this lane's row of a swizzled tile, 16-byte piece (piece_base + k) ^ out_swz
    uint4* dst = reinterpret_cast<uint4*>(outc);
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[(piece_base + k) ^ out_swz] = make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
  }
}

}  // namespace tc
}  // namespace l2i
