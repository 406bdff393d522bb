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

// Two pixels that share their per-channel vectors (the two rows of a vertical pair, the two tiles of a tile pair ...), 16 channels,
// packed fp32 arithmetic.  Why pairs: a warp-wide LDS.128 of a per-channel vector costs four shared-memory wavefronts even though
// every lane reads the same address, and with six vectors (demod, bias, next style, three ToRGB rows) the parameter traffic of a
// one-pixel-per-lane epilogue is larger than the MMA operand traffic of the small-channel layers (ncu: LDS 43-65 % of the
// shared-memory data pipe).  Every vector fetched here serves both pixels.
//   y = lrelu(acc * d + noise + bias) * sqrt2 (sqrt2 folded into d, bias, noise);  o = bf16(y * s_next);  rgb[c] += wr[c] . y
// rgb2[c] holds (even-channel, odd-channel) partial sums; the caller adds the halves.
template <bool RGB, bool WANT_Y>
__device__ __forceinline__ void epilogue_pair16(const uint32_t (&va)[16], const uint32_t (&vb)[16], const float* s_d, const float* s_b,
                                                const float* s_n, const float* s_w0, const float* s_w1, const float* s_w2, float nza,
                                                float nzb, uint64_t (&rgba)[3], uint64_t (&rgbb)[3], uint32_t (&oa)[8], uint32_t (&ob)[8],
                                                uint32_t (&ya)[8], uint32_t (&yb)[8]) {
  const uint64_t NZA = pk2(nza, nza), NZB = pk2(nzb, nzb), P2 = pk2(0.2f, 0.2f);
#pragma unroll
  for (int j4 = 0; j4 < 16; j4 += 4) {
    const ulonglong2 d4 = *reinterpret_cast<const ulonglong2*>(s_d + j4);
    const ulonglong2 b4 = *reinterpret_cast<const ulonglong2*>(s_b + j4);
    const ulonglong2 n4 = *reinterpret_cast<const ulonglong2*>(s_n + j4);
    ulonglong2 w0 = make_ulonglong2(0ull, 0ull), w1 = w0, w2 = w0;
    if (RGB) {
      w0 = *reinterpret_cast<const ulonglong2*>(s_w0 + j4);
      w1 = *reinterpret_cast<const ulonglong2*>(s_w1 + j4);
      w2 = *reinterpret_cast<const ulonglong2*>(s_w2 + j4);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint64_t D = h ? d4.y : d4.x, Bv = h ? b4.y : b4.x, Nn = h ? n4.y : n4.x;
      const uint64_t W0 = h ? w0.y : w0.x, W1 = h ? w1.y : w1.x, W2 = h ? w2.y : w2.x;
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const uint32_t* v = px ? vb : va;
        const uint64_t x = fma2(pk2u(v[j4 + 2 * h], v[j4 + 2 * h + 1]), D, add2(Bv, px ? NZB : NZA));
        const uint64_t m = mul2(x, P2);
        float x0, x1, m0, m1;
        upk2(x, x0, x1);
        upk2(m, m0, m1);
        const float y0 = fmaxf(x0, m0), y1 = fmaxf(x1, m1);   // leaky relu
        const uint64_t y = pk2(y0, y1);
        if (RGB) {
          uint64_t* rgb = px ? rgbb : rgba;
          rgb[0] = fma2(W0, y, rgb[0]);
          rgb[1] = fma2(W1, y, rgb[1]);
          rgb[2] = fma2(W2, y, rgb[2]);
        }
        float o0, o1;
        upk2(mul2(y, Nn), o0, o1);
        (px ? ob : oa)[(j4 >> 1) + h] = pack_bf16(o0, o1);
        if (WANT_Y) (px ? yb : ya)[(j4 >> 1) + h] = pack_bf16(y0, y1);   // training: the unscaled activation, kept for the backward pass
      }
    }
  }
}

}  // namespace tc
}  // namespace l2i
