// Data-gradient backward of the synthesis network: grad_image -> grad_latent (the walk-training
// gradient path).  The generator is frozen, so no weight gradient is ever
// formed; with the input-scaling formulation every conv's backward is an ordinary data-gradient
// conv with shared weights plus per-(sample, channel) reductions for the style gradients:
//
//   plain layer   acc = conv(x~, W~);  v = d*acc + nz + bias;  y = lrelu(v)*sqrt2;  x~' = s' * y;  rgb = wr . y
//     g_y   = s' * g_x~'  +  wr^T . g_rgb
//     g_v   = g_y * lrelu'(y) * sqrt2
//     g_acc = d * g_v                      ->  g_x~ = conv_dgrad(g_acc, W~)
//     R_d   = sum_p g_v * (d*acc)          (d*acc is recovered from the saved y, the noise and the bias)
//     R_s'  = sum_p g_x~' * y              (gradient of the NEXT layer's style)
//     R_rgb = sum_p g_rgb[c] * y           (-> ToRGB style gradient through Wrgb)
//   up layer      t = d * convT(x~, W~);  v = blur(t) + nz + bias;  y, x~' as above
//     g_t = blur^T(g_v);  g_acc = d * g_t;  R_d = sum g_t * t;  g_x~ = stride-2 gather conv of g_acc
//   style         s = latent_i . Wmod + b;   d = rsqrt(sum_ci s^2 Wsq + eps)
//     g_s[ci] = R_s(prev layer)[ci]  -  s[ci] * sum_co Wsq[co,ci] * d[co]^2 * R_d[co]
//     g_latent[i] = sum over the layers reading latent i of g_s . Wmod
//
// Reference for what is differentiated: networks.py:231-286, 330-358, 460-514; the reference obtains
// the same gradients from autograd through cuDNN.
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "generator_internal.cuh"

namespace l2i {

template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) GVec { T v[VEC]; };

// ------------------------------------------------------------------------------------------------
// act_bwd: activation / noise / bias / ToRGB backward + the three per-channel reductions.
// grid = (pixel chunks, B), 256 threads: thread = (pixel slot, 16-byte channel vector).
// ------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256, 2)
act_bwd_kernel(T* __restrict__ g_out, const T* __restrict__ g_xn, const float* __restrict__ g_rgb,
               const T* __restrict__ y, const float* __restrict__ s_next, int64_t s_next_bs,
               const float* __restrict__ wr, int64_t wr_bs, const float* __restrict__ demod, int64_t demod_bs,
               const float* __restrict__ noise, int64_t noise_bs, const float* __restrict__ noise_w,
               const float* __restrict__ bias, float* __restrict__ R_s, float* __restrict__ R_d, int64_t R_bs,
               float* __restrict__ R_rgb, int64_t R_rgb_bs, int HW, int C, int chunk) {
  using V = GVec<T, VEC>;
  constexpr float kSqrt2 = 1.4142135623730951f;
  // [5][C] block reductions followed by the [6][C] per-channel constants (kept out of registers: the kernel is
  // latency-bound, two resident CTAs per SM matter more than saving the broadcast LDS)
  extern __shared__ float red[];
  float* prm = red + 5 * C;
  float *p_sn = prm, *p_dm = prm + C, *p_bs = prm + 2 * C, *p_w0 = prm + 3 * C, *p_w1 = prm + 4 * C, *p_w2 = prm + 5 * C;
  const int CV = C / VEC;
  const int slots = 256 / CV;
  const int cv = threadIdx.x % CV, slot = threadIdx.x / CV;
  const int b = blockIdx.y;
  const int c = cv * VEC;
  for (int i = threadIdx.x; i < 5 * C; i += 256) red[i] = 0.f;
  for (int i = threadIdx.x; i < C; i += 256) {
    p_sn[i] = s_next ? s_next[(int64_t)b * s_next_bs + i] : 0.f;
    p_dm[i] = demod ? demod[(int64_t)b * demod_bs + i] : 1.f;
    p_bs[i] = bias ? bias[i] : 0.f;
    p_w0[i] = wr ? wr[(int64_t)b * wr_bs + i] : 0.f;
    p_w1[i] = wr ? wr[(int64_t)b * wr_bs + C + i] : 0.f;
    p_w2[i] = wr ? wr[(int64_t)b * wr_bs + 2 * C + i] : 0.f;
  }
  __syncthreads();
  const float nw = (noise != nullptr && noise_w != nullptr) ? *noise_w : 0.f;
  float rs[VEC], rd[VEC], r0[VEC], r1[VEC], r2[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) rs[k] = rd[k] = r0[k] = r1[k] = r2[k] = 0.f;
  const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, HW);
  if (slot < slots) {
    constexpr int U = 4;   // pixels in flight per thread: every load of the batch is issued before the first use
    for (int pb = p0 + slot; pb < p1; pb += U * slots) {
      V yv[U], gx[U];
      float gr[U][3], nzv[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = pb + u * slots;
        ok[u] = p < p1;
        const int pc = ok[u] ? p : p0;
        const int64_t off = ((int64_t)b * HW + pc) * C + c;
        yv[u] = *reinterpret_cast<const V*>(y + off);
        if (g_xn != nullptr) gx[u] = *reinterpret_cast<const V*>(g_xn + off);
        gr[u][0] = gr[u][1] = gr[u][2] = 0.f;
        if (g_rgb != nullptr) {
          gr[u][0] = g_rgb[((int64_t)b * 3 + 0) * HW + pc];
          gr[u][1] = g_rgb[((int64_t)b * 3 + 1) * HW + pc];
          gr[u][2] = g_rgb[((int64_t)b * 3 + 2) * HW + pc];
        }
        nzv[u] = noise != nullptr ? nw * noise[(int64_t)b * noise_bs + pc] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        const int p = pb + u * slots;
        const int64_t off = ((int64_t)b * HW + p) * C + c;
        const float gr0 = gr[u][0], gr1 = gr[u][1], gr2 = gr[u][2], nz = nzv[u];
        V go;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float yk = to_f32<T>(yv[u].v[k]);
          const float gxk = g_xn != nullptr ? to_f32<T>(gx[u].v[k]) : 0.f;
          const float gy = p_sn[c + k] * gxk + p_w0[c + k] * gr0 + p_w1[c + k] * gr1 + p_w2[c + k] * gr2;
          const float gv = gy * (yk > 0.f ? kSqrt2 : 0.2f * kSqrt2);
          const float v = yk > 0.f ? yk * (1.f / kSqrt2) : yk * (1.f / (0.2f * kSqrt2));
          rd[k] = fmaf(gv, v - nz - p_bs[c + k], rd[k]);
          rs[k] = fmaf(gxk, yk, rs[k]);
          r0[k] = fmaf(gr0, yk, r0[k]);
          r1[k] = fmaf(gr1, yk, r1[k]);
          r2[k] = fmaf(gr2, yk, r2[k]);
          go.v[k] = from_f32<T>(p_dm[c + k] * gv);
        }
        *reinterpret_cast<V*>(g_out + off) = go;
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      atomicAdd(&red[0 * C + c + k], rs[k]);
      atomicAdd(&red[1 * C + c + k], rd[k]);
      atomicAdd(&red[2 * C + c + k], r0[k]);
      atomicAdd(&red[3 * C + c + k], r1[k]);
      atomicAdd(&red[4 * C + c + k], r2[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    if (R_s != nullptr) atomicAdd(R_s + (int64_t)b * R_bs + i, red[i]);
    if (R_d != nullptr) atomicAdd(R_d + (int64_t)b * R_bs + i, red[C + i]);
    if (R_rgb != nullptr) {
      atomicAdd(R_rgb + (int64_t)b * R_rgb_bs + i, red[2 * C + i]);
      atomicAdd(R_rgb + (int64_t)b * R_rgb_bs + C + i, red[3 * C + i]);
      atomicAdd(R_rgb + (int64_t)b * R_rgb_bs + 2 * C + i, red[4 * C + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// act_bwd, bulk-async variant for the 16-bit path: same maths as act_bwd_kernel, but the two streamed tensors (saved
// activation y and incoming gradient g_x~') are fetched as contiguous 16 KB tiles with cp.async.bulk into a
// double-buffered shared-memory ring, so a CTA always has the next tile in flight and the arithmetic reads shared
// memory.  (The register-window version was load-latency bound: ncu long-scoreboard on the first use of every load.)
// A tile = P = 8192 / C pixels x C channels; every thread owns one 16-byte channel vector of 4 pixels of the tile.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bw_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bw_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(bw_smem_u32(bar))
               : "memory");
}

constexpr int kBwTileBytes = 16384;

__global__ void __launch_bounds__(256, 2)
act_bwd_bulk_kernel(__nv_bfloat16* __restrict__ g_out, const __nv_bfloat16* __restrict__ g_xn, const float* __restrict__ g_rgb,
                    const __nv_bfloat16* __restrict__ y, const float* __restrict__ s_next, int64_t s_next_bs,
                    const float* __restrict__ wr, int64_t wr_bs, const float* __restrict__ demod, int64_t demod_bs,
                    const float* __restrict__ noise, int64_t noise_bs, const float* __restrict__ noise_w,
                    const float* __restrict__ bias, float* __restrict__ R_s, float* __restrict__ R_d, int64_t R_bs,
                    float* __restrict__ R_rgb, int64_t R_rgb_bs, int HW, int C, int tiles_per_cta) {
  using T = __nv_bfloat16;
  using V = GVec<T, 8>;
  constexpr float kSqrt2 = 1.4142135623730951f;
  extern __shared__ __align__(128) uint8_t bw_smem[];
  uint8_t* ybuf = bw_smem;                                   // [2][16 KB]
  uint8_t* gbuf = bw_smem + 2 * kBwTileBytes;                // [2][16 KB]
  float* red = reinterpret_cast<float*>(bw_smem + 4 * kBwTileBytes);   // [5][C]
  float* prm = red + 5 * C;                                  // [6][C]
  float *p_sn = prm, *p_dm = prm + C, *p_bs = prm + 2 * C, *p_w0 = prm + 3 * C, *p_w1 = prm + 4 * C, *p_w2 = prm + 5 * C;
  __shared__ __align__(8) uint64_t bar[2];
  const int b = blockIdx.y;
  const int P = kBwTileBytes / (C * 2);                      // pixels per tile
  const int CV = C / 8, slots = 256 / CV;                    // P == 4 * slots
  const int cv = threadIdx.x % CV, slot = threadIdx.x / CV;
  const int c = cv * 8;
  const int tile0 = blockIdx.x * tiles_per_cta;
  const int ntiles = min(tiles_per_cta, (HW + P - 1) / P - tile0);
  const bool has_g = g_xn != nullptr;

  auto issue = [&](int i) {   // thread 0: tile i of this CTA into stage i & 1
    const int p0 = (tile0 + i) * P;
    const uint32_t bytes = (uint32_t)(min(P, HW - p0) * C * 2);
    const int st = i & 1;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bw_smem_u32(&bar[st])), "r"(has_g ? 2 * bytes : bytes) : "memory");
    bw_bulk_load(ybuf + st * kBwTileBytes, y + ((int64_t)b * HW + p0) * C, bytes, &bar[st]);
    if (has_g) bw_bulk_load(gbuf + st * kBwTileBytes, g_xn + ((int64_t)b * HW + p0) * C, bytes, &bar[st]);
  };

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bw_smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bw_smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (ntiles > 0) issue(0);
  }
  for (int i = threadIdx.x; i < 5 * C; i += 256) red[i] = 0.f;
  for (int i = threadIdx.x; i < C; i += 256) {
    p_sn[i] = s_next ? s_next[(int64_t)b * s_next_bs + i] : 0.f;
    p_dm[i] = demod ? demod[(int64_t)b * demod_bs + i] : 1.f;
    p_bs[i] = bias ? bias[i] : 0.f;
    p_w0[i] = wr ? wr[(int64_t)b * wr_bs + i] : 0.f;
    p_w1[i] = wr ? wr[(int64_t)b * wr_bs + C + i] : 0.f;
    p_w2[i] = wr ? wr[(int64_t)b * wr_bs + 2 * C + i] : 0.f;
  }
  __syncthreads();
  const float nw = (noise != nullptr && noise_w != nullptr) ? *noise_w : 0.f;
  float rs[8], rd[8], r0[8], r1[8], r2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) rs[k] = rd[k] = r0[k] = r1[k] = r2[k] = 0.f;

  for (int i = 0; i < ntiles; ++i) {
    if (threadIdx.x == 0 && i + 1 < ntiles) issue(i + 1);     // stage (i+1)&1 was released by the barrier ending iteration i-1
    const int st = i & 1;
    const int p0 = (tile0 + i) * P;
    // side inputs of this thread's 4 pixels (global, broadcast across the channel vectors), issued before the wait
    float gr[4][3], nzv[4];
    bool okp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + slot + u * slots;
      okp[u] = p < HW;
      const int pc = okp[u] ? p : p0;
      gr[u][0] = gr[u][1] = gr[u][2] = 0.f;
      if (g_rgb != nullptr) {
        gr[u][0] = g_rgb[((int64_t)b * 3 + 0) * HW + pc];
        gr[u][1] = g_rgb[((int64_t)b * 3 + 1) * HW + pc];
        gr[u][2] = g_rgb[((int64_t)b * 3 + 2) * HW + pc];
      }
      nzv[u] = noise != nullptr ? nw * noise[(int64_t)b * noise_bs + pc] : 0.f;
    }
    {
      const uint32_t parity = (uint32_t)((i >> 1) & 1);
      asm volatile("{\n.reg .pred P1;\nBWW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra BWD;\nbra BWW;\nBWD:\n}" ::"r"(
                       bw_smem_u32(&bar[st])),
                   "r"(parity)
                   : "memory");
    }
    const T* ys = reinterpret_cast<const T*>(ybuf + st * kBwTileBytes);
    const T* gs = reinterpret_cast<const T*>(gbuf + st * kBwTileBytes);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!okp[u]) continue;
      const int pl = slot + u * slots;
      const V yv = *reinterpret_cast<const V*>(ys + pl * C + c);
      V gx;
      if (has_g) gx = *reinterpret_cast<const V*>(gs + pl * C + c);
      const float gr0 = gr[u][0], gr1 = gr[u][1], gr2 = gr[u][2], nz = nzv[u];
      V go;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float yk = __bfloat162float(yv.v[k]);
        const float gxk = has_g ? __bfloat162float(gx.v[k]) : 0.f;
        const float gy = p_sn[c + k] * gxk + p_w0[c + k] * gr0 + p_w1[c + k] * gr1 + p_w2[c + k] * gr2;
        const float gv = gy * (yk > 0.f ? kSqrt2 : 0.2f * kSqrt2);
        const float v = yk > 0.f ? yk * (1.f / kSqrt2) : yk * (1.f / (0.2f * kSqrt2));
        rd[k] = fmaf(gv, v - nz - p_bs[c + k], rd[k]);
        rs[k] = fmaf(gxk, yk, rs[k]);
        r0[k] = fmaf(gr0, yk, r0[k]);
        r1[k] = fmaf(gr1, yk, r1[k]);
        r2[k] = fmaf(gr2, yk, r2[k]);
        go.v[k] = __float2bfloat16_rn(p_dm[c + k] * gv);
      }
      *reinterpret_cast<V*>(g_out + ((int64_t)b * HW + p0 + pl) * C + c) = go;
    }
    __syncthreads();   // every thread is done with stage st before it is refilled (by the issue of iteration i+1)
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    atomicAdd(&red[0 * C + c + k], rs[k]);
    atomicAdd(&red[1 * C + c + k], rd[k]);
    atomicAdd(&red[2 * C + c + k], r0[k]);
    atomicAdd(&red[3 * C + c + k], r1[k]);
    atomicAdd(&red[4 * C + c + k], r2[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) {
    if (R_s != nullptr) atomicAdd(R_s + (int64_t)b * R_bs + i, red[i]);
    if (R_d != nullptr) atomicAdd(R_d + (int64_t)b * R_bs + i, red[C + i]);
    if (R_rgb != nullptr) {
      atomicAdd(R_rgb + (int64_t)b * R_rgb_bs + i, red[2 * C + i]);
      atomicAdd(R_rgb + (int64_t)b * R_rgb_bs + C + i, red[3 * C + i]);
      atomicAdd(R_rgb + (int64_t)b * R_rgb_bs + 2 * C + i, red[4 * C + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// blur_bwd: g_t[u][v] = sum_{i,j} f[i] f[j] * g_v[u-i+1][v-j+1]  on the padded (TH x TW) grid of the
// up-conv output; writes d * g_t and reduces R_d += g_t * t_saved.
// ------------------------------------------------------------------------------------------------
template <typename T, typename TIN, int VEC>
__global__ void __launch_bounds__(256)
blur_bwd_kernel(T* __restrict__ g_acc, const T* __restrict__ g_v, const TIN* __restrict__ t_saved,
                const float* __restrict__ demod, int64_t demod_bs, float* __restrict__ R_d, int64_t R_bs, int OH, int OW,
                int TH, int TW, int C, int chunk, float f0, float f1, float f2, float f3) {
  using V = GVec<T, VEC>;
  using VIN = GVec<TIN, VEC>;
  extern __shared__ float red[];  // [C]
  const int CV = C / VEC;
  const int slots = 256 / CV;
  const int cv = threadIdx.x % CV, slot = threadIdx.x / CV;
  const int b = blockIdx.y, c = cv * VEC;
  const float f[4] = {f0, f1, f2, f3};
  for (int i = threadIdx.x; i < C; i += 256) red[i] = 0.f;
  __syncthreads();
  float dm[VEC], rd[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { dm[k] = demod[(int64_t)b * demod_bs + c + k]; rd[k] = 0.f; }
  const int total = TH * TW;
  const int p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, total);
  if (slot < slots) {
    for (int p = p0 + slot; p < p1; p += slots) {
      const int u = p / TW, v = p - u * TW;
      float acc[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
      if (u < TH - 1 && v < TW - 1) {  // the last padded row / column of t is structurally zero: no gradient
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int Y = u - i + 1;
          if (Y < 0 || Y >= OH) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int X = v - j + 1;
            if (X < 0 || X >= OW) continue;
            const V gv = *reinterpret_cast<const V*>(g_v + (((int64_t)b * OH + Y) * OW + X) * C + c);
            const float w = f[i] * f[j];
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] = fmaf(w, to_f32<T>(gv.v[k]), acc[k]);
          }
        }
      }
      const int64_t off = ((int64_t)b * total + p) * C + c;
      const VIN tv = *reinterpret_cast<const VIN*>(t_saved + off);
      V go;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        rd[k] = fmaf(acc[k], to_f32<TIN>(tv.v[k]), rd[k]);
        go.v[k] = from_f32<T>(dm[k] * acc[k]);
      }
      *reinterpret_cast<V*>(g_acc + off) = go;
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) atomicAdd(&red[c + k], rd[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) atomicAdd(R_d + (int64_t)b * R_bs + i, red[i]);
}

// R[b][ci] = sum_p g[b][p][ci] * cst[ci][p]      (gradient of conv1's style through the constant input)
template <typename T>
__global__ void const_grad_kernel(float* __restrict__ R, int64_t R_bs, const T* __restrict__ g, const float* __restrict__ cst,
                                  int B, int C, int HW) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int b = idx / C, c = idx % C;
  float acc = 0.f;
  for (int p = 0; p < HW; ++p) acc = fmaf(to_f32<T>(g[((int64_t)b * HW + p) * C + c]), cst[(int64_t)c * HW + p], acc);
  R[(int64_t)b * R_bs + c] = acc;
}

// gs[b][ci] = Rs[b][ci] - s[b][ci] * sum_co wsq[co][ci] * d[b][co]^2 * Rd[b][co]      (one warp per (b, ci))
__global__ void __launch_bounds__(256)
conv_style_grad_kernel(float* __restrict__ gs, int64_t gs_bs, const float* __restrict__ Rs, int64_t Rs_bs,
                       const float* __restrict__ s, int64_t s_bs, const float* __restrict__ wsq,
                       const float* __restrict__ d, const float* __restrict__ Rd, int64_t d_bs, int B, int Cin, int Cout) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * Cin) return;
  const int b = warp / Cin, ci = warp % Cin;
  float acc = 0.f;
  for (int co = lane; co < Cout; co += 32) {
    const float dd = d[(int64_t)b * d_bs + co];
    acc = fmaf(wsq[(int64_t)co * Cin + ci] * dd * dd, Rd[(int64_t)b * d_bs + co], acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) gs[(int64_t)b * gs_bs + ci] = Rs[(int64_t)b * Rs_bs + ci] - s[(int64_t)b * s_bs + ci] * acc;
}

// gs[b][ci] = sum_c wrgb[c][ci] * Rrgb[b][c][ci]
__global__ void rgb_style_grad_kernel(float* __restrict__ gs, int64_t gs_bs, const float* __restrict__ wrgb,
                                      const float* __restrict__ Rrgb, int64_t R_bs, int B, int Cin) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Cin) return;
  const int b = idx / Cin, ci = idx % Cin;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) acc = fmaf(wrgb[c * Cin + ci], Rrgb[(int64_t)b * R_bs + c * Cin + ci], acc);
  gs[(int64_t)b * gs_bs + ci] = acc;
}

// g_latent[b][i][k] = sum over the style rows r that read latent i of gs[b][r] * mod_w[r][k]
__global__ void __launch_bounds__(128)
latent_grad_kernel(float* __restrict__ g_latent, const float* __restrict__ gs, int64_t gs_bs,
                   const float* __restrict__ mod_w, const int* __restrict__ lat_seg, int n_latent, int D) {
  const int i = blockIdx.x, b = blockIdx.y;
  const int* seg = lat_seg + i * 7;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float acc = 0.f;
    for (int q = 0; q < seg[0]; ++q) {
      const int r0 = seg[1 + 2 * q], n = seg[2 + 2 * q];
      for (int r = r0; r < r0 + n; ++r) acc = fmaf(gs[(int64_t)b * gs_bs + r], mod_w[(int64_t)r * D + k], acc);
    }
    g_latent[((int64_t)b * n_latent + i) * D + k] = acc;
  }
}

// dst [tap][Cout][Cin] fp32 = scale * src[Cout][Cin][tap]
__global__ void pack_weight_t_kernel(float* __restrict__ dst, const float* __restrict__ src, int Cout, int Cin, int ntap,
                                     float scale) {
  const int64_t total = (int64_t)Cout * Cin;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x)
    for (int t = 0; t < ntap; ++t) dst[(int64_t)t * total + idx] = src[idx * ntap + t] * scale;
}

// dst [tap][Cin][Cout] bf16 = scale * src[Cout][Cin][tap]   (K-major B operand of the data-gradient GEMM)
__global__ void pack_weight_t_bf16_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, int Cout, int Cin,
                                          int ntap, float scale) {
  const int64_t total = (int64_t)Cout * Cin;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % Cin), co = (int)(idx / Cin);
    for (int t = 0; t < ntap; ++t) dst[((int64_t)t * Cin + ci) * Cout + co] = __float2bfloat16_rn(src[idx * ntap + t] * scale);
  }
}

namespace {

// data-gradient conv: tcgen05 kernel for bf16 (raw epilogue, no demodulation), CUDA-core kernel otherwise
template <typename T>
int run_dgrad_conv(l2i_generator* g, const StyledConvLayer& L, const void* in, const ConvGeom& geom, const EpiParams& e,
                   cudaStream_t st) {
  if (sizeof(T) == 2 && g->conv_impl != 1 && L.w_bf16_t != nullptr && conv_tc_supported(geom, e))
    return launch_conv_tc(in, L.w_bf16_t, geom, e, st);
  return launch_conv_simt<T>(in, L.w_f32_t, geom, e, st);
}

template <typename T>
int launch_act_bwd(void* g_out, const void* g_xn, const float* g_rgb, const void* y, const float* s_next, int64_t s_next_bs,
                   const float* wr, int64_t wr_bs, const float* demod, int64_t demod_bs, const float* noise, int64_t noise_bs,
                   const float* noise_w, const float* bias, float* R_s, float* R_d, int64_t R_bs, float* R_rgb,
                   int64_t R_rgb_bs, int B, int HW, int C, cudaStream_t st) {
  if (sizeof(T) == 2 && !g_switches.fir_simt && C >= 32 && C <= 512 && (C & (C - 1)) == 0) {
    // 16 KB tiles: P = 8192 / C pixels; 16 tiles per CTA keep the double-buffered ring busy
    const int P = kBwTileBytes / (C * 2);
    const int ntiles = ceil_div(HW, P);
    const int per_cta = std::min(ntiles, 16);
    dim3 grid(ceil_div(ntiles, per_cta), B);
    const size_t smem = 4 * kBwTileBytes + sizeof(float) * 11 * C;
    static bool attr_set = false;
    if (!attr_set) {
      L2I_CUDA_TRY(cudaFuncSetAttribute(act_bwd_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kBwTileBytes + 4 * 11 * 512));
      attr_set = true;
    }
    act_bwd_bulk_kernel<<<grid, 256, smem, st>>>((__nv_bfloat16*)g_out, (const __nv_bfloat16*)g_xn, g_rgb, (const __nv_bfloat16*)y, s_next,
                                                 s_next_bs, wr, wr_bs, demod, demod_bs, noise, noise_bs, noise_w, bias, R_s, R_d, R_bs,
                                                 R_rgb, R_rgb_bs, HW, C, per_cta);
    return check_launch("act_bwd_bulk");
  }
  constexpr int VEC = 4;   // 8-byte (bf16) / 16-byte (fp32) vectors: small register arrays leave room for 4 pixels in flight
  if (C % VEC != 0 || C / VEC > 256) { set_error("act_bwd: unsupported C=%d", C); return L2I_ERR_UNSUPPORTED; }
  const int chunk = std::max(256, std::min(HW, 4096));
  dim3 grid(ceil_div(HW, chunk), B);
  act_bwd_kernel<T, VEC><<<grid, 256, sizeof(float) * 11 * C, st>>>(
      (T*)g_out, (const T*)g_xn, g_rgb, (const T*)y, s_next, s_next_bs, wr, wr_bs, demod, demod_bs, noise, noise_bs, noise_w,
      bias, R_s, R_d, R_bs, R_rgb, R_rgb_bs, HW, C, chunk);
  return check_launch("act_bwd");
}

template <typename T, typename TIN>
int launch_blur_bwd(void* g_acc, const void* g_v, const void* t_saved, const float* demod, int64_t demod_bs, float* R_d,
                    int64_t R_bs, int B, int OH, int OW, int TH, int TW, int C, const float* f, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  if (C % VEC != 0 || C / VEC > 256) { set_error("blur_bwd: unsupported C=%d", C); return L2I_ERR_UNSUPPORTED; }
  const int total = TH * TW;
  const int chunk = std::max(256, std::min(total, 4096));
  dim3 grid(ceil_div(total, chunk), B);
  blur_bwd_kernel<T, TIN, VEC><<<grid, 256, sizeof(float) * C, st>>>((T*)g_acc, (const T*)g_v, (const TIN*)t_saved, demod,
                                                                    demod_bs, R_d, R_bs, OH, OW, TH, TW, C, chunk, f[0], f[1],
                                                                    f[2], f[3]);
  return check_launch("blur_bwd");
}

TapList dgrad_plain_taps() {
  TapList t{};
  t.n = 9;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      const int i = kh * 3 + kw;
      t.dy[i] = (int8_t)(1 - kh);
      t.dx[i] = (int8_t)(1 - kw);
      t.wtap[i] = (int8_t)i;
    }
  return t;
}

TapList dgrad_up_taps() {
  TapList t{};
  t.n = 9;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      const int i = kh * 3 + kw;
      t.dy[i] = (int8_t)kh;
      t.dx[i] = (int8_t)kw;
      t.wtap[i] = (int8_t)i;
    }
  return t;
}

template <typename T>
int backward_impl(l2i_generator* g, float* grad_latent, const float* grad_image, int B, cudaStream_t st) {
  using TIN = typename std::conditional<sizeof(T) == 2, __half, float>::type;
  const int D = g->D;
  const int64_t R_bs = g->d_rows, Rr_bs = g->wr_elems;
  L2I_CUDA_TRY(cudaMemsetAsync(g->R_s, 0, sizeof(float) * B * R_bs, st));
  L2I_CUDA_TRY(cudaMemsetAsync(g->R_d, 0, sizeof(float) * B * R_bs, st));
  L2I_CUDA_TRY(cudaMemsetAsync(g->R_rgb, 0, sizeof(float) * B * Rr_bs, st));

  const float* g_skip = grad_image;  // gradient w.r.t. the skip image at the current resolution
  const void* g_xn = nullptr;        // gradient w.r.t. the x~' output of the layer being processed
  int gx_sel = 0;                    // act[0] / act[1] hold the data gradients (forward scratch is free now)
  int gskip_sel = 0;
  int rgb_i = (int)g->rgbs.size() - 1;
  for (int li = (int)g->convs.size() - 1; li >= 0; --li) {
    const auto& L = g->convs[li];
    const StyledConvLayer* next = li + 1 < (int)g->convs.size() ? &g->convs[li + 1] : nullptr;
    const float* s_next = next ? g->s_all + next->s_off : nullptr;
    const float* demod = g->d_all + L.d_off;
    const int HWo = L.res_out * L.res_out;
    ConvGeom geom{};
    geom.B = B; geom.Cin = L.cout; geom.Cout = L.cin; geom.nphase = 1; geom.out_scale = 1; geom.weight_taps = 9;
    geom.OH = geom.OW = L.res_in; geom.out_H = geom.out_W = L.res_in;
    EpiParams e{};
    e.mode = 1; e.demod = nullptr;
    void* gx_dst = g->act[gx_sel];
    e.out = gx_dst;
    if (!L.up) {
      const auto& R = g->rgbs[rgb_i];
      L2I_TRY(launch_act_bwd<T>(g->gbuf, g_xn, g_skip, L.y_save, s_next, g->s_rows, g->wr_all + R.wr_off, g->wr_elems, demod,
                                g->d_rows, L.noise_ptr, L.noise_bs, P_noise_w(g, L), P_bias(g, L),
                                next ? g->R_s + L.d_off : nullptr, g->R_d + L.d_off, R_bs, g->R_rgb + R.wr_off, Rr_bs, B, HWo,
                                L.cout, st));
      geom.H = geom.W = L.res_out; geom.in_scale = 1; geom.taps[0] = dgrad_plain_taps();
      L2I_TRY(run_dgrad_conv<T>(g, L, g->gbuf, geom, e, st));
      // skip chain: g_skip(res/2) = Upsample^T(g_skip(res)) = upfirdn2d(down=2, flipped kernel, pad (1,1))
      if (rgb_i > 0) {
        float* dst = g->gskip[gskip_sel];
        L2I_TRY(l2i_upfirdn2d(dst, g_skip, g->fir2d_dev, (int64_t)B * 3, L.res_out, L.res_out, 1, 4, 4, 1, 1, 2, 2, 1, 1, 1, 1,
                              L2I_F32, st));
        g_skip = dst;
        gskip_sel ^= 1;
      }
      --rgb_i;
    } else {
      // y -> g_v (no ToRGB, no demod here), then blur^T and the stride-2 gather conv
      L2I_TRY(launch_act_bwd<T>(g->gbuf, g_xn, nullptr, L.y_save, s_next, g->s_rows, nullptr, 0, nullptr, 0, nullptr, 0, nullptr,
                                nullptr, g->R_s + L.d_off, nullptr, R_bs, nullptr, 0, B, HWo, L.cout, st));
      const int TH = 2 * L.res_in + 2;
      if (sizeof(T) == 2 && fir_tma_supported(L.cout))
        L2I_TRY(launch_blur_bwd_tma(g->tbuf, g->gbuf, L.t_save, demod, g->d_rows, g->R_d + L.d_off, R_bs, B, L.res_out, L.res_out, TH, TH,
                                    L.cout, g->fir, st));
      else
        L2I_TRY((launch_blur_bwd<T, TIN>(g->tbuf, g->gbuf, L.t_save, demod, g->d_rows, g->R_d + L.d_off, R_bs, B, L.res_out,
                                         L.res_out, TH, TH, L.cout, g->fir, st)));
      geom.H = geom.W = TH; geom.in_scale = 2; geom.taps[0] = dgrad_up_taps();
      L2I_TRY(run_dgrad_conv<T>(g, L, g->tbuf, geom, e, st));
    }
    g_xn = gx_dst;
    gx_sel ^= 1;
  }
  // conv1's input is the constant: its style gradient is sum_p g_x~ * const
  const auto& L0 = g->convs[0];
  const_grad_kernel<T><<<ceil_div(B * L0.cin, 256), 256, 0, st>>>(g->R_s0, L0.cin, (const T*)g_xn, P(g, "input.input"), B,
                                                                 L0.cin, 16);
  L2I_TRY(check_launch("const_grad"));

  // style gradients per modulated conv, then through the modulation linears to the W+ latent
  for (size_t li = 0; li < g->convs.size(); ++li) {
    const auto& L = g->convs[li];
    const float* Rs = li == 0 ? g->R_s0 : g->R_s + g->convs[li - 1].d_off;
    const int64_t Rs_bs = li == 0 ? L.cin : R_bs;
    conv_style_grad_kernel<<<ceil_div(B * L.cin * 32, 256), 256, 0, st>>>(g->gs_all + L.s_off, g->s_rows, Rs, Rs_bs,
                                                                         g->s_all + L.s_off, g->s_rows, g->wsq_all + L.wsq_off,
                                                                         g->d_all + L.d_off, g->R_d + L.d_off, g->d_rows, B,
                                                                         L.cin, L.cout);
    L2I_TRY(check_launch("conv_style_grad"));
  }
  for (auto& R : g->rgbs) {
    rgb_style_grad_kernel<<<ceil_div(B * R.cin, 256), 256, 0, st>>>(g->gs_all + R.s_off, g->s_rows, g->wrgb_all + R.wr_off,
                                                                   g->R_rgb + R.wr_off, Rr_bs, B, R.cin);
    L2I_TRY(check_launch("rgb_style_grad"));
  }
  latent_grad_kernel<<<dim3(g->n_latent, B), 128, 0, st>>>(grad_latent, g->gs_all, g->s_rows, g->mod_w_all, g->lat_seg,
                                                           g->n_latent, D);
  return check_launch("latent_grad");
}

}  // namespace
}  // namespace l2i

using namespace l2i;

extern "C" int l2i_generator_set_training(l2i_generator_t* g, int enable) {
  L2I_REQUIRE(g, "generator_set_training: null generator");
  if (!enable) { g->training = false; return L2I_OK; }
  if (!g->finalized) { set_error("generator_set_training: call l2i_generator_finalize first"); return L2I_ERR_STATE; }
  if (!g->train_buffers) {
    const int64_t B = g->max_batch;
    const int64_t es = (int64_t)g->elem_size();
    int64_t act_elems = 0;
    for (auto& L : g->convs) {
      char* p = nullptr;
      L2I_TRY(train_alloc(g, &p, B * L.res_out * L.res_out * (int64_t)L.cout * es));
      L.y_save = p;
      if (L.up) {
        const int64_t th = 2 * L.res_in + 2;
        L2I_TRY(train_alloc(g, &p, B * th * th * (int64_t)L.cout * es));
        L.t_save = p;
      }
      L2I_TRY(train_alloc(g, &L.w_f32_t, (int64_t)9 * L.cin * L.cout));
      if (g->dtype == L2I_BF16) L2I_TRY(train_alloc(g, &L.w_bf16_t, (int64_t)9 * L.cin * L.cout));
      act_elems = std::max(act_elems, B * L.res_out * L.res_out * (int64_t)L.cout);
    }
    char* gb = nullptr;
    L2I_TRY(train_alloc(g, &gb, act_elems * es));
    g->gbuf = gb;
    L2I_TRY(train_alloc(g, &g->R_s, B * (int64_t)g->d_rows));
    L2I_TRY(train_alloc(g, &g->R_d, B * (int64_t)g->d_rows));
    L2I_TRY(train_alloc(g, &g->R_rgb, B * (int64_t)g->wr_elems));
    L2I_TRY(train_alloc(g, &g->R_s0, B * (int64_t)g->convs[0].cin));
    L2I_TRY(train_alloc(g, &g->gs_all, B * (int64_t)g->s_rows));
    L2I_TRY(train_alloc(g, &g->gskip[0], B * 3 * (int64_t)(g->size / 2) * (g->size / 2)));
    L2I_TRY(train_alloc(g, &g->gskip[1], B * 3 * (int64_t)(g->size / 4) * (g->size / 4)));
    L2I_TRY(train_alloc(g, &g->fir2d_dev, 16));
    L2I_TRY(train_alloc(g, &g->lat_seg, (int64_t)g->n_latent * 7));
    // tables: which style rows read each latent index; the 2-D FIR of the skip up-sampling (flipped for the transpose)
    std::vector<int> seg(g->n_latent * 7, 0);
    auto add = [&](int lat, int r0, int n) {
      int* s = &seg[lat * 7];
      s[1 + 2 * s[0]] = r0; s[2 + 2 * s[0]] = n; ++s[0];
    };
    for (auto& L : g->convs) add(L.latent_idx, L.s_off, L.cin);
    for (auto& R : g->rgbs) add(R.latent_idx, R.s_off, R.cin);
    float k2[16];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) k2[i * 4 + j] = g->fir[i] * g->fir[j];  // fir is already flipped: this is flip(kernel)
    L2I_CUDA_TRY(cudaMemcpy(g->lat_seg, seg.data(), sizeof(int) * seg.size(), cudaMemcpyHostToDevice));
    L2I_CUDA_TRY(cudaMemcpy(g->fir2d_dev, k2, sizeof(k2), cudaMemcpyHostToDevice));
    g->train_buffers = true;
  }
  // data-gradient weight copies follow the current parameters (repacked only after a finalize)
  if (g->train_weights_packed) { g->training = true; return L2I_OK; }
  for (auto& L : g->convs) {
    const float scale = 1.0f / std::sqrt((float)(L.cin * 9));
    const int64_t total = (int64_t)L.cout * L.cin;
    const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 8);
    pack_weight_t_kernel<<<blocks, 256>>>(L.w_f32_t, g->params.at(L.name + ".conv.weight").ptr, L.cout, L.cin, 9, scale);
    L2I_TRY(check_launch("pack_weight_t"));
    if (L.w_bf16_t != nullptr) {
      pack_weight_t_bf16_kernel<<<blocks, 256>>>(L.w_bf16_t, g->params.at(L.name + ".conv.weight").ptr, L.cout, L.cin, 9, scale);
      L2I_TRY(check_launch("pack_weight_t_bf16"));
    }
  }
  L2I_CUDA_TRY(cudaDeviceSynchronize());
  g->train_weights_packed = true;
  g->training = true;
  return L2I_OK;
}

extern "C" int l2i_generator_backward(l2i_generator_t* g, float* grad_latent, const float* grad_image, int batch, void* stream) {
  L2I_REQUIRE(g && grad_latent && grad_image, "generator_backward: null argument");
  if (!g->train_buffers || g->last_train_batch != batch) {
    set_error("generator_backward: needs a forward in training mode with the same batch (last training batch %d, got %d)",
              g->last_train_batch, batch);
    return L2I_ERR_STATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (g->dtype == L2I_F32) return backward_impl<float>(g, grad_latent, grad_image, batch, st);
  return backward_impl<__nv_bfloat16>(g, grad_latent, grad_image, batch, st);
}
