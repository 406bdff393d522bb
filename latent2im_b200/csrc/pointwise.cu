// Bandwidth-bound NHWC kernels of the synthesis network:
//   * blur_act      : 4x4 FIR blur of the up-conv output (Blur, networks.py:72-88, pad (1,1)) fused
//                     with NoiseInjection (:275-286), FusedLeakyReLU (bias + lrelu*sqrt2) and the
//                     next layer's style scale - one read of t, one write of x~.
//   * skip_combine  : ToRGB tail (:349-358): sum of rgb partials + bias + 2x FIR upsample of the skip.
//   * const_input   : ConstantInput (:289-299) * style of conv1.
//   * style_finish  : demodulation coefficients and ToRGB effective weights from the styles.
//   * weight packing kernels and the uint8 image epilogue (transform_base.py:625-626).
#include "conv_common.cuh"

namespace l2i {

// ------------------------------------------------------------------------------------------------
// blur + noise + bias + lrelu + next-style scale.  t: [B][TH][TW][C] (TH = 2H+2 allocated, rows
// 0..2H valid, row 2H+1 zero), out: [B][OH][OW][C] with OH = TH-2.  out[Y][X] = sum_{i,j} f[i] f[j] *
// t[Y+i-1][X+j-1] (zero outside).  Each thread owns VEC consecutive channels of one output pixel and
// walks RY consecutive rows so vertical neighbours are reused from registers.
// ------------------------------------------------------------------------------------------------
template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) TVec { T v[VEC]; };

// Thread mapping: a CTA of 256 threads covers XT = 256 / VP adjacent output columns x one chunk of
// VP 16-byte channel vectors, and marches down RY output rows.  Each thread keeps the horizontally
// filtered rows Y-1 .. Y+2 of its column in registers (separable FIR), so every t element is fetched
// by four neighbouring threads of the same CTA (one DRAM/L2 fetch + three L1 hits) and every output
// element is written once with a 16-byte store.  Per-thread bias / next-style values live in registers.
constexpr int kBlurRows = 32;

template <typename T, typename TIN, int VEC>
__global__ void __launch_bounds__(256, 4)
blur_act_kernel(T* __restrict__ out, T* __restrict__ y_out, const TIN* __restrict__ t, int B, int OH, int OW, int C, int TH, int TW,
                const float* __restrict__ noise, int64_t noise_bs, const float* __restrict__ noise_w,
                const float* __restrict__ bias, const float* __restrict__ s_next, int64_t s_next_bs,
                float f0, float f1, float f2, float f3, int VP, int tiles_x, int tiles_y, int chunks, int pair_pack) {
  using V = TVec<T, VEC>;
  using VIN = TVec<TIN, VEC>;
  constexpr float kSqrt2 = 1.4142135623730951f;
  const int XT = 256 / VP;
  int r = blockIdx.x;
  const int ck = r % chunks; r /= chunks;
  const int tx = r % tiles_x; r /= tiles_x;
  const int ty = r % tiles_y;
  const int b = r / tiles_y;
  const int X = tx * XT + threadIdx.x / VP;
  const int c = (ck * VP + threadIdx.x % VP) * VEC;
  if (X >= OW) return;
  const int Y0 = ty * kBlurRows;
  const int Y1 = min(Y0 + kBlurRows, OH);
  const float nw = (noise != nullptr && noise_w != nullptr) ? __ldg(noise_w) * kSqrt2 : 0.f;
  float bs[VEC], sn[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    bs[k] = __ldg(bias + c + k) * kSqrt2;
    sn[k] = s_next != nullptr ? __ldg(s_next + (int64_t)b * s_next_bs + c + k) : 1.f;
  }
  const TIN* tb = t + (int64_t)b * TH * TW * C + c;
  const float f[4] = {f0, f1, f2, f3};

  auto hrow = [&](int u, float (&h)[VEC]) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) h[k] = 0.f;
    if (u < 0 || u >= TH) return;
    const TIN* row = tb + (int64_t)u * TW * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int v = X + j - 1;
      if (v < 0 || v >= TW) continue;
      const VIN tv = *reinterpret_cast<const VIN*>(row + (int64_t)v * C);
#pragma unroll
      for (int k = 0; k < VEC; ++k) h[k] = fmaf(f[j], to_f32<TIN>(tv.v[k]), h[k]);
    }
  };

  float h0[VEC], h1[VEC], h2[VEC], h3[VEC];
  hrow(Y0 - 1, h0);
  hrow(Y0, h1);
  hrow(Y0 + 1, h2);
  const float* nrow = noise != nullptr ? noise + (int64_t)b * noise_bs + X : nullptr;
  T* ob = out + ((int64_t)b * OH * OW + X) * C + c;
#pragma unroll 4
  for (int Y = Y0; Y < Y1; ++Y) {
    hrow(Y + 2, h3);
    const float nz = nrow != nullptr ? nw * __ldg(nrow + (int64_t)Y * OW) : 0.f;
    V ov, yv;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float a = f0 * h0[k];
      a = fmaf(f1, h1[k], a);
      a = fmaf(f2, h2[k], a);
      a = fmaf(f3, h3[k], a);
      float x = fmaf(a, kSqrt2, bs[k] + nz);
      x = fmaxf(x, 0.2f * x);
      yv.v[k] = from_f32<T>(x);
      ov.v[k] = from_f32<T>(x * sn[k]);
      h0[k] = h1[k]; h1[k] = h2[k]; h2[k] = h3[k];
    }
    if (pair_pack)  // [B][OH/2][OW][2][C]: the rows 2j, 2j+1 of a column sit side by side (128-byte units for C = 32)
      *reinterpret_cast<V*>(out + ((((int64_t)b * (OH >> 1) + (Y >> 1)) * OW + X) * 2 + (Y & 1)) * C + c) = ov;
    else
      *reinterpret_cast<V*>(ob + (int64_t)Y * OW * C) = ov;
    if (y_out != nullptr) *reinterpret_cast<V*>(y_out + ((int64_t)b * OH * OW + X) * C + c + (int64_t)Y * OW * C) = yv;
  }
}

template <typename T, typename TIN>
int launch_blur_act(void* out, void* y_out, const void* t, int B, int OH, int OW, int C, int TH, int TW, const float* noise,
                    int64_t noise_bs, const float* noise_w, const float* bias, const float* s_next,
                    int64_t s_next_bs, const float* f, int pair_pack, cudaStream_t st) {
  constexpr int VEC = 4;  // 4 channels per thread (8-byte bf16 / 16-byte fp32 accesses) keeps the register windows small
  if (C % VEC != 0) {
    set_error("blur_act: C=%d not a multiple of %d", C, VEC);
    return L2I_ERR_UNSUPPORTED;
  }
  if ((int64_t)B * OH * OW == 0) return L2I_OK;
  int VP = C / VEC;          // 16-byte vectors per pixel handled by one CTA pass
  int chunks = 1;
  if (VP > 8) {
    if (VP % 8 != 0) {
      set_error("blur_act: C=%d unsupported", C);
      return L2I_ERR_UNSUPPORTED;
    }
    chunks = VP / 8;
    VP = 8;
  }
  if (256 % VP != 0) {
    set_error("blur_act: C=%d unsupported", C);
    return L2I_ERR_UNSUPPORTED;
  }
  const int XT = 256 / VP;
  const int tiles_x = ceil_div(OW, XT), tiles_y = ceil_div(OH, kBlurRows);
  const int64_t blocks = (int64_t)B * tiles_x * tiles_y * chunks;
  if (blocks > 0x7fffffff) {
    set_error("blur_act: grid too large");
    return L2I_ERR_INVALID_ARG;
  }
  blur_act_kernel<T, TIN, VEC><<<(unsigned)blocks, 256, 0, st>>>((T*)out, (T*)y_out, (const TIN*)t, B, OH, OW, C, TH, TW, noise, noise_bs,
                                                           noise_w, bias, s_next, s_next_bs, f[0], f[1], f[2], f[3], VP,
                                                           tiles_x, tiles_y, chunks, pair_pack);
  return check_launch("blur_act");
}
template int launch_blur_act<float, float>(void*, void*, const void*, int, int, int, int, int, int, const float*, int64_t,
                                           const float*, const float*, const float*, int64_t, const float*, int, cudaStream_t);
template int launch_blur_act<__nv_bfloat16, __half>(void*, void*, const void*, int, int, int, int, int, int, const float*,
                                                    int64_t, const float*, const float*, const float*, int64_t,
                                                    const float*, int, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// skip_out[b,c,Y,X] = sum_p rgb_part[p][b][c][Y][X] + bias[c] + upsample2x(skip_in)[b,c,Y,X]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
skip_combine_kernel(float* __restrict__ skip_out, const float* __restrict__ rgb_part, int nparts,
                    const float* __restrict__ bias, const float* __restrict__ skip_in, int B, int H, int W,
                    float f0, float f1, float f2, float f3) {
  const int64_t plane = (int64_t)H * W;
  const int64_t total = (int64_t)B * 3 * plane;
  const float f[4] = {f0, f1, f2, f3};
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bc = idx / plane;
    const int64_t p = idx - bc * plane;
    const int c = (int)(bc % 3);
    const int Y = (int)(p / W), X = (int)(p - (int64_t)Y * W);
    float v = bias[c];
    for (int q = 0; q < nparts; ++q) v += rgb_part[(int64_t)q * total + idx];
    if (skip_in != nullptr) v += upsample2x_at(skip_in + bc * (plane / 4), H / 2, W / 2, Y, X, f);
    skip_out[idx] = v;
  }
}

int launch_skip_combine(float* skip_out, const float* rgb_part, int nparts, const float* bias,
                        const float* skip_in, int B, int H, int W, const float* f, cudaStream_t st) {
  const int64_t total = (int64_t)B * 3 * H * W;
  if (total == 0) return L2I_OK;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 16);
  skip_combine_kernel<<<blocks, 256, 0, st>>>(skip_out, rgb_part, nparts, bias, skip_in, B, H, W, f[0], f[1],
                                              f[2], f[3]);
  return check_launch("skip_combine");
}

// ------------------------------------------------------------------------------------------------
// x~0[b, y, x, c] = const[c, y, x] * s[b, c]
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void const_input_kernel(T* __restrict__ out, const float* __restrict__ cst, const float* __restrict__ s,
                                   int64_t s_bs, int B, int C, int HW) {
  const int64_t total = (int64_t)B * HW * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int p = (int)((idx / C) % HW);
    const int b = (int)(idx / ((int64_t)C * HW));
    out[idx] = from_f32<T>(cst[(int64_t)c * HW + p] * s[(int64_t)b * s_bs + c]);
  }
}

template <typename T>
int launch_const_input(void* out, const float* cst, const float* s, int64_t s_bs, int B, int C, int HW,
                       cudaStream_t st) {
  const int64_t total = (int64_t)B * HW * C;
  if (total == 0) return L2I_OK;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 8);
  const_input_kernel<T><<<blocks, 256, 0, st>>>((T*)out, cst, s, s_bs, B, C, HW);
  return check_launch("const_input");
}
template int launch_const_input<float>(void*, const float*, const float*, int64_t, int, int, int, cudaStream_t);
template int launch_const_input<__nv_bfloat16>(void*, const float*, const float*, int64_t, int, int, int, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// style_finish: one warp per row.
//   demod rows : d[b][row] = rsqrt(sum_ci s[b][s_off+ci]^2 * wsq[wsq_off + ci] + 1e-8)
//   rgb rows   : wr[b][row*cin + ci] ... handled by a plain elementwise kernel below
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
demod_kernel(float* __restrict__ d, int64_t d_bs, const float* __restrict__ s, int64_t s_bs,
             const float* __restrict__ wsq, const int64_t* __restrict__ row_wsq_off,
             const int* __restrict__ row_s_off, const int* __restrict__ row_cin, int nrows, int B) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nrows) return;
  const float* wrow = wsq + row_wsq_off[warp];
  const int cin = row_cin[warp], soff = row_s_off[warp];
  // four samples per pass share every load of the squared-weight row (one sample per pass re-read the 13 MB table B times: 58 us
  // at batch 32, L2-bound)
  for (int b0 = 4 * blockIdx.y; b0 < B; b0 += 4 * gridDim.y) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* sr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sr[j] = s + (int64_t)min(b0 + j, B - 1) * s_bs + soff;
    for (int k = lane; k < cin; k += 32) {
      const float w = wrow[k];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float v = sr[j][k]; acc[j] = fmaf(v * v, w, acc[j]); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
      if (lane == 0 && b0 + j < B) d[(int64_t)(b0 + j) * d_bs + warp] = rsqrtf(acc[j] + 1e-8f);
    }
  }
}

int launch_demod(float* d, int64_t d_bs, const float* s, int64_t s_bs, const float* wsq,
                 const int64_t* row_wsq_off, const int* row_s_off, const int* row_cin, int nrows, int B,
                 cudaStream_t st) {
  if (nrows == 0 || B == 0) return L2I_OK;
  dim3 grid(ceil_div(nrows * 32, 256), std::min(ceil_div(B, 4), 64));
  demod_kernel<<<grid, 256, 0, st>>>(d, d_bs, s, s_bs, wsq, row_wsq_off, row_s_off, row_cin, nrows, B);
  return check_launch("demod");
}

// wr[b][e] = wrgb[e] * s[b][elem_s_off[e]]   for e over all ToRGB layers' [3][Cin] weights
__global__ void __launch_bounds__(256)
rgb_weight_kernel(float* __restrict__ wr, int64_t wr_bs, const float* __restrict__ wrgb,
                  const int* __restrict__ elem_s_off, const float* __restrict__ s, int64_t s_bs, int n, int B) {
  const int64_t total = (int64_t)n * B;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx % n);
    const int b = (int)(idx / n);
    wr[(int64_t)b * wr_bs + e] = wrgb[e] * s[(int64_t)b * s_bs + elem_s_off[e]];
  }
}

int launch_rgb_weight(float* wr, int64_t wr_bs, const float* wrgb, const int* elem_s_off, const float* s,
                      int64_t s_bs, int n, int B, cudaStream_t st) {
  if (n == 0 || B == 0) return L2I_OK;
  const int blocks = (int)std::min<int64_t>(ceil_div64((int64_t)n * B, 256), (int64_t)kNumSMs * 8);
  rgb_weight_kernel<<<blocks, 256, 0, st>>>(wr, wr_bs, wrgb, elem_s_off, s, s_bs, n, B);
  return check_launch("rgb_weight");
}

// ------------------------------------------------------------------------------------------------
// gather latents with arbitrary strides into [B][n_latent][D]
// ------------------------------------------------------------------------------------------------
__global__ void gather_latent_kernel(float* __restrict__ out, const float* __restrict__ in, int64_t bs, int64_t ls,
                                     int B, int L, int D) {
  const int64_t total = (int64_t)B * L * D;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % D);
    const int i = (int)((idx / D) % L);
    const int b = (int)(idx / ((int64_t)D * L));
    out[idx] = in[(int64_t)b * bs + (int64_t)i * ls + k];
  }
}

int launch_gather_latent(float* out, const float* in, int64_t bs, int64_t ls, int B, int L, int D, cudaStream_t st) {
  const int64_t total = (int64_t)B * L * D;
  if (total == 0) return L2I_OK;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 8);
  gather_latent_kernel<<<blocks, 256, 0, st>>>(out, in, bs, ls, B, L, D);
  return check_launch("gather_latent");
}

// ------------------------------------------------------------------------------------------------
// weight packing (run once at finalize)
// ------------------------------------------------------------------------------------------------
// src: [Cout][Cin][KH][KW] fp32.  dst_f32: [tap][Cin][Cout] * scale; dst_bf16: [tap][Cout][Cin] * scale;
// wsq: [Cout][Cin] = sum_tap (scale*w)^2
__global__ void pack_conv_weight_kernel(float* __restrict__ dst_f32, __nv_bfloat16* __restrict__ dst_bf16,
                                        float* __restrict__ wsq, const float* __restrict__ src, int Cout, int Cin,
                                        int ntap, float scale, int split) {
  const int64_t total = (int64_t)Cout * Cin;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % Cin);
    const int co = (int)(idx / Cin);
    float ss = 0.f;
    for (int t = 0; t < ntap; ++t) {
      const float w = src[idx * ntap + t] * scale;
      ss = fmaf(w, w, ss);
      if (dst_f32) dst_f32[((int64_t)t * Cin + ci) * Cout + co] = w;
      if (dst_bf16) {
        const int64_t o = ((int64_t)t * Cout + co) * Cin + ci;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        dst_bf16[o] = hi;
        // split mode: second half of the tensor holds the rounding residual, so hi + lo carries ~16 significand bits
        if (split) dst_bf16[(int64_t)ntap * Cout * Cin + o] = __float2bfloat16_rn(w - __bfloat162float(hi));
      }
    }
    if (wsq) wsq[idx] = ss;
  }
}

int launch_pack_conv_weight(float* dst_f32, __nv_bfloat16* dst_bf16, float* wsq, const float* src, int Cout,
                            int Cin, int ntap, float scale, int split, cudaStream_t st) {
  const int64_t total = (int64_t)Cout * Cin;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 8);
  pack_conv_weight_kernel<<<blocks, 256, 0, st>>>(dst_f32, dst_bf16, wsq, src, Cout, Cin, ntap, scale, split);
  return check_launch("pack_conv_weight");
}

// Composite up-conv weights: conv_transpose2d(3x3, stride 2) followed by the 4x4 FIR blur (pad (1,1)) is one
// stride-2 transposed conv with a 6x6 kernel; per output phase (py, px) it is a 3x3 conv over the input pixels
// (y + dy, x + dx), dy, dx in {-1, 0, 1}:
//   Wc[py,px][dy,dx] = sum_{i,kh : py + i - 1 - kh = 2 dy} sum_{j,kw : px + j - 1 - kw = 2 dx} f[i] f[j] W[kh,kw]
// (t[2y+kh] += W[kh] x[y]; out[Y] = sum_i f[i] t[Y + i - 1]).  dst: [tap = (dy+1)*3 + (dx+1)][(py*2+px)*Cout + co][Cin] bf16.
__global__ void pack_composite_weight_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, int Cout,
                                             int Cin, float scale, float f0, float f1, float f2, float f3) {
  const float f[4] = {f0, f1, f2, f3};
  const int64_t total = (int64_t)9 * 4 * Cout * Cin;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % Cin);
    const int row = (int)((idx / Cin) % (4 * Cout));
    const int tap = (int)(idx / ((int64_t)Cin * 4 * Cout));
    const int co = row % Cout, ph = row / Cout;
    const int py = ph >> 1, px = ph & 1, dy = tap / 3 - 1, dx = tap % 3 - 1;
    const float* w = src + ((int64_t)co * Cin + ci) * 9;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kh = py + i - 1 - 2 * dy;
      if (kh < 0 || kh > 2) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kw = px + j - 1 - 2 * dx;
        if (kw < 0 || kw > 2) continue;
        acc = fmaf(f[i] * f[j], w[kh * 3 + kw], acc);
      }
    }
    dst[idx] = __float2bfloat16_rn(acc * scale);
  }
}

int launch_pack_composite_weight(__nv_bfloat16* dst, const float* src, int Cout, int Cin, float scale, const float* fir,
                                 cudaStream_t st) {
  const int64_t total = (int64_t)36 * Cout * Cin;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 8);
  pack_composite_weight_kernel<<<blocks, 256, 0, st>>>(dst, src, Cout, Cin, scale, fir[0], fir[1], fir[2], fir[3]);
  return check_launch("pack_composite_weight");
}

__global__ void scale_copy_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i] * scale;
}

int launch_scale_copy(float* dst, const float* src, int64_t n, float scale, cudaStream_t st) {
  if (n == 0) return L2I_OK;
  const int blocks = (int)std::min<int64_t>(ceil_div64(n, 256), (int64_t)kNumSMs * 8);
  scale_copy_kernel<<<blocks, 256, 0, st>>>(dst, src, n, scale);
  return check_launch("scale_copy");
}

// ------------------------------------------------------------------------------------------------
// image epilogues
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
image_to_uint8_kernel(uint8_t* __restrict__ out, const float* __restrict__ img, int B, int H, int W) {
  const int64_t plane = (int64_t)H * W;
  const int64_t total = (int64_t)B * plane;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / plane, p = idx - b * plane;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // np.uint8(np.clip(((x + 1) / 2.0) * 255, 0, 255)) in fp32, truncating
      float v = ((img[(b * 3 + c) * plane + p] + 1.0f) / 2.0f) * 255.0f;
      v = fminf(fmaxf(v, 0.f), 255.f);
      out[idx * 3 + c] = (uint8_t)v;
    }
  }
}

// NHWC activation (T) -> NCHW fp32, for the debug taps
template <typename T>
__global__ void nhwc_to_nchw_kernel(float* __restrict__ out, const T* __restrict__ in, int B, int H, int W, int C,
                                    const float* __restrict__ inv_scale, int64_t inv_bs) {
  const int64_t total = (int64_t)B * H * W * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    int64_t r = idx / C;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    float v = to_f32<T>(in[idx]);
    if (inv_scale != nullptr) v /= inv_scale[(int64_t)b * inv_bs + c];
    out[(((int64_t)b * C + c) * H + y) * W + x] = v;
  }
}

template <typename T>
int launch_nhwc_to_nchw(float* out, const void* in, int B, int H, int W, int C, const float* inv_scale,
                        int64_t inv_bs, cudaStream_t st) {
  const int64_t total = (int64_t)B * H * W * C;
  if (total == 0) return L2I_OK;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 8);
  nhwc_to_nchw_kernel<T><<<blocks, 256, 0, st>>>(out, (const T*)in, B, H, W, C, inv_scale, inv_bs);
  return check_launch("nhwc_to_nchw");
}
template int launch_nhwc_to_nchw<float>(float*, const void*, int, int, int, int, const float*, int64_t, cudaStream_t);
template int launch_nhwc_to_nchw<__nv_bfloat16>(float*, const void*, int, int, int, int, const float*, int64_t,
                                                cudaStream_t);

}  // namespace l2i

using namespace l2i;

extern "C" int l2i_image_to_uint8(uint8_t* out, const float* image, int B, int H, int W, void* stream) {
  L2I_REQUIRE(B >= 0 && H >= 0 && W >= 0, "image_to_uint8: bad shape");
  const int64_t total = (int64_t)B * H * W;
  if (total == 0) return L2I_OK;
  L2I_REQUIRE(out && image, "image_to_uint8: null tensor");
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 16);
  image_to_uint8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, image, B, H, W);
  return check_launch("image_to_uint8");
}
