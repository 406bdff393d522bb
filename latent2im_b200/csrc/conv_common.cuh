// Geometry + fused-epilogue description shared by the CUDA-core (fp32) and tcgen05 (bf16)
// implicit-GEMM modulated convolutions.
//
// Maths (reference networks.py:231-286, 330-358 restated with the input-scaling identity, SURVEY
// section 0.6a):  the producer of an activation stores  x~ = x * s_next[b, ci]  so every conv is a
// shared-weight contraction  acc[p, co] = sum_{tap, ci} Wt[tap][ci][co] * x~[p + tap, ci],
// followed by this epilogue:
//     v   = acc * demod[b, co]                                   (demodulation as an output scale)
//     raw mode (up-conv, before the blur):  store v
//     act mode:  y = lrelu(v + noise_w * noise[b, p] + bias[co], 0.2) * sqrt(2)
//                store  y * s_next[b, co]            (omitted for the last layer)
//                rgb[c] += wr[b, c, co] * y          (ToRGB 1x1 modulated conv, no demod)
// and per pixel either  rgb partial sums -> rgb_part  or, when one CTA owns all Cout channels,
// skip_out = rgb + rgb_bias + upsample2x(skip_in)   (ToRGB.forward incl. Upsample, :349-358).
#pragma once
#include "common.cuh"

namespace l2i {

constexpr int kMaxTaps = 9;

struct TapList {
  int n;
  int8_t dy[kMaxTaps];
  int8_t dx[kMaxTaps];
  int8_t wtap[kMaxTaps];  // index into the packed weight's tap dimension (kh*3+kw)
};

struct ConvGeom {
  int B, H, W, Cin, Cout;  // input activation [B, H, W, Cin] (NHWC)
  int OH, OW;              // output grid per phase
  int nphase;              // 1 (plain) or 4 (stride-2 transposed conv, phase-decomposed)
  TapList taps[4];
  int out_scale;           // 1 or 2: output pixel = (oy*out_scale + py, ox*out_scale + px)
  int in_scale;            // 1, or 2 for the data gradient of the stride-2 transposed conv: input pixel = oy*in_scale + dy
  int out_H, out_W;        // allocated dims of the output tensor
};

struct EpiParams {
  int mode;                 // 0 = act, 1 = raw
  int raw_fp16;             // raw mode with 16-bit storage: write fp16 (the demodulated conv output is bounded)
  const float* demod;       // [B][demod_bs]; nullptr = 1 (data-gradient convs)
  int64_t demod_bs;
  const float* bias;        // [Cout]
  const float* noise;       // [Bn][H][W] fp32 or nullptr
  int64_t noise_bs;         // 0 for broadcast
  const float* noise_w;     // device scalar
  const float* s_next;      // [B][s_next_bs] or nullptr (no activation output)
  int64_t s_next_bs;
  void* out;                // NHWC [B][out_H][out_W][Cout]
  void* y_out;              // training: unscaled activation y (same layout), saved for the backward pass
  // ToRGB
  const float* wr;          // [B][wr_bs] laid out [3][Cout], or nullptr
  int64_t wr_bs;
  float* rgb_part;          // [nparts][B][3][H][W]
  int fused_skip;           // 1: write skip_out directly (requires a single N tile)
  const float* rgb_bias;    // [3]
  const float* skip_in;     // [B][3][H/2][W/2] or nullptr
  float* skip_out;          // [B][3][H][W]
  float fir[4];             // separable up-sampling taps (already * factor), flipped order
};

// 2x FIR upsample of the low-res skip at output pixel (Y, X): reference Upsample (networks.py:30-48)
// = upfirdn2d(up=2, pad=(2,1)) with kernel outer(taps)/sum*4.  f[] holds the *flipped* 1-D taps times 2.
__device__ __forceinline__ float upsample2x_at(const float* __restrict__ plane, int h, int w, int Y, int X,
                                               const float* f) {
  // out[Y] = sum_i f[i] * U[Y + i - 2], U[2y] = in[y], zero elsewhere / outside
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int u = Y + i - 2;
    if (u < 0 || (u & 1)) continue;
    const int y = u >> 1;
    if (y >= h) continue;
    float row = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int v = X + j - 2;
      if (v < 0 || (v & 1)) continue;
      const int x = v >> 1;
      if (x >= w) continue;
      row = fmaf(f[j], plane[(int64_t)y * w + x], row);
    }
    acc = fmaf(f[i], row, acc);
  }
  return acc;
}

}  // namespace l2i
