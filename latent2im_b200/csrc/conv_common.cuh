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

// Kernel-selection switches (debug / test): read from the environment by every l2i_generator_create.
//   L2I_HALO=0, L2I_HALO_MASK=<bits: 1 plain Cin=64, 2 up-conv, 4 pair-packed Cin=32, 8 composite>, L2I_QUAD=0, L2I_ARES=0, L2I_VPAIR=0,
//   L2I_UPROW=0 (composite 6x6 up-conv kernels instead of the row-marching fused up-conv), L2I_UPROW_MASK=<bits: 1 64->32, 2 128->64, 4 256->128>,
//   L2I_CLUSTER=1 (experiment: CTA pairs with TMA-multicast weight rings in the weight-streaming kernels; measured no gain),
//   L2I_HRING_STORE=0 (direct 16-byte stores instead of staging + TMA tensor stores in the halo-ring kernel),
//   L2I_HRING=0 (general kernel instead of the halo-ring kernel for the plain layers with Cin >= 256),
//   L2I_ARES_PAIR=0 (one-tile-per-epilogue A-resident kernel for the plain 128 -> 128 layer),
//   L2I_FIR_SIMT=1 (register-window FIR kernels instead of the TMA-fed ones), L2I_HALO_BASE_OFFSET=1 (descriptor experiment)
struct KernelSwitches {
  int halo = 1, halo_mask = 15, halo_base_offset = 0, quad = 1, ares = 1, vpair = 1, fir_simt = 0, uprow = 1, uprow_mask = 7, cluster = 0, ares_pair = 1, hring = 1, hring_store = 1;
};
extern KernelSwitches g_switches;
void refresh_kernel_switches();

constexpr int kMaxTaps = 18;  // 9 spatial taps x (hi, lo) halves of split-bf16 weights

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
  int in_pair_packed;      // Cin == 32 input stored as [B][H/2][W][2][32]: vertical pixel pairs form 128-byte units
  int weight_taps;         // tap slices in the packed weight tensor: 9, or 18 when it holds bf16 hi + lo halves
  // Composite up-conv (stride-2 transposed 3x3 conv + 4x4 blur folded into one 6x6 stride-2 kernel):
  // run as a plain 3x3 conv at INPUT resolution whose GEMM N = 4 * up_cout enumerates (phase = py*2+px, co);
  // column n lands at output pixel (2*oy + py, 2*ox + px), channel co.  0 = not composite.
  int up_cout;
  int out_pair_packed;     // composite only: write the Cout == 32 output as [B][H][2W][2][32] vertical pixel pairs
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
  float* skip_out;          // [B][3][H][W]; may be nullptr when only image_u8 is wanted (2x2-block kernel)
  uint8_t* image_u8;        // optional fused output of the LAST ToRGB: [B][H][W][3] = clip((x+1)/2*255) truncated (transform_base.py:625-626)
  float fir[4];             // separable up-sampling taps (already * factor), flipped order
};

// 2x FIR upsample of the low-res skip at output pixel (Y, X): reference Upsample (networks.py:30-48)
// = upfirdn2d(up=2, pad=(2,1)) with kernel outer(taps)/sum*4.  f[] holds the *flipped* 1-D taps times 2.
//   out[Y] = sum_i f[i] * U[Y + i - 2],  U[2y] = in[y], zero elsewhere / outside.
// Exactly two input rows (and two columns) contribute to any output pixel:
//   Y = 2m   : rows m-1 (f[0]) and m   (f[2]);      Y = 2m+1 : rows m (f[1]) and m+1 (f[3])
// so the gather is 4 branch-free loads per channel (out-of-range taps get weight 0 and a clamped address).
struct Up2Taps {
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
};

__device__ __forceinline__ Up2Taps make_up2_taps(int h, int w, int Y, int X, const float* f) {
  Up2Taps t;
  const int py = Y & 1, px = X & 1;
  const int ya = (Y >> 1) - 1 + py, xa = (X >> 1) - 1 + px;
  t.wy0 = (ya >= 0 && ya < h) ? f[py] : 0.f;
  t.wy1 = (ya + 1 >= 0 && ya + 1 < h) ? f[py + 2] : 0.f;
  t.wx0 = (xa >= 0 && xa < w) ? f[px] : 0.f;
  t.wx1 = (xa + 1 >= 0 && xa + 1 < w) ? f[px + 2] : 0.f;
  t.y0 = min(max(ya, 0), h - 1); t.y1 = min(max(ya + 1, 0), h - 1);
  t.x0 = min(max(xa, 0), w - 1); t.x1 = min(max(xa + 1, 0), w - 1);
  return t;
}

__device__ __forceinline__ float up2_apply(const float* __restrict__ plane, int w, const Up2Taps& t) {
  const float a = __ldg(plane + (int64_t)t.y0 * w + t.x0), b = __ldg(plane + (int64_t)t.y0 * w + t.x1);
  const float c = __ldg(plane + (int64_t)t.y1 * w + t.x0), d = __ldg(plane + (int64_t)t.y1 * w + t.x1);
  return t.wy0 * (t.wx0 * a + t.wx1 * b) + t.wy1 * (t.wx0 * c + t.wx1 * d);
}

__device__ __forceinline__ float upsample2x_at(const float* __restrict__ plane, int h, int w, int Y, int X,
                                               const float* f) {
  const Up2Taps t = make_up2_taps(h, w, Y, X, f);
  return up2_apply(plane, w, t);
}

}  // namespace l2i
