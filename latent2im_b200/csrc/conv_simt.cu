// CUDA-core implicit-GEMM modulated convolution (fp32 accumulate, fp32 FFMA).
// This is the fp32 arbiter path on the GPU (max-abs <= 1e-3 gate) and the bring-up reference for
// the tcgen05 kernel; activations are NHWC in T (float or bf16), weights fp32 [tap][Cin][Cout].
//
// CTA tile: 64 output pixels (linear index over B*OH*OW) x 64 output channels, 256 threads,
// each thread a 4x4 register block; K loop = taps x Cin in steps of 16 staged through shared memory.
#include "conv_common.cuh"

namespace l2i {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const T* __restrict__ in, const float* __restrict__ wt, ConvGeom g, EpiParams e) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN];

  const int phase = blockIdx.z;
  const TapList& taps = g.taps[phase];
  const int py = phase >> 1, px = phase & 1;
  const int64_t M = (int64_t)g.B * g.OH * g.OW;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;

  // A-load role: pixel (tid / 4), ci sub-chunk (tid % 4) * 4
  const int a_m = tid >> 2, a_c = (tid & 3) * 4;
  int a_b = 0, a_oy = 0, a_ox = 0;
  const bool a_valid = (m0 + a_m) < M;
  if (a_valid) {
    int64_t m = m0 + a_m;
    a_ox = (int)(m % g.OW); m /= g.OW;
    a_oy = (int)(m % g.OH);
    a_b = (int)(m / g.OH);
  }
  // B-load role: k row (tid / 16), co (tid % 16) * 4
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;

  const int tm = tid >> 4, tn = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < taps.n; ++t) {
    const int iy = a_oy * g.in_scale + taps.dy[t], ix = a_ox * g.in_scale + taps.dx[t];
    const bool inb = a_valid && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
    const T* arow = in + (((int64_t)a_b * g.H + (inb ? iy : 0)) * g.W + (inb ? ix : 0)) * g.Cin;
    const float* wtap = wt + (int64_t)taps.wtap[t] * g.Cin * g.Cout;
    for (int c0 = 0; c0 < g.Cin; c0 += BK) {
      float av[4] = {0.f, 0.f, 0.f, 0.f};
      if (inb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) av[j] = to_f32<T>(arow[c0 + a_c + j]);
      }
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const float* wrow = wtap + (int64_t)(c0 + b_k) * g.Cout + n0 + b_n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n0 + b_n + j < g.Cout) bv[j] = wrow[j];
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) As[a_c + j][a_m] = av[j];
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[b_k][b_n + j] = bv[j];
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][tm * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tn * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

  // ---- fused epilogue ----
  const float nw = (e.noise != nullptr && e.noise_w != nullptr) ? *e.noise_w : 0.f;
  const int co0 = n0 + tn * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + tm * 4 + i;
    const bool valid = m < M;  // warp-uniform per half-warp? no: keep shuffles unconditional below
    int64_t r = valid ? m : 0;
    const int ox = (int)(r % g.OW); r /= g.OW;
    const int oy = (int)(r % g.OH);
    const int b = (int)(r / g.OH);
    const int Y = oy * g.out_scale + py, X = ox * g.out_scale + px;
    float rgb[3] = {0.f, 0.f, 0.f};
    float outv[4], yv[4];
    float nz = 0.f;
    if (e.mode == 0 && e.noise != nullptr && valid)
      nz = nw * e.noise[(int64_t)b * e.noise_bs + (int64_t)Y * g.out_W + X];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + j;
      float v = 0.f;
      if (co < g.Cout && valid) {
        v = acc[i][j] * (e.demod != nullptr ? e.demod[(int64_t)b * e.demod_bs + co] : 1.f);
        if (e.mode == 0) {
          v = lrelu(v + nz + e.bias[co], 0.2f) * 1.4142135623730951f;
          if (e.wr != nullptr) {
            const float* wr = e.wr + (int64_t)b * e.wr_bs;
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb[c] = fmaf(wr[c * g.Cout + co], v, rgb[c]);
          }
          yv[j] = v;
          if (e.s_next != nullptr) v *= e.s_next[(int64_t)b * e.s_next_bs + co];
        }
      }
      outv[j] = v;
    }
    if (valid && e.out != nullptr && (e.mode == 1 || e.s_next != nullptr)) {
      const int64_t off = (((int64_t)b * g.out_H + Y) * g.out_W + X) * g.Cout + co0;
      if (sizeof(T) == 2 && e.mode == 1 && e.raw_fp16) {
        __half* op = (__half*)e.out + off;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (co0 + j < g.Cout) op[j] = __float2half_rn(outv[j]);
      } else {
        T* op = (T*)e.out + off;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (co0 + j < g.Cout) op[j] = from_f32<T>(outv[j]);
      }
    }
    if (valid && e.mode == 0 && e.y_out != nullptr) {
      T* yp = (T*)e.y_out + (((int64_t)b * g.out_H + Y) * g.out_W + X) * g.Cout + co0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (co0 + j < g.Cout) yp[j] = from_f32<T>(yv[j]);
    }
    if (e.mode == 0 && e.wr != nullptr) {
      // reduce the three partial sums over the 16 lanes that share this pixel
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) rgb[c] += __shfl_xor_sync(0xffffffffu, rgb[c], off);
      }
      if (tn == 0 && valid) {
        const int64_t plane = (int64_t)g.out_H * g.out_W;
        if (e.fused_skip) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float v = rgb[c] + e.rgb_bias[c];
            if (e.skip_in != nullptr)
              v += upsample2x_at(e.skip_in + ((int64_t)b * 3 + c) * (plane / 4), g.out_H / 2, g.out_W / 2, Y, X, e.fir);
            e.skip_out[((int64_t)b * 3 + c) * plane + (int64_t)Y * g.out_W + X] = v;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            e.rgb_part[(((int64_t)blockIdx.y * g.B + b) * 3 + c) * plane + (int64_t)Y * g.out_W + X] = rgb[c];
        }
      }
    }
  }
}

template <typename T>
int launch_conv_simt(const void* in, const float* wt, const ConvGeom& g, const EpiParams& e, cudaStream_t st) {
  const int64_t M = (int64_t)g.B * g.OH * g.OW;
  if (M == 0) return L2I_OK;
  dim3 grid((unsigned)ceil_div64(M, BM), (unsigned)ceil_div(g.Cout, BN), (unsigned)g.nphase);
  if (g.Cin % BK != 0) {
    set_error("conv_simt: Cin=%d must be a multiple of %d", g.Cin, BK);
    return L2I_ERR_UNSUPPORTED;
  }
  if (e.fused_skip && grid.y != 1) {
    set_error("conv_simt: fused skip needs a single N tile (Cout=%d)", g.Cout);
    return L2I_ERR_INVALID_ARG;
  }
  conv_simt_kernel<T><<<grid, 256, 0, st>>>((const T*)in, wt, g, e);
  return check_launch("conv_simt");
}

template int launch_conv_simt<float>(const void*, const float*, const ConvGeom&, const EpiParams&, cudaStream_t);
template int launch_conv_simt<__nv_bfloat16>(const void*, const float*, const ConvGeom&, const EpiParams&, cudaStream_t);

}  // namespace l2i
