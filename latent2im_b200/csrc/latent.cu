// Latent-side kernels: fused linear (mapping network / walk MLPs), PixelNorm, and the walk steps.
// These are launch- and latency-bound (B <= a few hundred rows of 512 floats); each output element
// is one warp-level dot product, weights stream once from L2, inputs sit in shared memory.
#include "common.cuh"

namespace l2i {

constexpr int kLinRowsPerBlock = 8;   // one warp per output feature
constexpr int kLinBatchTile = 8;      // batch rows staged in shared memory per pass

// y[b, n] = act(wscale * <x[b,:], W[n,:]> + bscale * bias[n]) * gain
// Optional per-output-row input offset table (row_xoff[n], in elements) lets one launch serve
// rows that read different slices of x (used for the per-layer style modulations).
__global__ void __launch_bounds__(kLinRowsPerBlock * 32)
linear_kernel(float* __restrict__ y, int64_t y_stride, const float* __restrict__ x, int64_t x_stride,
              const int* __restrict__ row_xoff, const float* __restrict__ W, const float* __restrict__ bias,
              int B, int N, int K, float wscale, float bscale, int act, float alpha, float gain) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * kLinRowsPerBlock + warp;
  const bool active = n < N;
  const float* wrow = W + (int64_t)(active ? n : 0) * K;
  const int xoff = (active && row_xoff != nullptr) ? row_xoff[n] : 0;
  const float b = (active && bias != nullptr) ? bias[n] * bscale : 0.f;
  for (int b0 = blockIdx.y * kLinBatchTile; b0 < B; b0 += gridDim.y * kLinBatchTile) {
    float acc[kLinBatchTile];
#pragma unroll
    for (int j = 0; j < kLinBatchTile; ++j) acc[j] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float w = active ? wrow[k] : 0.f;
#pragma unroll
      for (int j = 0; j < kLinBatchTile; ++j) {
        const int bb = b0 + j;
        if (bb < B) acc[j] = fmaf(w, x[(int64_t)bb * x_stride + xoff + k], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < kLinBatchTile; ++j) {
      float v = acc[j];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0 && active && b0 + j < B) {
        v = v * wscale + b;
        if (act == 1) v = lrelu(v, alpha);
        y[(int64_t)(b0 + j) * y_stride + n] = v * gain;
      }
    }
  }
}

int launch_linear(float* y, int64_t y_stride, const float* x, int64_t x_stride, const int* row_xoff,
                  const float* W, const float* bias, int B, int N, int K, float wscale, float bscale,
                  int act, float alpha, float gain, cudaStream_t st) {
  if (B == 0 || N == 0) return L2I_OK;
  dim3 grid(ceil_div(N, kLinRowsPerBlock), std::min(ceil_div(B, kLinBatchTile), 64));
  linear_kernel<<<grid, kLinRowsPerBlock * 32, 0, st>>>(y, y_stride, x, x_stride, row_xoff, W, bias, B, N, K,
                                                        wscale, bscale, act, alpha, gain);
  return check_launch("linear");
}

// Whole mapping network in one launch (Generator.style, networks.py:374-382): PixelNorm, then n_mlp x
// [EqualLinear(D, D, lr_mul) + fused leaky relu].  One CTA per latent row keeps the activation vector in shared
// memory between layers, so the 8 dependent layers cost one launch instead of 9; every warp produces D/16 outputs per
// layer, four weight rows (16 independent 16-byte loads per lane) in flight at a time.  Weights stream from L2.
constexpr int kMapThreads = 1024;
constexpr int kMapRows = 8;      // output rows per warp pass: 8 x 4 independent 16-byte weight loads per lane in flight
__global__ void __launch_bounds__(kMapThreads)
mapping_fused_kernel(float* __restrict__ w_out, const float* __restrict__ z, const float* const* __restrict__ Ws,
                     const float* const* __restrict__ bs, int n_mlp, int D, float wscale, float bscale) {
  extern __shared__ float xbuf[];   // [2][D]
  __shared__ float red[kMapThreads / 32];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kMapThreads / 32;
  // PixelNorm (networks.py:11-16)
  float ss = 0.f;
  for (int k = threadIdx.x; k < D; k += kMapThreads) { const float v = z[(int64_t)b * D + k]; xbuf[k] = v; ss = fmaf(v, v, ss); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < kWarps; ++i) tot += red[i];
  const float rn = rsqrtf(tot / (float)D + 1e-8f);
  for (int k = threadIdx.x; k < D; k += kMapThreads) xbuf[k] *= rn;
  __syncthreads();
  float* cur = xbuf;
  float* nxt = xbuf + D;
  const int D4 = D >> 2;
  for (int l = 0; l < n_mlp; ++l) {
    const float* W = Ws[l];
    const float* bias = bs[l];
    for (int n0 = warp * kMapRows; n0 < D; n0 += kWarps * kMapRows) {   // kMapRows output rows per pass (rows past D are skipped)
      float acc[kMapRows];
#pragma unroll
      for (int r = 0; r < kMapRows; ++r) acc[r] = 0.f;
      for (int k4 = lane; k4 < D4; k4 += 32) {
        const float4 xv = *reinterpret_cast<const float4*>(cur + 4 * k4);
#pragma unroll
        for (int r = 0; r < kMapRows; ++r) {
          const int n = min(n0 + r, D - 1);
          const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * D) + k4);
          acc[r] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[r]))));
        }
      }
#pragma unroll
      for (int r = 0; r < kMapRows; ++r) {
        float v = acc[r];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0 && n0 + r < D) {
          v = v * wscale + bias[n0 + r] * bscale;
          nxt[n0 + r] = lrelu(v, 0.2f) * 1.4142135623730951f;
        }
      }
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  for (int k = threadIdx.x; k < D; k += kMapThreads) w_out[(int64_t)b * D + k] = cur[k];
}

int launch_mapping_fused(float* w, const float* z, const float* const* Ws, const float* const* bs, int B, int n_mlp, int D,
                         float wscale, float bscale, cudaStream_t st) {
  if (B == 0) return L2I_OK;
  mapping_fused_kernel<<<B, kMapThreads, sizeof(float) * 2 * D, st>>>(w, z, Ws, bs, n_mlp, D, wscale, bscale);
  return check_launch("mapping_fused");
}

__global__ void __launch_bounds__(128) pixel_norm_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                         int B, int D) {
  const int b = blockIdx.x;
  const float* xr = x + (int64_t)b * D;
  float s = 0.f;
  for (int k = threadIdx.x; k < D; k += blockDim.x) s += xr[k] * xr[k];
  __shared__ float red[4];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  const float tot = red[0] + red[1] + red[2] + red[3];
  const float r = rsqrtf(tot / (float)D + 1e-8f);
  for (int k = threadIdx.x; k < D; k += blockDim.x) y[(int64_t)b * D + k] = xr[k] * r;
}

// out[b,i,:] = in[b,i,:] + sum_a alpha[b,a] * w[a,i,:]
__global__ void __launch_bounds__(128)
walk_linear_fwd_kernel(float* __restrict__ out, const float* __restrict__ in, int64_t in_bs, int64_t in_ls,
                       const float* __restrict__ alpha, const float* __restrict__ w, int B, int A,
                       int n_latent, int D, uint64_t mask) {
  const int i = blockIdx.x, b = blockIdx.y;
  const bool on = (mask >> i) & 1ull;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float v = in[(int64_t)b * in_bs + (int64_t)i * in_ls + k];
    if (on) {
      float d = 0.f;
      for (int a = 0; a < A; ++a) d = fmaf(alpha[b * A + a], w[((int64_t)a * n_latent + i) * D + k], d);
      v += d;
    }
    out[((int64_t)b * n_latent + i) * D + k] = v;
  }
}

// grad_w[a,i,:] = sum_b alpha[b,a] * grad_out[b,i,:]
__global__ void __launch_bounds__(128)
walk_linear_bwd_kernel(float* __restrict__ gw, const float* __restrict__ gout, const float* __restrict__ alpha,
                       int B, int A, int n_latent, int D, uint64_t mask) {
  const int i = blockIdx.x, a = blockIdx.y;
  const bool on = (mask >> i) & 1ull;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float s = 0.f;
    if (on)
      for (int b = 0; b < B; ++b) s = fmaf(alpha[b * A + a], gout[((int64_t)b * n_latent + i) * D + k], s);
    gw[((int64_t)a * n_latent + i) * D + k] = s;
  }
}

__global__ void __launch_bounds__(128)
walk_combine_kernel(float* __restrict__ out, const float* __restrict__ in, int64_t in_bs, int64_t in_ls,
                    const float* __restrict__ d, int64_t d_bs, int64_t d_ls, const float* __restrict__ coef,
                    int n_latent, int D, uint64_t mask, int normalize) {
  const int i = blockIdx.x, b = blockIdx.y;
  const bool on = (mask >> i) & 1ull;
  const float* dr = d + (int64_t)b * d_bs + (int64_t)i * d_ls;
  float c = coef != nullptr ? coef[b] : 1.f;
  if (on && normalize) {
    float s = 0.f;
    for (int k = threadIdx.x; k < D; k += blockDim.x) s += dr[k] * dr[k];
    __shared__ float red[4];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    c *= 1.f / sqrtf(red[0] + red[1] + red[2] + red[3]);
  }
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float v = in[(int64_t)b * in_bs + (int64_t)i * in_ls + k];
    if (on) v += c * dr[k];
    out[((int64_t)b * n_latent + i) * D + k] = v;
  }
}

}  // namespace l2i

using namespace l2i;

extern "C" int l2i_linear_fwd(float* y, int64_t y_stride, const float* x, int64_t x_stride, const float* W,
                              const float* bias, int B, int N, int K, float wscale, float bscale, int act,
                              float alpha, float gain, void* stream) {
  L2I_REQUIRE(B >= 0 && N >= 0 && K >= 1, "linear_fwd: bad shape");
  L2I_REQUIRE(act == 0 || act == 1, "linear_fwd: act must be 0 or 1");
  if (B == 0 || N == 0) return L2I_OK;
  L2I_REQUIRE(y && x && W, "linear_fwd: null tensor");
  return launch_linear(y, y_stride, x, x_stride, nullptr, W, bias, B, N, K, wscale, bscale, act, alpha, gain,
                       (cudaStream_t)stream);
}

extern "C" int l2i_pixel_norm(float* y, const float* x, int B, int D, void* stream) {
  L2I_REQUIRE(B >= 0 && D >= 1, "pixel_norm: bad shape");
  if (B == 0) return L2I_OK;
  L2I_REQUIRE(y && x, "pixel_norm: null tensor");
  pixel_norm_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(y, x, B, D);
  return check_launch("pixel_norm");
}

extern "C" int l2i_walk_linear_fwd(float* out, const float* in, int64_t in_batch_stride,
                                   int64_t in_layer_stride, const float* alpha, const float* w, int B, int A,
                                   int n_latent, int D, uint64_t layer_mask, void* stream) {
  L2I_REQUIRE(B >= 0 && A >= 0 && n_latent >= 1 && n_latent <= 64 && D >= 1, "walk_linear_fwd: bad shape");
  if (B == 0) return L2I_OK;
  L2I_REQUIRE(out && in && (A == 0 || (alpha && w)), "walk_linear_fwd: null tensor");
  walk_linear_fwd_kernel<<<dim3(n_latent, B), 128, 0, (cudaStream_t)stream>>>(
      out, in, in_batch_stride, in_layer_stride, alpha, w, B, A, n_latent, D, layer_mask);
  return check_launch("walk_linear_fwd");
}

extern "C" int l2i_walk_linear_bwd(float* grad_w, const float* grad_out, const float* alpha, int B, int A,
                                   int n_latent, int D, uint64_t layer_mask, void* stream) {
  L2I_REQUIRE(B >= 0 && A >= 1 && n_latent >= 1 && n_latent <= 64 && D >= 1, "walk_linear_bwd: bad shape");
  L2I_REQUIRE(grad_w && (B == 0 || (grad_out && alpha)), "walk_linear_bwd: null tensor");
  walk_linear_bwd_kernel<<<dim3(n_latent, A), 128, 0, (cudaStream_t)stream>>>(grad_w, grad_out, alpha, B, A,
                                                                            n_latent, D, layer_mask);
  return check_launch("walk_linear_bwd");
}

extern "C" int l2i_walk_combine(float* out, const float* in, int64_t in_batch_stride, int64_t in_layer_stride,
                                const float* d, int64_t d_batch_stride, int64_t d_layer_stride,
                                const float* coef, int B, int n_latent, int D, uint64_t layer_mask,
                                int normalize, void* stream) {
  L2I_REQUIRE(B >= 0 && n_latent >= 1 && n_latent <= 64 && D >= 1, "walk_combine: bad shape");
  if (B == 0) return L2I_OK;
  L2I_REQUIRE(out && in && d, "walk_combine: null tensor");
  walk_combine_kernel<<<dim3(n_latent, B), 128, 0, (cudaStream_t)stream>>>(
      out, in, in_batch_stride, in_layer_stride, d, d_batch_stride, d_layer_stride, coef, n_latent, D,
      layer_mask, normalize);
  return check_launch("walk_combine");
}
