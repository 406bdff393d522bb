// Latent-side kernels: fused linear (mapping network / walk MLPs), PixelNorm, and the walk steps.
// These are launch- and latency-bound (B <= a few hundred rows of 512 floats); each output element
// is one warp-level dot product, weights stream once from L2, inputs sit in shared memory.
#include <cstdlib>

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

// Batch-in-lanes variant (the 26 modulation linears of a forward: ~7000 rows x 512 at B = 32; the walk MLPs):
// lane = batch row, a warp owns 4 consecutive output rows, the input tile sits TRANSPOSED in shared memory ([K][32], so
// lane b reads xs[k][b] without bank conflicts) and the weight row is a warp-uniform 16-byte load.  Per 4 k: 4 uniform LDG.128
// + 4 LDS + 16 FMA, no cross-lane reduction, one 16-byte store per lane.  Weights are read once per 32 batch rows.
// Requires the rows of one 4-row group to share row_xoff (true for the generator: every Cin is a multiple of 4).
constexpr int kLinBtWarps = 8;
constexpr int kXsPitch = 33;
__global__ void __launch_bounds__(kLinBtWarps * 32)
linear_bt_kernel(float* __restrict__ y, int64_t y_stride, const float* __restrict__ x, int64_t x_stride,
                 const int* __restrict__ row_xoff, const float* __restrict__ W, const float* __restrict__ bias,
                 int B, int N, int K, float wscale, float bscale, int act, float alpha, float gain, int groups_per_block) {
  extern __shared__ float xs[];   // [K][33] input tile (transposed, padded), then [kLinBtWarps][4][K] weight slabs (16-byte aligned)
  float* ws_all = xs + ((K * kXsPitch + 3) & ~3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.y * 32;
  const int g_begin = blockIdx.x * groups_per_block, g_end = min(g_begin + groups_per_block, (N + 3) / 4);
  int staged_xoff = -1;
  for (int gb = g_begin; gb < g_end; gb += kLinBtWarps) {
    // all warps of the block walk the groups together so that a change of the input slice restages the tile once
    const int g = gb + warp;
    const int xoff_blk = row_xoff != nullptr ? row_xoff[min(gb * 4, N - 1)] : 0;
    const int xoff_last = row_xoff != nullptr ? row_xoff[min((min(gb + kLinBtWarps, g_end)) * 4 - 1, N - 1)] : 0;
    if (xoff_blk != staged_xoff) {
      __syncthreads();
      // x tile: one burst of 16-byte cp.async copies (row-major, rows >= B zero-filled) into the weight-slab region, then a
      // shared -> shared transpose into the padded [K][33] layout (a register-staged loop costs one L2 round trip per element:
      // 28 us per tile, measured)
      {
        const int kq = K >> 2;                                       // 16-byte chunks per row
        for (int c = threadIdx.x; c < 32 * kq; c += blockDim.x) {
          const int bb = c / kq, q4 = c - bb * kq;
          const bool ok = b0 + bb < B;
          const float* src = x + (int64_t)(ok ? b0 + bb : 0) * x_stride + xoff_blk + 4 * q4;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(ws_all + bb * K + 4 * q4)),
                       "l"(src), "r"(ok ? 16 : 0)
                       : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        for (int i = threadIdx.x; i < K * 32; i += blockDim.x) {
          const int bb = i / K, k = i - bb * K;
          xs[k * kXsPitch + bb] = ws_all[i];
        }
      }
      __syncthreads();
      staged_xoff = xoff_blk;
    }
    if (g < g_end) {
      const int n0 = g * 4;
      const bool uniform = xoff_last == xoff_blk;       // the whole pass reads the staged slice
      const int my_xoff = row_xoff != nullptr ? row_xoff[n0] : 0;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (uniform || my_xoff == staged_xoff) {
        const float* w0 = W + (int64_t)min(n0, N - 1) * K;
        const float* w1 = W + (int64_t)min(n0 + 1, N - 1) * K;
        const float* w2 = W + (int64_t)min(n0 + 2, N - 1) * K;
        const float* w3 = W + (int64_t)min(n0 + 3, N - 1) * K;
        // The four weight rows (4 x K floats) are fetched with ONE burst of cp.async copies into this warp's shared-memory
        // slab and read back as broadcast LDS.128: a register-fed loop keeps only ~4 loads in flight per lane (ptxas interleaves
        // loads and FMAs) and is L2-latency bound at ~60 us per 512 x 512 layer (measured).
        float* wsl = ws_all + warp * 4 * K;
        for (int c = lane; c < K; c += 32) {          // c: 16-byte chunk index over the 4 rows (K / 4 chunks per row)
          const int r = c / (K >> 2), kc = c - r * (K >> 2);
          const float* src = (r == 0 ? w0 : (r == 1 ? w1 : (r == 2 ? w2 : w3))) + 4 * kc;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(wsl + r * K + 4 * kc)), "l"(src) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
#pragma unroll 4
        for (int k = 0; k < K; k += 4) {
          const float4 a0 = *reinterpret_cast<const float4*>(wsl + k);
          const float4 a1 = *reinterpret_cast<const float4*>(wsl + K + k);
          const float4 a2 = *reinterpret_cast<const float4*>(wsl + 2 * K + k);
          const float4 a3 = *reinterpret_cast<const float4*>(wsl + 3 * K + k);
          const float x0 = xs[(k + 0) * kXsPitch + lane], x1 = xs[(k + 1) * kXsPitch + lane], x2 = xs[(k + 2) * kXsPitch + lane], x3 = xs[(k + 3) * kXsPitch + lane];
          acc[0] = fmaf(a0.x, x0, fmaf(a0.y, x1, fmaf(a0.z, x2, fmaf(a0.w, x3, acc[0]))));
          acc[1] = fmaf(a1.x, x0, fmaf(a1.y, x1, fmaf(a1.z, x2, fmaf(a1.w, x3, acc[1]))));
          acc[2] = fmaf(a2.x, x0, fmaf(a2.y, x1, fmaf(a2.z, x2, fmaf(a2.w, x3, acc[2]))));
          acc[3] = fmaf(a3.x, x0, fmaf(a3.y, x1, fmaf(a3.z, x2, fmaf(a3.w, x3, acc[3]))));
        }
        __syncwarp();                                  // the slab is rewritten by this warp's next group
      } else {
        // a pass that straddles two input slices (never the case for the generator's tables): read x from global
        for (int r = 0; r < 4; ++r) {
          const int n = min(n0 + r, N - 1);
          const int xo = row_xoff[n];
          if (b0 + lane < B)
            for (int k = 0; k < K; ++k) acc[r] = fmaf(W[(int64_t)n * K + k], x[(int64_t)(b0 + lane) * x_stride + xo + k], acc[r]);
        }
      }
      if (b0 + lane < B) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (n0 + r >= N) break;
          float v = acc[r] * wscale + (bias != nullptr ? bias[n0 + r] * bscale : 0.f);
          if (act == 1) v = lrelu(v, alpha);
          y[(int64_t)(b0 + lane) * y_stride + n0 + r] = v * gain;
        }
      }
    }
  }
}

int launch_linear(float* y, int64_t y_stride, const float* x, int64_t x_stride, const int* row_xoff,
                  const float* W, const float* bias, int B, int N, int K, float wscale, float bscale,
                  int act, float alpha, float gain, cudaStream_t st) {
  if (B == 0 || N == 0) return L2I_OK;
  const size_t bt_smem = ((((size_t)K * kXsPitch + 3) & ~(size_t)3) + (size_t)K * 4 * kLinBtWarps) * sizeof(float);
  // taken for EVERY batch size (not only the large ones): the summation order must not depend on the batch, image i of a
  // batch is bit-identical to the same latent run alone (tests/test_gpu_generator.py::test_large_batch_matches_single_sample_runs)
  static const bool bt_off = std::getenv("L2I_LINEAR_BT") != nullptr && std::atoi(std::getenv("L2I_LINEAR_BT")) == 0;   // debug A/B
  if (!bt_off && K % 4 == 0 && bt_smem <= 200 * 1024 && ((uintptr_t)W % 16 == 0) && ((uintptr_t)x % 16 == 0) && x_stride % 4 == 0) {
    static bool attr_set = false;
    if (!attr_set) {
      L2I_CUDA_TRY(cudaFuncSetAttribute(linear_bt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    const int groups = (N + 3) / 4;
    // enough blocks to fill the machine, each a multiple of kLinBtWarps groups so a block restages x as rarely as possible
    const int target_blocks = std::max(1, 2 * kNumSMs / ceil_div(B, 32));
    const int gpb = ceil_div(ceil_div(groups, target_blocks), kLinBtWarps) * kLinBtWarps;
    dim3 grid(ceil_div(groups, gpb), ceil_div(B, 32));
    linear_bt_kernel<<<grid, kLinBtWarps * 32, bt_smem, st>>>(y, y_stride, x, x_stride, row_xoff, W, bias, B, N, K, wscale, bscale,
                                                               act, alpha, gain, gpb);
    return check_launch("linear_bt");
  }
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


// ---- backward of the fused linear (walk MLP training; transform_base.py:175-179, 214-217 under autograd) ----------------
// g = gy * act'(y) is formed on the fly from the saved OUTPUT's sign (leaky relu: y > 0 <=> pre-activation > 0).
__device__ __forceinline__ float lin_g(const float* __restrict__ gy, const float* __restrict__ y, int64_t i, int act, float alpha) {
  const float g = gy[i];
  return (act == 1 && y[i] <= 0.f) ? g * alpha : g;
}

// gx[b, k] = sum_n g[b, n] W[n, k]: one thread per k (coalesced weight rows), kLinBatchTile batch rows per block
__global__ void __launch_bounds__(128)
linear_bwd_x_kernel(float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ y,
                    const float* __restrict__ W, int B, int N, int K, int act, float alpha) {
  extern __shared__ float gs[];   // [kLinBatchTile][N]
  const int b0 = blockIdx.y * kLinBatchTile;
  for (int i = threadIdx.x; i < kLinBatchTile * N; i += blockDim.x) {
    const int j = i / N, n = i - j * N;
    gs[i] = (b0 + j < B) ? lin_g(gy, y, (int64_t)(b0 + j) * N + n, act, alpha) : 0.f;
  }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float acc[kLinBatchTile];
#pragma unroll
  for (int j = 0; j < kLinBatchTile; ++j) acc[j] = 0.f;
  for (int n = 0; n < N; ++n) {
    const float w = W[(int64_t)n * K + k];
#pragma unroll
    for (int j = 0; j < kLinBatchTile; ++j) acc[j] = fmaf(gs[j * N + n], w, acc[j]);
  }
#pragma unroll
  for (int j = 0; j < kLinBatchTile; ++j)
    if (b0 + j < B) gx[(int64_t)(b0 + j) * K + k] = acc[j];
}

// gW[n, k] = sum_b g[b, n] x[b, k];  gb[n] = sum_b g[b, n]: one block per n, threads over k
__global__ void __launch_bounds__(256)
linear_bwd_w_kernel(float* __restrict__ gW, float* __restrict__ gb, const float* __restrict__ gy, const float* __restrict__ y,
                    const float* __restrict__ x, int B, int N, int K, int act, float alpha) {
  const int n = blockIdx.x;
  float bsum = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(lin_g(gy, y, (int64_t)b * N + n, act, alpha), x[(int64_t)b * K + k], acc);
    gW[(int64_t)n * K + k] = acc;
  }
  if (gb != nullptr && threadIdx.x == 0) {
    for (int b = 0; b < B; ++b) bsum += lin_g(gy, y, (int64_t)b * N + n, act, alpha);
    gb[n] = bsum;
  }
}

// Gradient of walk_combine w.r.t. d (one block per (layer, sample)); with d_ls == 0 (one MLP output serves every layer)
// the per-layer contributions are summed by the caller's second kernel below.
//   out = in + c * d            : gd = c * g
//   out = in + d / ||d||        : gd = (g - <g, u> u) / ||d||,  u = d / ||d||
__global__ void __launch_bounds__(128)
walk_combine_bwd_kernel(float* __restrict__ gd, const float* __restrict__ g, const float* __restrict__ d, int64_t d_bs,
                        int64_t d_ls, const float* __restrict__ coef, int n_latent, int D, uint64_t mask, int normalize) {
  const int i = blockIdx.x, b = blockIdx.y;
  const bool on = (mask >> i) & 1ull;
  const float* dr = d + (int64_t)b * d_bs + (int64_t)i * d_ls;
  const float* gr = g + ((int64_t)b * n_latent + i) * D;
  float* o = gd + ((int64_t)b * n_latent + i) * D;
  const float c = coef != nullptr ? coef[b] : 1.f;
  __shared__ float red[2][4];
  float inv = 1.f, dot = 0.f;
  if (on && normalize) {
    float s = 0.f, t = 0.f;
    for (int k = threadIdx.x; k < D; k += blockDim.x) { s += dr[k] * dr[k]; t += gr[k] * dr[k]; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); t += __shfl_xor_sync(0xffffffffu, t, off); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = t; }
    __syncthreads();
    const float nn = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    inv = 1.f / sqrtf(nn);
    dot = (red[1][0] + red[1][1] + red[1][2] + red[1][3]) / nn;      // <g, d> / ||d||^2
  }
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float v = 0.f;
    if (on) v = normalize ? c * inv * (gr[k] - dot * dr[k]) : c * gr[k];
    o[k] = v;
  }
}

// gd_shared[b, :] = sum_i gd[b, i, :]
__global__ void __launch_bounds__(128)
layer_sum_kernel(float* __restrict__ out, const float* __restrict__ in, int n_latent, int D) {
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < n_latent; ++i) s += in[((int64_t)b * n_latent + i) * D + k];
    out[(int64_t)b * D + k] = s;
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

extern "C" int l2i_linear_bwd(float* gx, float* gW, float* gb, const float* gy, const float* y, const float* x, const float* W,
                              int B, int N, int K, int act, float alpha, void* stream) {
  L2I_REQUIRE(B >= 0 && N >= 1 && K >= 1, "linear_bwd: bad shape");
  L2I_REQUIRE(act == 0 || act == 1, "linear_bwd: act must be 0 or 1");
  if (B == 0) return L2I_OK;
  L2I_REQUIRE(gy && x && W && (act == 0 || y), "linear_bwd: null tensor");
  L2I_REQUIRE((size_t)kLinBatchTile * N * sizeof(float) <= 48 * 1024, "linear_bwd: N = %d too large for the staged gradient tile", N);
  cudaStream_t st = (cudaStream_t)stream;
  if (gx != nullptr) {
    linear_bwd_x_kernel<<<dim3(ceil_div(K, 128), ceil_div(B, kLinBatchTile)), 128, kLinBatchTile * N * sizeof(float), st>>>(
        gx, gy, y, W, B, N, K, act, alpha);
    L2I_TRY(check_launch("linear_bwd_x"));
  }
  if (gW != nullptr) {
    linear_bwd_w_kernel<<<N, 256, 0, st>>>(gW, gb, gy, y, x, B, N, K, act, alpha);
    L2I_TRY(check_launch("linear_bwd_w"));
  }
  return L2I_OK;
}

extern "C" int l2i_walk_combine_bwd(float* grad_d, float* scratch, const float* grad_out, const float* d, int64_t d_batch_stride,
                                    int64_t d_layer_stride, const float* coef, int B, int n_latent, int D, uint64_t layer_mask,
                                    int normalize, void* stream) {
  L2I_REQUIRE(B >= 0 && n_latent >= 1 && n_latent <= 64 && D >= 1, "walk_combine_bwd: bad shape");
  if (B == 0) return L2I_OK;
  L2I_REQUIRE(grad_d && grad_out && d, "walk_combine_bwd: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  const bool shared = d_layer_stride == 0;
  L2I_REQUIRE(!shared || scratch != nullptr, "walk_combine_bwd: shared d needs a [B, n_latent, D] scratch buffer");
  walk_combine_bwd_kernel<<<dim3(n_latent, B), 128, 0, st>>>(shared ? scratch : grad_d, grad_out, d, d_batch_stride, d_layer_stride,
                                                            coef, n_latent, D, layer_mask, normalize);
  L2I_TRY(check_launch("walk_combine_bwd"));
  if (shared) {
    layer_sum_kernel<<<B, 128, 0, st>>>(grad_d, scratch, n_latent, D);
    L2I_TRY(check_launch("walk_combine_bwd_sum"));
  }
  return L2I_OK;
}
