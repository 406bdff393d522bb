// upfirdn2d: pad -> zero-insert upsample -> 2-D FIR -> decimate, on [major, in_h, in_w, minor].
// Replaces op/upfirdn2d_kernel.cu of the reference.  Two kernels:
//   * upfirdn2d_planar_kernel  (minor == 1, the layout every reference call site uses): a CTA
//     stages the input footprint of a 32x64 output tile in shared memory with coalesced loads,
//     the (flipped) FIR taps sit in shared memory, each thread produces 8 outputs.
//   * upfirdn2d_gather_kernel  generic fallback for minor > 1.
// Unlike the reference every (up, down, pad, kernel<=8x8) combination is computed; nothing
// returns uninitialised memory.  All index arithmetic that can exceed 2^31 is 64-bit.
#include "common.cuh"

namespace l2i {

struct UpfirdnParams {
  int64_t major;
  int in_h, in_w, minor, kh, kw;
  int up_x, up_y, down_x, down_y;
  int pad_x0, pad_y0;
  int out_h, out_w;
};

__device__ __forceinline__ int floor_div_i(int a, int b) {
  int q = a / b;
  return (q * b > a) ? q - 1 : q;
}

// ---- generic gather ----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_gather_kernel(T* __restrict__ y, const T* __restrict__ x,
                                                               const float* __restrict__ kernel,
                                                               UpfirdnParams p) {
  __shared__ float sk[64];  // flipped taps
  for (int t = threadIdx.x; t < p.kh * p.kw; t += blockDim.x) {
    int ky = t / p.kw, kx = t - ky * p.kw;
    sk[t] = kernel[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
  }
  __syncthreads();
  const int64_t total = p.major * p.out_h * (int64_t)p.out_w * p.minor;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int mi = (int)(r % p.minor); r /= p.minor;
    const int ox = (int)(r % p.out_w); r /= p.out_w;
    const int oy = (int)(r % p.out_h);
    const int64_t mj = r / p.out_h;
    const int mid_x = ox * p.down_x + p.up_x - 1 - p.pad_x0;
    const int mid_y = oy * p.down_y + p.up_y - 1 - p.pad_y0;
    const int in_x0 = floor_div_i(mid_x, p.up_x);
    const int in_y0 = floor_div_i(mid_y, p.up_y);
    const int kx0 = (in_x0 + 1) * p.up_x - mid_x - 1;
    const int ky0 = (in_y0 + 1) * p.up_y - mid_y - 1;
    float acc = 0.f;
    for (int ky = ky0, iy = in_y0; ky < p.kh; ky += p.up_y, ++iy) {
      if (iy < 0 || iy >= p.in_h) continue;
      for (int kx = kx0, ix = in_x0; kx < p.kw; kx += p.up_x, ++ix) {
        if (ix < 0 || ix >= p.in_w) continue;
        acc += to_f32<T>(x[((mj * p.in_h + iy) * p.in_w + ix) * p.minor + mi]) * sk[ky * p.kw + kx];
      }
    }
    y[i] = from_f32<T>(acc);
  }
}

// ---- planar tiled kernel (minor == 1) -----------------------------------------------------------
// Output tile TH x TW per CTA iteration; the input footprint is at most
//   ((TH-1)*down + kh - 1)/up + 2 rows (same for columns).
constexpr int kTileH = 32, kTileW = 64;

template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_planar_kernel(T* __restrict__ y, const T* __restrict__ x,
                                                               const float* __restrict__ kernel,
                                                               UpfirdnParams p, int tiles_x, int tiles_y,
                                                               int fh, int fw) {
  extern __shared__ float smem[];
  float* sk = smem;        // 64 flipped taps
  float* sx = smem + 64;   // fh x fw input footprint
  for (int t = threadIdx.x; t < p.kh * p.kw; t += blockDim.x) {
    int ky = t / p.kw, kx = t - ky * p.kw;
    sk[t] = kernel[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
  }
  const int64_t tiles_per_plane = (int64_t)tiles_x * tiles_y;
  const int64_t ntiles = tiles_per_plane * p.major;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t plane = tile / tiles_per_plane;
    const int tr = (int)(tile - plane * tiles_per_plane);
    const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
    const int oy0 = ty * kTileH, ox0 = tx * kTileW;
    const int mid_y0 = oy0 * p.down_y + p.up_y - 1 - p.pad_y0;
    const int mid_x0 = ox0 * p.down_x + p.up_x - 1 - p.pad_x0;
    const int in_y_base = floor_div_i(mid_y0, p.up_y);
    const int in_x_base = floor_div_i(mid_x0, p.up_x);
    const T* xp = x + plane * (int64_t)p.in_h * p.in_w;
    __syncthreads();
    for (int i = threadIdx.x; i < fh * fw; i += blockDim.x) {
      int ry = i / fw, rx = i - ry * fw;
      int iy = in_y_base + ry, ix = in_x_base + rx;
      float v = 0.f;
      if (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) v = to_f32<T>(xp[(int64_t)iy * p.in_w + ix]);
      sx[i] = v;
    }
    __syncthreads();
    T* yp = y + plane * (int64_t)p.out_h * p.out_w;
    for (int o = threadIdx.x; o < kTileH * kTileW; o += blockDim.x) {
      const int ry = o / kTileW, rx = o - ry * kTileW;
      const int oy = oy0 + ry, ox = ox0 + rx;
      if (oy >= p.out_h || ox >= p.out_w) continue;
      const int mid_y = mid_y0 + ry * p.down_y, mid_x = mid_x0 + rx * p.down_x;
      const int in_y = floor_div_i(mid_y, p.up_y), in_x = floor_div_i(mid_x, p.up_x);
      const int ky0 = (in_y + 1) * p.up_y - mid_y - 1, kx0 = (in_x + 1) * p.up_x - mid_x - 1;
      const int sy = in_y - in_y_base, sxo = in_x - in_x_base;
      float acc = 0.f;
      for (int ky = ky0, r = sy; ky < p.kh; ky += p.up_y, ++r)
        for (int kx = kx0, c = sxo; kx < p.kw; kx += p.up_x, ++c) acc += sx[r * fw + c] * sk[ky * p.kw + kx];
      yp[(int64_t)oy * p.out_w + ox] = from_f32<T>(acc);
    }
  }
}

// ---- 4 x 4 kernels with up, down in {1, 2} (the reference's modes 1, 3, 5: Blur, Upsample, Downsample / their transposes) -----
// op/upfirdn2d_kernel.cu:177-211.  A CTA owns a 32 x 128 output tile: the input footprint is staged once in shared memory
// (coalesced loads, zero fill outside the image = the op's padding), every thread produces 4 adjacent outputs of 4 rows from a
// register window of the footprint rows (7 / 10 / 3 shared-memory loads per 4 outputs and tap row instead of 16), all index
// arithmetic (the op's floor divisions) is hoisted out of the pixel loops, taps live in registers (up = 1) or are read as
// broadcast LDS (up = 2), and the 4 outputs leave as one 16- or 8-byte store.
constexpr int kP4TileH = 32, kP4TileW = 128;

template <typename T> struct OutVec4;
template <> struct OutVec4<float> {
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
  static constexpr int kAlign = 16;
};
template <> struct OutVec4<__nv_bfloat16> {
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  }
  static constexpr int kAlign = 8;
};
template <> struct OutVec4<__half> {
  static __device__ __forceinline__ void store(__half* p, const float (&v)[4]) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  }
  static constexpr int kAlign = 8;
};

template <typename T, int UP, int DOWN>
__global__ void __launch_bounds__(256) upfirdn2d_planar4_kernel(T* __restrict__ y, const T* __restrict__ x,
                                                                const float* __restrict__ kernel, UpfirdnParams p,
                                                                int tiles_x, int tiles_y, int fh, int fw) {
  extern __shared__ float smem[];
  float* sk = smem;        // 16 flipped taps
  float* sx = smem + 16;   // fh x fw input footprint
  if (threadIdx.x < 16) sk[threadIdx.x] = kernel[15 - threadIdx.x];   // flip both axes of the 4 x 4 kernel
  constexpr int NT = 4 / UP;                 // taps per axis that hit a real input sample
  constexpr int SPAN = (3 * DOWN) / UP + NT + (UP > 1 ? 1 : 0);   // footprint columns touched by 4 adjacent outputs
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t tiles_per_plane = (int64_t)tiles_x * tiles_y;
  const int64_t ntiles = tiles_per_plane * p.major;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t plane = tile / tiles_per_plane;
    const int tr = (int)(tile - plane * tiles_per_plane);
    const int tyi = tr / tiles_x, txi = tr - tyi * tiles_x;
    const int oy0 = tyi * kP4TileH, ox0 = txi * kP4TileW;
    const int mid_y0 = oy0 * DOWN + UP - 1 - p.pad_y0;
    const int mid_x0 = ox0 * DOWN + UP - 1 - p.pad_x0;
    const int in_y_base = floor_div_i(mid_y0, UP);
    const int in_x_base = floor_div_i(mid_x0, UP);
    const T* xp = x + plane * (int64_t)p.in_h * p.in_w;
    __syncthreads();
    for (int ry = ty; ry < fh; ry += 8) {
      const int iy = in_y_base + ry;
      const bool rok = iy >= 0 && iy < p.in_h;
      const T* xr = xp + (int64_t)iy * p.in_w;
      for (int rx = tx; rx < fw; rx += 32) {
        const int ix = in_x_base + rx;
        sx[ry * fw + rx] = (rok && ix >= 0 && ix < p.in_w) ? to_f32<T>(xr[ix]) : 0.f;
      }
    }
    __syncthreads();
    // this thread's 4 adjacent output columns: first footprint column and first tap of each
    int cx[4], kx0[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int mid_x = mid_x0 + (4 * tx + j) * DOWN;
      const int in_x = floor_div_i(mid_x, UP);
      kx0[j] = (in_x + 1) * UP - mid_x - 1;
      cx[j] = in_x - in_x_base;
    }
    float kreg[16];
    if (UP == 1) {
#pragma unroll
      for (int t = 0; t < 16; ++t) kreg[t] = sk[t];
    }
    T* yp = y + plane * (int64_t)p.out_h * p.out_w;
    const int ox = ox0 + 4 * tx;
    if (UP == 1) {
      // 4 x 4 output block per thread: each footprint row is read once into a register window and feeds up to four output rows
      // (49 / 100 shared-memory loads per 16 outputs for down = 1 / 2 instead of 112 / 160)
      constexpr int RSPAN = 3 * DOWN + 4;
      const int oyb = oy0 + 4 * ty;
      if (oyb < p.out_h && ox < p.out_w) {
        const float* srow = sx + (4 * ty * DOWN) * fw + cx[0];      // in_y - in_y_base = ry * DOWN for up = 1
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[a][j] = 0.f;
#pragma unroll
        for (int r = 0; r < RSPAN; ++r) {
          float win[SPAN];
#pragma unroll
          for (int c = 0; c < SPAN; ++c) win[c] = srow[r * fw + c];
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int t = r - a * DOWN;                              // tap row of output row a (compile time)
            if (t >= 0 && t < 4) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[a][j] = fmaf(win[j * DOWN + u], kreg[t * 4 + u], acc[a][j]);
            }
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int oy = oyb + a;
          if (oy >= p.out_h) break;
          T* dst = yp + (int64_t)oy * p.out_w + ox;
          if (ox + 3 < p.out_w && ((uintptr_t)dst % OutVec4<T>::kAlign) == 0) {
            OutVec4<T>::store(dst, acc[a]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (ox + j < p.out_w) dst[j] = from_f32<T>(acc[a][j]);
          }
        }
      }
      continue;
    }
#pragma unroll
    for (int rr = 0; rr < kP4TileH / 8; ++rr) {
      const int ry = ty + 8 * rr;
      const int oy = oy0 + ry;
      if (oy >= p.out_h || ox >= p.out_w) continue;
      const int mid_y = mid_y0 + ry * DOWN;
      const int in_y = floor_div_i(mid_y, UP);
      const int ky0 = (in_y + 1) * UP - mid_y - 1;
      const float* srow = sx + (in_y - in_y_base) * fw + cx[0];
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        float win[SPAN];
#pragma unroll
        for (int c = 0; c < SPAN; ++c) win[c] = srow[t * fw + c];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int u = 0; u < NT; ++u) {
            // up = 2: column offset and first tap depend on the parity of mid_x (runtime, fixed per thread and j)
            const int off = cx[j] - cx[0] + u;
            float v = win[0];
#pragma unroll
            for (int c = 1; c < SPAN; ++c) v = off == c ? win[c] : v;
            acc[j] = fmaf(v, sk[(ky0 + t * UP) * 4 + kx0[j] + u * UP], acc[j]);
          }
        }
      }
      T* dst = yp + (int64_t)oy * p.out_w + ox;
      if (ox + 3 < p.out_w && ((uintptr_t)dst % OutVec4<T>::kAlign) == 0) {
        OutVec4<T>::store(dst, acc);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ox + j < p.out_w) dst[j] = from_f32<T>(acc[j]);
      }
    }
  }
}

template <typename T, int UP, int DOWN>
static int launch_planar4(void* y, const void* x, const float* kernel, const UpfirdnParams& p, cudaStream_t st) {
  const int fh = ((kP4TileH - 1) * DOWN + 3) / UP + 2;
  const int fw = ((kP4TileW - 1) * DOWN + 3) / UP + 2;
  const size_t smem = sizeof(float) * (16 + (size_t)fh * fw);
  auto kern = upfirdn2d_planar4_kernel<T, UP, DOWN>;
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    L2I_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int tiles_x = ceil_div(p.out_w, kP4TileW), tiles_y = ceil_div(p.out_h, kP4TileH);
  const int64_t ntiles = (int64_t)tiles_x * tiles_y * p.major;
  const int per_sm = smem > 48 * 1024 ? 3 : 8;
  const int blocks = (int)std::min<int64_t>(ntiles, (int64_t)kNumSMs * per_sm);
  kern<<<blocks, 256, smem, st>>>((T*)y, (const T*)x, kernel, p, tiles_x, tiles_y, fh, fw);
  return check_launch("upfirdn2d_planar4");
}

template <typename T>
static int launch_upfirdn(void* y, const void* x, const float* kernel, const UpfirdnParams& p, cudaStream_t st) {
  const int64_t total = p.major * p.out_h * (int64_t)p.out_w * p.minor;
  if (total == 0) return L2I_OK;
  if (p.minor == 1 && p.kh == 4 && p.kw == 4 && p.up_x == p.up_y && p.down_x == p.down_y) {
    if (p.up_x == 1 && p.down_x == 1) return launch_planar4<T, 1, 1>(y, x, kernel, p, st);
    if (p.up_x == 2 && p.down_x == 1) return launch_planar4<T, 2, 1>(y, x, kernel, p, st);
    if (p.up_x == 1 && p.down_x == 2) return launch_planar4<T, 1, 2>(y, x, kernel, p, st);
  }
  if (p.minor == 1) {
    const int fh = ((kTileH - 1) * p.down_y + p.kh - 1) / p.up_y + 2;
    const int fw = ((kTileW - 1) * p.down_x + p.kw - 1) / p.up_x + 2;
    const size_t smem = sizeof(float) * (64 + (size_t)fh * fw);
    if (smem <= 48 * 1024) {
      const int tiles_x = ceil_div(p.out_w, kTileW), tiles_y = ceil_div(p.out_h, kTileH);
      const int64_t ntiles = (int64_t)tiles_x * tiles_y * p.major;
      const int blocks = (int)std::min<int64_t>(ntiles, (int64_t)kNumSMs * 8);
      upfirdn2d_planar_kernel<T><<<blocks, 256, smem, st>>>((T*)y, (const T*)x, kernel, p, tiles_x, tiles_y, fh, fw);
      return check_launch("upfirdn2d_planar");
    }
  }
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 16);
  upfirdn2d_gather_kernel<T><<<blocks, 256, 0, st>>>((T*)y, (const T*)x, kernel, p);
  return check_launch("upfirdn2d_gather");
}

}  // namespace l2i

using namespace l2i;

extern "C" int l2i_upfirdn2d(void* y, const void* x, const float* kernel, int64_t major, int in_h, int in_w,
                             int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                             int pad_x0, int pad_x1, int pad_y0, int pad_y1, int dtype, void* stream) {
  L2I_REQUIRE(major >= 0 && in_h >= 0 && in_w >= 0 && minor >= 1, "upfirdn2d: bad input shape");
  L2I_REQUIRE(kh >= 1 && kw >= 1 && kh <= 8 && kw <= 8, "upfirdn2d: kernel %dx%d not in 1..8", kh, kw);
  L2I_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down must be >= 1");
  UpfirdnParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  // op/upfirdn2d_kernel.cu:167-168
  int num_h = in_h * up_y + pad_y0 + pad_y1 - kh + down_y;
  int num_w = in_w * up_x + pad_x0 + pad_x1 - kw + down_x;
  p.out_h = num_h > 0 ? num_h / down_y : 0;
  p.out_w = num_w > 0 ? num_w / down_x : 0;
  if (major == 0 || p.out_h == 0 || p.out_w == 0) return L2I_OK;
  L2I_REQUIRE(y && x && kernel, "upfirdn2d: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case L2I_F32: return launch_upfirdn<float>(y, x, kernel, p, st);
    case L2I_BF16: return launch_upfirdn<__nv_bfloat16>(y, x, kernel, p, st);
    case L2I_F16: return launch_upfirdn<__half>(y, x, kernel, p, st);
    default: set_error("upfirdn2d: unsupported dtype %d", dtype); return L2I_ERR_UNSUPPORTED;
  }
}
