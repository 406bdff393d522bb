// 2x2-block ("quad") tcgen05 implicit-GEMM conv for the 32 -> 32 channel 3x3 layer at the top resolution.
//
// A tcgen05.mma with M = 128 streams its A rows at 2 rows / clock whatever N is (measured: ~64-70 clocks per
// M128 x K16 instruction for every N <= 128), so a 32-channel conv issued tap by tap (N = 32) runs the tensor
// pipe at a quarter of its rate.  Here one GEMM row is a 2x2 block of OUTPUT pixels:
//     N = 128 = (a, b, co)       a, b in {0,1}: pixel inside the block, co: 32 output channels
//     K = 512 = (r, c, ci)       r, c in 0..3: the 4x4 input patch the block reads, ci: 32 input channels
//     Wq[(a,b,co)][(r,c,ci)] = W[co][ci][r-a][c-b]   (zero outside the 3x3 support: 9/16 dense)
// i.e. 16384 issued MACs per pixel at the full N = 128 rate (4 clocks / pixel) instead of 12288 at a quarter of it.
// The A operand needs no im2col: the activation is plain NHWC with 32 bf16 = 64 bytes per pixel, so the four
// pixels of a patch row are 256 contiguous bytes and horizontally adjacent blocks are 128 bytes apart - exactly the
// row pitch of the K-major SWIZZLE_128B layout.  One 4-D TMA box load per tile lands the (32+2) x (16+4) pixel halo
// tile in shared memory (out-of-image pixels zero-filled = the conv padding); the UMMA descriptor of patch row r,
// K step kk starts at  tile + r * row_pitch + 64 + kk * 32 B  with the 8-row-group stride (SBO) = two image rows.
// The 128 KB weight matrix stays resident in shared memory for the lifetime of the persistent CTA.
//
// Epilogue (per thread = one block = four pixels x 32 channels): demod, noise, bias, leaky-relu, ToRGB 1x1,
// bias + 2x FIR up-sampling of the skip image (ToRGB.forward, networks.py:349-358), float2 stores.
#include "tc_epilogue.cuh"

namespace l2i {

using namespace tc;

namespace {

constexpr int kQTileW = 16, kQTileH = 32;                 // output pixels per tile (8 x 16 blocks = 128 GEMM rows)
// The halo tile is loaded as 10 horizontal pixel PAIRS per row starting at x0 - 2, so the TMA box's inner extent is
// exactly the 128-byte swizzle span; the left halo pixel x0 - 1 sits 64 bytes into each row.
constexpr int kQHaloPairs = kQTileW / 2 + 2, kQHaloH = kQTileH + 2;
constexpr int kQRowPitch = kQHaloPairs * 128;             // 1280 bytes per halo image row
constexpr int kQHaloBytes = kQHaloH * kQRowPitch;         // 43520
constexpr int kQStageBytes = (kQHaloBytes + 1023) & ~1023;
constexpr int kQStages = 2;
constexpr int kQGroups = 3;                               // 512 threads -> 128 registers per thread, no spills
constexpr int kQTmemCols = 512;                           // 3 accumulators x 128 columns, rounded up to a power of two
constexpr int kQThreads = 128 + kQGroups * 128;
constexpr int kQN = 128, kQK = 512;
constexpr int kQWBytes = kQN * kQK * 2;                   // 131072
constexpr int kQSmem = kQStages * kQStageBytes + kQWBytes + 1024;

struct QuadParams {
  int B, H, W;
  int tiles_x, tiles_y, total_tiles;
  uint32_t idesc;
  EpiParams e;
};

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void quad_group_sync(int group) {
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

__global__ void __launch_bounds__(kQThreads, 1)
conv_tc_quad_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ QuadParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_w = smem + kQStages * kQStageBytes;
  __shared__ __align__(16) float epi_smem[kQGroups * 6 * 32];
  __shared__ __align__(8) uint64_t full_bar[kQStages];
  __shared__ __align__(8) uint64_t empty_bar[kQStages];
  __shared__ __align__(8) uint64_t tmem_full[kQGroups];
  __shared__ __align__(8) uint64_t tmem_empty[kQGroups];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kQStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < kQGroups; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
    mbar_init(&w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, kQTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  auto decode = [&](int tile, int& x0, int& y0, int& b) {
    const int tx = tile % p.tiles_x;
    const int r = tile / p.tiles_x;
    const int ty = r % p.tiles_y;
    b = r / p.tiles_y;
    x0 = tx * kQTileW; y0 = ty * kQTileH;
  };

  if (warp == 0) {
    // ===================== TMA producer: the weight matrix once, then one halo tile per output tile =====
    if (lane == 0) {
      mbar_expect_tx(&w_bar, (uint32_t)kQWBytes);
      for (int t = 0; t < kQK / 64; ++t) tma_load_3d(smem_w + t * (kQN * 128), &tmap_w, &w_bar, 0, 0, t);
      int stage = 0;
      uint32_t phase_bit = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int x0, y0, b;
        decode(tile, x0, y0, b);
        mbar_wait(&empty_bar[stage], phase_bit ^ 1);
        mbar_expect_tx(&full_bar[stage], kQHaloBytes);
        tma_load_4d(smem + stage * kQStageBytes, &tmap_a, &full_bar[stage], 0, (x0 >> 1) - 1, y0 - 1, b);
        if (++stage == kQStages) { stage = 0; phase_bit ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 4 patch rows x 8 K16 steps, one N = 128 accumulator per tile =====
    if (lane == 0) {
      mbar_wait(&w_bar, 0);
      tc_fence_after();
      const uint32_t w_base = smem_u32(smem_w);
      int stage = 0;
      uint32_t phase_bit = 0;
      int grp = 0;
      uint32_t grp_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[grp], grp_phase ^ 1);
        mbar_wait(&full_bar[stage], phase_bit);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + stage * kQStageBytes);
        const uint32_t tmem_d = tmem_base + (uint32_t)(grp * kQN);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            // A: rows = blocks along x (128 B apart), 8-row groups = block rows (two image rows apart)
            const uint64_t adesc = make_desc_sw128(a_base + (uint32_t)(r * kQRowPitch + 64 + kk * 32), 2 * kQRowPitch);
            // B: K-major [128 x 64] atoms, atom = (r*128 + kk*16) / 64
            const uint64_t bdesc = make_desc_sw128(w_base + (uint32_t)((r * 2 + (kk >> 2)) * (kQN * 128) + (kk & 3) * 32), 1024);
            umma_bf16(tmem_d, adesc, bdesc, p.idesc, (r | kk) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[grp]);
        if (++stage == kQStages) { stage = 0; phase_bit ^= 1; }
        if (++grp == kQGroups) { grp = 0; grp_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const EpiParams& e = p.e;
    const int group = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gtid = threadIdx.x - (128 + group * 128);
    float* sp = epi_smem + group * (6 * 32);
    float* s_d = sp;
    float* s_b = sp + 32;
    float* s_n = sp + 64;
    float* s_w = sp + 96;
    constexpr float kSqrt2 = 1.4142135623730951f;
    const float nw = (e.noise != nullptr && e.noise_w != nullptr) ? __ldg(e.noise_w) * kSqrt2 : 0.f;
    const int64_t plane = (int64_t)p.H * p.W;
    const int bx = row & 7, by = row >> 3;
    const int h2 = p.H >> 1, w2 = p.W >> 1;
    uint32_t grp_phase = 0;
    int staged_b = -1;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      if (it % kQGroups != group) continue;
      int x0, y0, b;
      decode(tile, x0, y0, b);
      const int X = x0 + 2 * bx, Y = y0 + 2 * by;          // top-left pixel of this thread's block
      const bool ok = X < p.W && Y < p.H;                   // H, W are even: a block is inside or outside as a whole

      if (b != staged_b) {
        quad_group_sync(group);
        for (int j = gtid; j < 32; j += 128) {
          s_d[j] = (e.demod != nullptr ? __ldg(e.demod + (int64_t)b * e.demod_bs + j) : 1.f) * kSqrt2;
          s_b[j] = __ldg(e.bias + j) * kSqrt2;
          s_n[j] = e.s_next ? __ldg(e.s_next + (int64_t)b * e.s_next_bs + j) : 1.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) s_w[c * 32 + j] = e.wr ? __ldg(e.wr + (int64_t)b * e.wr_bs + c * 32 + j) : 0.f;
        }
        quad_group_sync(group);
        staged_b = b;
      }

      // ---- all global loads of the block before waiting for the accumulator ----
      float nzq[4] = {0.f, 0.f, 0.f, 0.f};
      float up[4][3];
#pragma unroll
      for (int i = 0; i < 4; ++i) up[i][0] = up[i][1] = up[i][2] = 0.f;
      if (ok) {
        if (e.noise != nullptr) {
          const float* np = e.noise + (int64_t)b * e.noise_bs + (int64_t)Y * p.W + X;
          const float2 n01 = __ldg(reinterpret_cast<const float2*>(np));
          const float2 n23 = __ldg(reinterpret_cast<const float2*>(np + p.W));
          nzq[0] = nw * n01.x; nzq[1] = nw * n01.y; nzq[2] = nw * n23.x; nzq[3] = nw * n23.y;
        }
        if (e.fused_skip) {
          // 2x FIR up-sampling of the skip image for the four pixels of the block: they share the 3x3 low-res
          // patch around (m, n) = (Y/2, X/2); even outputs use rows m-1 (f0), m (f2), odd ones m (f1), m+1 (f3)
          const int m = Y >> 1, n = X >> 1;
          const float wy[2][3] = {{m > 0 ? e.fir[0] : 0.f, e.fir[2], 0.f}, {0.f, e.fir[1], m + 1 < h2 ? e.fir[3] : 0.f}};
          const float wx[2][3] = {{n > 0 ? e.fir[0] : 0.f, e.fir[2], 0.f}, {0.f, e.fir[1], n + 1 < w2 ? e.fir[3] : 0.f}};
          const int ym = max(m - 1, 0), yp = min(m + 1, h2 - 1), xm = max(n - 1, 0), xp = min(n + 1, w2 - 1);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float bias_c = __ldg(e.rgb_bias + c);
            float pv[3][3];
            if (e.skip_in != nullptr) {
              const float* pl = e.skip_in + ((int64_t)b * 3 + c) * (plane >> 2);
              const int ys[3] = {ym, m, yp}, xs[3] = {xm, n, xp};
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) pv[i][j] = __ldg(pl + (int64_t)ys[i] * w2 + xs[j]);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
              for (int bb = 0; bb < 2; ++bb) {
                float acc = bias_c;
                if (e.skip_in != nullptr) {
#pragma unroll
                  for (int i = 0; i < 3; ++i) {
                    const float hrow = wx[bb][0] * pv[i][0] + wx[bb][1] * pv[i][1] + wx[bb][2] * pv[i][2];
                    acc = fmaf(wy[a][i], hrow, acc);
                  }
                }
                up[a * 2 + bb][c] = acc;
              }
          }
        }
      }

      mbar_wait(&tmem_full[group], grp_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(group * kQN);
      float rgb[4][3];
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) {   // 32-column chunk ph = pixel (a, b) = (ph >> 1, ph & 1) of the block
        uint32_t v[32];
        tmem_ld32(taddr + ph * 32, v);
        __nv_bfloat16* outc = nullptr;
        __nv_bfloat16* yc = nullptr;
        if (ok) {
          const int64_t pix = ((int64_t)b * p.H + Y + (ph >> 1)) * p.W + X + (ph & 1);
          if (e.out != nullptr && e.s_next != nullptr) outc = (__nv_bfloat16*)e.out + pix * 32;
          if (e.y_out != nullptr) yc = (__nv_bfloat16*)e.y_out + pix * 32;
        }
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        tmem_ld_wait();
        epilogue_chunk32<EPI_ACT_RGB>(v, s_d, s_b, s_n, s_w, s_w + 32, s_w + 64, nzq[ph], false, r0, r1, r2, outc, yc);
        rgb[ph][0] = r0; rgb[ph][1] = r1; rgb[ph][2] = r2;
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[group]);
      grp_phase ^= 1;

      if (e.wr != nullptr && ok) {
        float* dst = e.fused_skip ? e.skip_out : e.rgb_part;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float* pl = dst + ((int64_t)b * 3 + c) * plane + (int64_t)Y * p.W + X;
          *reinterpret_cast<float2*>(pl) = make_float2(rgb[0][c] + up[0][c], rgb[1][c] + up[1][c]);
          *reinterpret_cast<float2*>(pl + p.W) = make_float2(rgb[2][c] + up[2][c], rgb[3][c] + up[3][c]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kQTmemCols);
  }
}

// dst [8 atoms][128 n][64 k] bf16 with n = (a*2+b)*32 + co, k_global = atom*64 + k = (r*4 + c)*32 + ci;
// src [32 co][32 ci][3][3] fp32
__global__ void pack_quad_weight_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, float scale) {
  const int total = kQN * kQK;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx % 64, n = (idx / 64) % kQN, atom = idx / (64 * kQN);
    const int kg = atom * 64 + k;
    const int ci = kg % 32, c = (kg / 32) % 4, r = kg / 128;
    const int co = n % 32, b = (n / 32) % 2, a = n / 64;
    const int kh = r - a, kw = c - b;
    float v = 0.f;
    if (kh >= 0 && kh < 3 && kw >= 0 && kw < 3) v = src[(((int64_t)co * 32 + ci) * 3 + kh) * 3 + kw] * scale;
    dst[idx] = __float2bfloat16_rn(v);
  }
}

}  // namespace

// The 2x2-block kernel takes plain (non-transposed) 32 -> 32 convs with the act + ToRGB epilogue on even-sized images
// stored as plain NHWC (64 bytes per pixel).
bool conv_tc_quad_supported(const ConvGeom& g, const EpiParams& e) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* env = std::getenv("L2I_QUAD");
    enabled = (env != nullptr && env[0] == '0') ? 0 : 1;
  }
  if (!enabled || !tmap_available()) return false;
  if (g.nphase != 1 || g.in_scale != 1 || g.up_cout != 0 || g.in_pair_packed) return false;
  if (g.Cin != 32 || g.Cout != 32 || e.mode != 0 || e.wr == nullptr) return false;
  if (g.H < 32 || g.W < 16 || (g.H & 1) || (g.W & 1) || g.OH != g.H || g.OW != g.W) return false;
  if (!e.fused_skip) return false;
  return true;
}

int launch_pack_quad_weight(__nv_bfloat16* dst, const float* src, float scale, cudaStream_t st) {
  pack_quad_weight_kernel<<<ceil_div(kQN * kQK, 256), 256, 0, st>>>(dst, src, scale);
  return check_launch("pack_quad_weight");
}

// w: the [8][128][64] bf16 matrix written by launch_pack_quad_weight
int launch_conv_tc_quad(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st) {
  QuadParams p{};
  p.B = g.B; p.H = g.H; p.W = g.W; p.e = e;
  p.idesc = make_idesc_bf16(128, kQN, 0);
  CUtensorMap ta, tw;
  {
    const uint64_t dims[4] = {64, (uint64_t)g.W / 2, (uint64_t)g.H, (uint64_t)g.B};   // horizontal pixel pairs
    const uint64_t str[4] = {2, 128, (uint64_t)g.W * 64, (uint64_t)g.H * g.W * 64};
    const uint32_t box[4] = {64, kQHaloPairs, kQHaloH, 1};
    L2I_TRY(make_tmap(&ta, in, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    const uint64_t dims[3] = {64, kQN, kQK / 64};
    const uint64_t str[3] = {2, 128, (uint64_t)kQN * 128};
    const uint32_t box[3] = {64, kQN, 1};
    L2I_TRY(make_tmap(&tw, w, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  p.tiles_x = ceil_div(g.W, kQTileW); p.tiles_y = ceil_div(g.H, kQTileH);
  const int64_t total = (int64_t)p.tiles_x * p.tiles_y * g.B;
  if (total <= 0 || total > 0x7fffffff) { set_error("conv_tc_quad: bad tile count"); return L2I_ERR_INVALID_ARG; }
  p.total_tiles = (int)total;
  static bool attr_set = false;
  if (!attr_set) {
    L2I_CUDA_TRY(cudaFuncSetAttribute(conv_tc_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQSmem));
    attr_set = true;
  }
  const int grid = std::min(p.total_tiles, kNumSMs);
  conv_tc_quad_kernel<<<grid, kQThreads, kQSmem, st>>>(ta, tw, p);
  return check_launch("conv_tc_quad");
}

}  // namespace l2i
