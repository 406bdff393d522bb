// Vertical-pair tcgen05 implicit-GEMM conv for the plain 64 -> 64 channel 3x3 layer (512 x 512 px).
//
// Issued tap by tap this layer is N = 64 per MMA, and an M = 128 tcgen05.mma costs ~64-77 clocks per K = 16 step
// whatever N <= 128 is (A rows stream at 2 / clock), so the tensor pipe runs at half rate (conv_tc_halo.cu: 36 MMAs per
// 128 pixels).  Here one GEMM row is a VERTICAL PAIR of output pixels (rows 2j, 2j+1 of one column):
//     N = 128 = (parity, co);  the pair reads input rows 2j-1+r, r = 0..3, and columns x-1+kw, kw = 0..2
//     parity p takes input row r with kernel row kh = r - p  ->  r = 0: (p0, kh0);  r = 1: (p0, kh1), (p1, kh0);
//                                                               r = 2: (p0, kh2), (p1, kh1);  r = 3: (p1, kh2)
// i.e. 12 (r, kw) steps x 4 K16 = 48 MMAs per 256 pixels instead of 72.  The B operand needs no zero padding and no
// duplicated weights: the nine 64 x 64 weight tiles are stored per kw in the order W[kh=2], W[kh=1], W[kh=0], so the
// N = 128 operand of r = 1 is the 16 KB window starting at W[kh=1] and that of r = 2 the window starting at W[kh=2]
// (overlapping UMMA descriptors); r = 0 and r = 3 are N = 64 MMAs into the lower / upper half of the accumulator.
// A comes straight from the NHWC halo tile (TMA, once per output tile): GEMM rows of one 8-row group are 8 adjacent
// pixels (128 B apart), groups are two image rows apart (stride-byte-offset = 2 x halo pitch).
//
// The layer is epilogue-bound (demod, noise, bias, leaky-relu, ToRGB and the skip up-sampling for 256 pixels x 64
// channels per tile), so the epilogue issues no global gathers: warp 3 TMA-loads the tile's noise patch (32 x 8 fp32)
// and low-resolution skip patch (3 x 18 x 12 fp32, out-of-image samples zero-filled = upfirdn2d's zero padding) into a
// six-deep ring, two tiles ahead of each epilogue warpgroup, and the MMA issuer has four accumulators to run ahead in.
#include "tc_epilogue.cuh"

namespace l2i {

using namespace tc;

namespace {

constexpr int kVW = 10, kVH = 34;                          // halo tile: (8 + 2) x (32 + 2) pixels of 64 channels
constexpr int kVHaloBytes = kVW * kVH * 128;               // 43520
constexpr int kVStageBytes = (kVHaloBytes + 1023) & ~1023;
constexpr int kVStages = 2;
constexpr int kVGroups = 3;                               // 512 threads -> 128 registers per thread
constexpr int kVAccs = 4;                                 // accumulators of 128 TMEM columns
constexpr int kVPatches = 2 * kVGroups;                   // noise / skip patch ring (tile it -> buffer it % 6, always the same warpgroup)
constexpr int kVSkipW = 12, kVSkipH = 18;                 // low-res skip patch: columns n0-4 .. n0+7 (16-byte aligned start), rows m0-1 .. m0+16
constexpr int kVSkipFloats = 3 * kVSkipH * kVSkipW;       // 648
constexpr int kVSkipStride = (kVSkipFloats + 31) & ~31;   // 672 floats: buffers stay 128-byte aligned
constexpr int kVNoiseFloats = 32 * 8;                     // the tile's 32 rows x 8 columns
constexpr int kVThreads = 128 + kVGroups * 128;
constexpr int kVWBytes = 9 * 64 * 128;                     // 73728: nine 64 x 64 bf16 tiles
constexpr int kVStageOutBytes = 2048;                      // per epilogue warp: [parity][32 pixels][16 channels] bf16 (TMA store sources)
constexpr int kVSmem = kVStages * kVStageBytes + kVWBytes + kVGroups * 4 * kVStageOutBytes + 1024;
constexpr int kVTileW = 8, kVTileH = 32;

struct VPairParams {
  int B, H, W;
  int tiles_x, tiles_y, total_tiles;
  uint32_t idesc128, idesc64;
  int tma_store;             // activation output goes through per-warp staging tiles + strided TMA tensor stores
  int has_skip, has_noise;   // tmap_s / tmap_n are valid: warp 3 loads the patches of every tile
  int noise_per_sample;      // 0: one noise image broadcast over the batch
  EpiParams e;
};

__device__ __forceinline__ uint64_t vp_desc(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void vp_group_sync(int group) {
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

__global__ void __launch_bounds__(kVThreads, 1)
conv_tc_vpair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                     const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_s,
                     const __grid_constant__ CUtensorMap tmap_n, const __grid_constant__ VPairParams p) {
  constexpr int N = 128, CO = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_w = smem + kVStages * kVStageBytes;
  uint8_t* smem_stage_out = smem_w + kVWBytes;
  __shared__ __align__(16) float epi_smem[kVGroups * 6 * CO];
  __shared__ __align__(8) uint64_t full_bar[kVStages];
  __shared__ __align__(8) uint64_t empty_bar[kVStages];
  __shared__ __align__(8) uint64_t tmem_full[kVAccs];
  __shared__ __align__(8) uint64_t tmem_empty[kVAccs];
  __shared__ __align__(8) uint64_t patch_full[kVPatches];
  __shared__ __align__(8) uint64_t patch_empty[kVPatches];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(128) float skip_smem[kVPatches * kVSkipStride];
  __shared__ __align__(128) float noise_smem[kVPatches * kVNoiseFloats];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    if (p.tma_store) prefetch_tmap(&tmap_o);
    if (p.has_skip) prefetch_tmap(&tmap_s);
    if (p.has_noise) prefetch_tmap(&tmap_n);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kVStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < kVAccs; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
    for (int a = 0; a < kVPatches; ++a) { mbar_init(&patch_full[a], 1); mbar_init(&patch_empty[a], 128); }
    mbar_init(&w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  auto decode = [&](int tile, int& x0, int& y0, int& b) {
    const int tx = tile % p.tiles_x;
    const int r = tile / p.tiles_x;
    const int ty = r % p.tiles_y;
    b = r / p.tiles_y;
    x0 = tx * kVTileW; y0 = ty * kVTileH;
  };

  if (warp == 0) {
    // ===================== TMA producer: the nine weight tiles once, then one halo tile per output tile ===
    if (lane == 0) {
      mbar_expect_tx(&w_bar, (uint32_t)kVWBytes);
      for (int t = 0; t < 9; ++t) tma_load_3d(smem_w + t * (64 * 128), &tmap_w, &w_bar, 0, 0, t);
      int stage = 0;
      uint32_t phase_bit = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int x0, y0, b;
        decode(tile, x0, y0, b);
        mbar_wait(&empty_bar[stage], phase_bit ^ 1);
        mbar_expect_tx(&full_bar[stage], kVHaloBytes);
        tma_load_4d(smem + stage * kVStageBytes, &tmap_a, &full_bar[stage], 0, x0 - 1, y0 - 1, b);
        if (++stage == kVStages) { stage = 0; phase_bit ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================== patch producer: the tile's noise and low-res skip boxes, up to six tiles ahead ========
    if (lane == 0 && (p.has_skip || p.has_noise)) {
      int buf = 0;
      uint32_t buf_phase = 0;
      const uint32_t bytes = (p.has_skip ? kVSkipFloats * 4u : 0u) + (p.has_noise ? kVNoiseFloats * 4u : 0u);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int x0, y0, b;
        decode(tile, x0, y0, b);
        mbar_wait(&patch_empty[buf], buf_phase ^ 1);   // the epilogue that used this buffer six tiles ago has read it
        mbar_expect_tx(&patch_full[buf], bytes);
        if (p.has_skip) tma_load_4d(skip_smem + buf * kVSkipStride, &tmap_s, &patch_full[buf], (x0 >> 1) - 4, (y0 >> 1) - 1, 0, b);
        if (p.has_noise) tma_load_3d(noise_smem + buf * kVNoiseFloats, &tmap_n, &patch_full[buf], x0, y0, p.noise_per_sample ? b : 0);
        if (++buf == kVPatches) { buf = 0; buf_phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 48 MMAs per tile, no barrier inside a tile =====================
    // warp-uniform control flow, one elected lane issues (tc_ptx.cuh: elect_one)
    {
      mbar_wait(&w_bar, 0);
      tc_fence_after();
      const uint32_t w_base = smem_u32(smem_w), smem_a0 = smem_u32(smem);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint64_t kHiA = kmajor_desc_hi(2 * kVW * 128, 2), kHiB = kmajor_desc_hi(1024, 2);
      const uint64_t b_desc0 = kmajor_desc_at(kHiB, w_base);
      int stage = 0, acc = 0;
      uint32_t phase_bit = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        mbar_wait(&full_bar[stage], phase_bit);
        tc_fence_after();
        const uint64_t a_desc0 = kmajor_desc_at(kHiA, smem_a0 + (uint32_t)(stage * kVStageBytes));
        const uint32_t tmem_d = tmem_u + (uint32_t)(acc * N);
        if (elect_one()) {
          // input-row order 1, 2, 0, 3: the first MMA (r = 1, N = 128) initialises all 128 accumulator columns
#pragma unroll
          for (int ri = 0; ri < 4; ++ri) {
            const int r = ri == 0 ? 1 : (ri == 1 ? 2 : (ri == 2 ? 0 : 3));
            // weight window: slots per kw are W[kh=2], W[kh=1], W[kh=0]; parity 0 uses kh = r, parity 1 kh = r - 1
            const int slot = r == 0 ? 2 : (r == 1 ? 1 : 0);            // first 64-row slot of the operand
            const bool wide = (r == 1 || r == 2);                      // N = 128 (both parities) or N = 64
            const uint32_t d_col = r == 3 ? 64u : 0u;                  // r = 3 feeds parity 1 only
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_d + d_col, a_desc0 + (uint64_t)(((r * kVW + kw) * 128 + k * 32) >> 4),
                          b_desc0 + (uint64_t)(((kw * 3 + slot) * (64 * 128) + k * 32) >> 4), wide ? p.idesc128 : p.idesc64,
                          (ri | kw | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&tmem_full[acc]);
        }
        if (++stage == kVStages) { stage = 0; phase_bit ^= 1; }
        if (++acc == kVAccs) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread = one column x two rows (parity = accumulator half) ==========
    const EpiParams& e = p.e;
    const int group = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gtid = threadIdx.x - (128 + group * 128);
    float* sp = epi_smem + group * (6 * CO);
    float* s_d = sp;
    float* s_b = sp + CO;
    float* s_n = sp + 2 * CO;
    float* s_w = sp + 3 * CO;
    constexpr float kSqrt2 = 1.4142135623730951f;
    const float nw = (e.noise != nullptr && e.noise_w != nullptr) ? __ldg(e.noise_w) * kSqrt2 : 0.f;
    const int64_t plane = (int64_t)p.H * p.W;
    const int lx = row & 7, lj = row >> 3;
    int staged_b = -1;
    int it = group;
    for (int tile = blockIdx.x + group * gridDim.x; tile < p.total_tiles; tile += kVGroups * gridDim.x, it += kVGroups) {
      int x0, y0, b;
      decode(tile, x0, y0, b);
      const int acc = it & (kVAccs - 1), buf = it % kVPatches;
      const uint32_t acc_parity = (uint32_t)(it / kVAccs) & 1u, buf_parity = (uint32_t)(it / kVPatches) & 1u;
      const int X = x0 + lx, Y = y0 + 2 * lj;              // upper pixel of the pair
      const bool ok = X < p.W && Y < p.H;                   // H is even: both pixels inside or both outside

      if (b != staged_b) {
        vp_group_sync(group);
        for (int j = gtid; j < CO; j += 128) {
          s_d[j] = (e.demod != nullptr ? __ldg(e.demod + (int64_t)b * e.demod_bs + j) : 1.f) * kSqrt2;
          s_b[j] = __ldg(e.bias + j) * kSqrt2;
          s_n[j] = e.s_next ? __ldg(e.s_next + (int64_t)b * e.s_next_bs + j) : 1.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) s_w[c * CO + j] = e.wr ? __ldg(e.wr + (int64_t)b * e.wr_bs + c * CO + j) : 0.f;
        }
        vp_group_sync(group);
        staged_b = b;
      }

      float nzp[2] = {0.f, 0.f};
      if (p.has_skip || p.has_noise) mbar_wait(&patch_full[buf], buf_parity);
      if (p.has_noise) {
        const float* ns = noise_smem + buf * kVNoiseFloats + (2 * lj) * 8 + lx;
        nzp[0] = nw * ns[0];
        nzp[1] = nw * ns[8];
      }

      mbar_wait(&tmem_full[acc], acc_parity);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N);
      uint64_t rgb2[2][3] = {{0ull, 0ull, 0ull}, {0ull, 0ull, 0ull}};        // (even, odd channel) partial ToRGB sums of the two pixels
      uint8_t* stage_tile = smem_stage_out + (warp - 4) * kVStageOutBytes;   // [parity][32 pixels][16 channels] bf16
      uint4* stage_row = reinterpret_cast<uint4*>(stage_tile + lane * 32);
      const int stage_swp = (lane >> 2) & 1;                                 // bank-conflict-free order of the two 16-byte halves
      const bool want_out = e.out != nullptr && e.s_next != nullptr;
      const bool want_y = ok && e.y_out != nullptr;
      const int64_t pix = ((int64_t)b * p.H + Y) * p.W + X;
#pragma unroll 1
      for (int cs = 0; cs < CO; cs += 16) {     // 16 channels of BOTH pixels of the pair: every per-channel vector is fetched once
        uint32_t v0[16], v1[16];
        tmem_ld16(taddr + cs, v0);
        tmem_ld16(taddr + CO + cs, v1);
        tmem_ld_wait();
        uint32_t o0[8], o1[8], y0v[8], y1v[8];
        if (want_y) {
          epilogue_pair16<true, true>(v0, v1, s_d + cs, s_b + cs, s_n + cs, s_w + cs, s_w + CO + cs, s_w + 2 * CO + cs, nzp[0], nzp[1],
                                      rgb2[0], rgb2[1], o0, o1, y0v, y1v);
          uint4* yd = reinterpret_cast<uint4*>((__nv_bfloat16*)e.y_out + pix * CO + cs);
          yd[0] = make_uint4(y0v[0], y0v[1], y0v[2], y0v[3]);
          yd[1] = make_uint4(y0v[4], y0v[5], y0v[6], y0v[7]);
          yd = reinterpret_cast<uint4*>((__nv_bfloat16*)e.y_out + (pix + p.W) * CO + cs);
          yd[0] = make_uint4(y1v[0], y1v[1], y1v[2], y1v[3]);
          yd[1] = make_uint4(y1v[4], y1v[5], y1v[6], y1v[7]);
        } else {
          epilogue_pair16<true, false>(v0, v1, s_d + cs, s_b + cs, s_n + cs, s_w + cs, s_w + CO + cs, s_w + 2 * CO + cs, nzp[0], nzp[1],
                                       rgb2[0], rgb2[1], o0, o1, y0v, y1v);
        }
        if (want_out && p.tma_store) {
          if (lane == 0) tma_store_wait_read();       // the previous stores have finished reading this warp's staging tiles
          __syncwarp();
          const uint4 a_lo = make_uint4(o0[0], o0[1], o0[2], o0[3]), a_hi = make_uint4(o0[4], o0[5], o0[6], o0[7]);
          const uint4 b_lo = make_uint4(o1[0], o1[1], o1[2], o1[3]), b_hi = make_uint4(o1[4], o1[5], o1[6], o1[7]);
          stage_row[stage_swp] = stage_swp ? a_hi : a_lo;
          stage_row[stage_swp ^ 1] = stage_swp ? a_lo : a_hi;
          stage_row[64 + stage_swp] = stage_swp ? b_hi : b_lo;
          stage_row[64 + (stage_swp ^ 1)] = stage_swp ? b_lo : b_hi;
          // the warp's 4 pair-rows x 8 columns of one parity = every other image row: one strided TMA tensor store per parity
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d_nocommit(&tmap_o, stage_tile, cs, x0, y0 + 8 * q, b);
            tma_store_4d(&tmap_o, stage_tile + 1024, cs, x0, y0 + 8 * q + 1, b);
          }
        } else if (ok && want_out) {
          uint4* od = reinterpret_cast<uint4*>((__nv_bfloat16*)e.out + pix * CO + cs);
          od[0] = make_uint4(o0[0], o0[1], o0[2], o0[3]);
          od[1] = make_uint4(o0[4], o0[5], o0[6], o0[7]);
          od = reinterpret_cast<uint4*>((__nv_bfloat16*)e.out + (pix + p.W) * CO + cs);
          od[0] = make_uint4(o1[0], o1[1], o1[2], o1[3]);
          od[1] = make_uint4(o1[4], o1[5], o1[6], o1[7]);
        }
      }
      float rgb[2][3];
#pragma unroll
      for (int par = 0; par < 2; ++par)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float lo, hi;
          upk2(rgb2[par][c], lo, hi);
          rgb[par][c] = lo + hi;
        }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);

      // ---- ToRGB tail: bias + 2x FIR up-sampling of the skip image (ToRGB.forward, networks.py:349-358) from the patch.
      // Output rows Y = 2m, 2m+1 read low-res rows m-1 (f0), m (f2) and m (f1), m+1 (f3); the column parity picks the taps.
      float up[2][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) up[0][c] = up[1][c] = e.fused_skip ? __ldg(e.rgb_bias + c) : 0.f;
      if (p.has_skip) {
        const int px = lx & 1;
        const float wa = px ? e.fir[1] : e.fir[0], wb = px ? e.fir[3] : e.fir[2];
        const float* sk = skip_smem + buf * kVSkipStride + lj * kVSkipW + (lx >> 1) + 3 + px;   // patch origin = (m0 - 1, n0 - 4)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float h[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float* rowp = sk + (c * kVSkipH + i) * kVSkipW;
            h[i] = fmaf(wa, rowp[0], wb * rowp[1]);
          }
          up[0][c] += fmaf(e.fir[0], h[0], e.fir[2] * h[1]);
          up[1][c] += fmaf(e.fir[1], h[1], e.fir[3] * h[2]);
        }
      }
      if (e.wr != nullptr && ok) {
        float* dst = e.fused_skip ? e.skip_out : e.rgb_part;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float* pl = dst + ((int64_t)b * 3 + c) * plane + (int64_t)Y * p.W + X;
          pl[0] = rgb[0][c] + up[0][c];
          pl[p.W] = rgb[1][c] + up[1][c];
        }
      }
      // Hand the patch buffer back only after the stores that consume the loaded values have been issued: an
      // mbarrier.arrive does not wait for shared-memory loads that are still queued.
      if (p.has_skip || p.has_noise) mbar_arrive(&patch_empty[buf]);
    }
  }

  if (p.tma_store && warp >= 4 && lane == 0) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dst [kw][2 - kh][co][ci] bf16 (tile t = kw*3 + (2 - kh)); src [64 co][64 ci][3][3] fp32
__global__ void pack_vpair_weight_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, float scale) {
  const int total = 9 * 64 * 64;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ci = idx % 64, co = (idx / 64) % 64, t = idx / 4096;
    const int kw = t / 3, kh = 2 - t % 3;
    dst[idx] = __float2bfloat16_rn(src[(((int64_t)co * 64 + ci) * 3 + kh) * 3 + kw] * scale);
  }
}

}  // namespace

bool conv_tc_vpair_supported(const ConvGeom& g, const EpiParams& e) {
  if (!g_switches.vpair || !tmap_available()) return false;
  if (g.nphase != 1 || g.in_scale != 1 || g.up_cout != 0 || g.in_pair_packed || g.weight_taps != 9) return false;
  if (g.Cin != 64 || g.Cout != 64 || e.mode != 0 || e.wr == nullptr || !e.fused_skip) return false;
  if (g.H < 32 || g.W < 8 || (g.H & 1) || g.OH != g.H || g.OW != g.W) return false;
  return true;
}

int launch_pack_vpair_weight(__nv_bfloat16* dst, const float* src, float scale, cudaStream_t st) {
  pack_vpair_weight_kernel<<<ceil_div(9 * 64 * 64, 256), 256, 0, st>>>(dst, src, scale);
  return check_launch("pack_vpair_weight");
}

// w: the [9][64][64] bf16 tiles written by launch_pack_vpair_weight
int launch_conv_tc_vpair(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st) {
  VPairParams p{};
  p.B = g.B; p.H = g.H; p.W = g.W; p.e = e;
  p.idesc128 = make_idesc_bf16(128, 128, 0);
  p.idesc64 = make_idesc_bf16(128, 64, 0);
  CUtensorMap ta, tw;
  {
    const uint64_t dims[4] = {64, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
    const uint64_t str[4] = {2, 128, (uint64_t)g.W * 128, (uint64_t)g.H * g.W * 128};
    const uint32_t box[4] = {64, kVW, kVH, 1};
    L2I_TRY(make_tmap(&ta, in, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    const uint64_t dims[3] = {64, 64, 9};
    const uint64_t str[3] = {2, 128, 64 * 128};
    const uint32_t box[3] = {64, 64, 1};
    L2I_TRY(make_tmap(&tw, w, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  CUtensorMap ts = ta, tn = ta;   // placeholders when there is no skip image / noise
  p.has_skip = (e.fused_skip && e.skip_in != nullptr) ? 1 : 0;
  if (p.has_skip) {
    const uint64_t h2 = (uint64_t)g.H / 2, w2 = (uint64_t)g.W / 2;
    const uint64_t dims[4] = {w2, h2, 3, (uint64_t)g.B};
    const uint64_t str[4] = {4, w2 * 4, h2 * w2 * 4, 3 * h2 * w2 * 4};
    const uint32_t box[4] = {kVSkipW, kVSkipH, 3, 1};
    if ((uintptr_t)e.skip_in % 16 != 0 || (w2 * 4) % 16 != 0) { set_error("conv_tc_vpair: skip image must be 16-byte aligned"); return L2I_ERR_INVALID_ARG; }
    L2I_TRY(make_tmap_f32(&ts, e.skip_in, 4, dims, str, box));
  }
  p.has_noise = (e.noise != nullptr && e.noise_w != nullptr) ? 1 : 0;
  p.noise_per_sample = e.noise_bs != 0 ? 1 : 0;
  if (p.has_noise) {
    const uint64_t nb = p.noise_per_sample ? (uint64_t)g.B : 1;
    const uint64_t bs = p.noise_per_sample ? (uint64_t)e.noise_bs * 4 : (uint64_t)g.H * g.W * 4;
    const uint64_t dims[3] = {(uint64_t)g.W, (uint64_t)g.H, nb};
    const uint64_t str[3] = {4, (uint64_t)g.W * 4, bs};
    const uint32_t box[3] = {8, 32, 1};
    if ((uintptr_t)e.noise % 16 != 0 || bs % 16 != 0) { set_error("conv_tc_vpair: noise must be 16-byte aligned"); return L2I_ERR_INVALID_ARG; }
    L2I_TRY(make_tmap_f32(&tn, e.noise, 3, dims, str, box));
  }
  CUtensorMap to = ta;
  p.tma_store = 0;
  if (e.out != nullptr && e.s_next != nullptr && e.y_out == nullptr && (uintptr_t)e.out % 16 == 0) {
    // a warp's 32 pixels of one parity: 8 columns x 4 rows, rows two apart; 16 of the 64 channels per store
    const uint64_t dims[4] = {64, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
    const uint64_t str[4] = {2, 128, (uint64_t)g.W * 128, (uint64_t)g.H * g.W * 128};
    const uint32_t box[4] = {16, 8, 8, 1};
    const uint32_t estr[4] = {1, 1, 2, 1};
    L2I_TRY(make_tmap_strided(&to, e.out, 4, dims, str, box, estr, CU_TENSOR_MAP_SWIZZLE_NONE));
    p.tma_store = 1;
  }
  p.tiles_x = ceil_div(g.W, kVTileW); p.tiles_y = ceil_div(g.H, kVTileH);
  const int64_t total = (int64_t)p.tiles_x * p.tiles_y * g.B;
  if (total <= 0 || total > 0x7fffffff) { set_error("conv_tc_vpair: bad tile count"); return L2I_ERR_INVALID_ARG; }
  p.total_tiles = (int)total;
  static bool attr_set = false;
  if (!attr_set) {
    L2I_CUDA_TRY(cudaFuncSetAttribute(conv_tc_vpair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kVSmem));
    attr_set = true;
  }
  const int grid = std::min(p.total_tiles, kNumSMs);
  conv_tc_vpair_kernel<<<grid, kVThreads, kVSmem, st>>>(ta, tw, to, ts, tn, p);
  return check_launch("conv_tc_vpair");
}

}  // namespace l2i
