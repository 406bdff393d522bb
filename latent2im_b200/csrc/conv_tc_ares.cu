// A-resident / B-streamed tcgen05 implicit-GEMM conv for the Cin = 128 layers (256 x 256 px, plain 128 -> 128 with the
// act + ToRGB epilogue, and the composite 128 -> 4 x 64 up-conv).
//
// The general kernel (conv_tc.cu) reloads the activation tile for every tap: 9 x 16 KB of A plus 9 x 16/32 KB of B per
// 128-pixel tile and 64-channel chunk.  With N <= 256 and K = 128 per tap its MMAs are short, so that traffic (and the
// depth of the TMA ring needed to cover its latency) bounds the layer (ncu: tensor pipe 47 % / 65 % active).  Here the
// (8+2) x (16+2) pixel halo tile of the activation (two 64-channel planes, 46 KB) is TMA-loaded ONCE per output tile and
// the nine taps are UMMA descriptors at tap-shifted start addresses (as in conv_tc_halo.cu); only the weight tiles
// stream through a shared-memory ring, fed by their own producer warp.  L2 -> SMEM bytes per tile: 46 KB + 9 x 2 x B-tile
// instead of 9 x 2 x (16 KB + B-tile).
#include <type_traits>

#include "tc_epilogue.cuh"

namespace l2i {

using namespace tc;

namespace {

constexpr int kRW = 10, kRH = 18;                          // halo tile (pixels): 8 + 2 wide, 16 + 2 tall
constexpr int kRPlaneBytes = kRW * kRH * 128;              // 23040: one 64-channel plane
constexpr int kRPlaneStride = (kRPlaneBytes + 1023) & ~1023;
constexpr int kRAStages = 2;
constexpr int kRTileW = 8, kRTileH = 16;

struct AresParams {
  int B, H, W;               // input = output grid of the GEMM rows
  int Cout;                  // real output channels (N for the plain layer, N / 4 for the composite up-conv)
  int out_H, out_W;
  int tiles_x, tiles_y, total_tiles;
  uint32_t idesc;
  EpiParams e;
};

__device__ __forceinline__ uint64_t ares_desc(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void ares_group_sync(int group) {
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

// N = GEMM N, KC = 64-channel chunks of Cin, BP = chunks per weight-ring stage (1 or KC), WST = weight ring stages,
// COMP = composite up-conv (N = 4 phases x N/4).  The single MMA-issuing thread pays ~100 clocks of mbarrier latency per
// ring stage; with N = 128 a 64-channel stage is only 256 tensor clocks, so those layers take whole taps (BP = KC) per stage.
template <int N, int KC, int BP, int WST, int GROUPS, int EPI, bool COMP>
__global__ void __launch_bounds__(128 + GROUPS * 128, 1)
conv_tc_ares_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ AresParams p) {
  constexpr int kAStageBytes = KC * kRPlaneStride;
  constexpr int kBPlaneBytes = N * 128;
  constexpr int kBStageBytes = BP * kBPlaneBytes;
  static_assert(KC % BP == 0, "ring stage = BP chunks of one tap");
  constexpr int CO = COMP ? N / 4 : N;
  constexpr int kEpiFloats = 6 * CO;
  static_assert(GROUPS * N <= 512, "TMEM budget");
  static_assert(!COMP || EPI == EPI_ACT, "composite variant uses the act-only epilogue");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_b = smem + kRAStages * kAStageBytes;
  __shared__ __align__(16) float epi_smem[GROUPS * kEpiFloats];
  __shared__ __align__(8) uint64_t a_full[kRAStages];
  __shared__ __align__(8) uint64_t a_empty[kRAStages];
  __shared__ __align__(8) uint64_t w_full[WST];
  __shared__ __align__(8) uint64_t w_empty[WST];
  __shared__ __align__(8) uint64_t tmem_full[GROUPS];
  __shared__ __align__(8) uint64_t tmem_empty[GROUPS];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < WST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int a = 0; a < GROUPS; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
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
    x0 = tx * kRTileW; y0 = ty * kRTileH;
  };

  if (warp == 0) {
    // ===================== A producer: the halo tile (KC planes) once per output tile =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase_bit = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int x0, y0, b;
        decode(tile, x0, y0, b);
        mbar_wait(&a_empty[stage], phase_bit ^ 1);
        mbar_expect_tx(&a_full[stage], KC * kRPlaneBytes);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc)
          tma_load_4d(smem + stage * kAStageBytes + kc * kRPlaneStride, &tmap_a, &a_full[stage], kc * 64, x0 - 1, y0 - 1, b);
        if (++stage == kRAStages) { stage = 0; phase_bit ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================== B producer: 9 taps x KC weight tiles per output tile through the ring ===========
    if (lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        for (int t = 0; t < 9; ++t) {
#pragma unroll
          for (int kc = 0; kc < KC; kc += BP) {
            mbar_wait(&w_empty[ws], wphase ^ 1);
            mbar_expect_tx(&w_full[ws], kBStageBytes);
#pragma unroll
            for (int j = 0; j < BP; ++j)
              tma_load_3d(smem_b + ws * kBStageBytes + j * kBPlaneBytes, &tmap_w, &w_full[ws], (kc + j) * 64, 0, t);
            if (++ws == WST) { ws = 0; wphase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform control flow, one elected lane issues (tc_ptx.cuh: elect_one) =============
    {
      int stage = 0, ws = 0, grp = 0;
      uint32_t phase_bit = 0, wphase = 0, grp_phase = 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_a0 = smem_u32(smem), smem_b0 = smem_u32(smem_b);
      constexpr uint64_t kHiA = kmajor_desc_hi(kRW * 128, 2), kHiB = kmajor_desc_hi(1024, 2);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[grp], grp_phase ^ 1);
        mbar_wait(&a_full[stage], phase_bit);
        tc_fence_after();
        const uint32_t a_base = smem_a0 + (uint32_t)(stage * kAStageBytes);
        const uint32_t tmem_d = tmem_u + (uint32_t)(grp * N);
#pragma unroll 1
        for (int t = 0; t < 9; ++t) {
          const uint32_t shift = (uint32_t)(((t / 3) * kRW + (t % 3)) * 128);   // tap (dy, dx) = (t/3 - 1, t%3 - 1)
#pragma unroll
          for (int kc = 0; kc < KC; kc += BP) {
            mbar_wait(&w_full[ws], wphase);
            tc_fence_after();
            const uint64_t a_desc = kmajor_desc_at(kHiA, a_base + (uint32_t)(kc * kRPlaneStride) + shift);
            const uint64_t b_desc = kmajor_desc_at(kHiB, smem_b0 + (uint32_t)(ws * kBStageBytes));
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < BP; ++j) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(tmem_d, a_desc + (uint64_t)((j * kRPlaneStride + k * 32) >> 4), b_desc + (uint64_t)((j * kBPlaneBytes + k * 32) >> 4),
                            p.idesc, (t | kc | j | k) != 0 ? 1u : 0u);
              }
              umma_commit(&w_empty[ws]);
            }
            if (++ws == WST) { ws = 0; wphase ^= 1; }
          }
        }
        if (elect_one()) {
          umma_commit(&a_empty[stage]);
          umma_commit(&tmem_full[grp]);
        }
        if (++stage == kRAStages) { stage = 0; phase_bit ^= 1; }
        if (++grp == GROUPS) { grp = 0; grp_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const EpiParams& e = p.e;
    const int group = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gtid = threadIdx.x - (128 + group * 128);
    float* sp = epi_smem + group * kEpiFloats;
    float* s_d = sp;
    float* s_b = sp + CO;
    float* s_n = sp + 2 * CO;
    float* s_w = sp + 3 * CO;
    constexpr float kSqrt2 = 1.4142135623730951f;
    const float nw = (e.noise != nullptr && e.noise_w != nullptr) ? __ldg(e.noise_w) * kSqrt2 : 0.f;
    const int64_t plane = (int64_t)p.out_H * p.out_W;
    const int lx = row & 7, ly = row >> 3;
    uint32_t grp_phase = 0;
    int staged_b = -1;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      if (it % GROUPS != group) continue;
      int x0, y0, b;
      decode(tile, x0, y0, b);
      const int ox = x0 + lx, oy = y0 + ly;
      const bool ok = ox < p.W && oy < p.H;

      if (b != staged_b) {
        ares_group_sync(group);
        for (int j = gtid; j < CO; j += 128) {
          s_d[j] = (e.demod != nullptr ? __ldg(e.demod + (int64_t)b * e.demod_bs + j) : 1.f) * kSqrt2;
          s_b[j] = __ldg(e.bias + j) * kSqrt2;
          s_n[j] = e.s_next ? __ldg(e.s_next + (int64_t)b * e.s_next_bs + j) : 1.f;
          if (EPI == EPI_ACT_RGB) {
#pragma unroll
            for (int c = 0; c < 3; ++c) s_w[c * CO + j] = e.wr ? __ldg(e.wr + (int64_t)b * e.wr_bs + c * p.Cout + j) : 0.f;
          }
        }
        ares_group_sync(group);
        staged_b = b;
      }

      // ---- global loads of the tile before waiting for the accumulator ----
      float nzq[4] = {0.f, 0.f, 0.f, 0.f};
      float up[3] = {0.f, 0.f, 0.f};
      if (ok && e.noise != nullptr) {
        if (COMP) {
          const float* np = e.noise + (int64_t)b * e.noise_bs + (int64_t)(2 * oy) * p.out_W + 2 * ox;
          const float2 n01 = __ldg(reinterpret_cast<const float2*>(np));
          const float2 n23 = __ldg(reinterpret_cast<const float2*>(np + p.out_W));
          nzq[0] = nw * n01.x; nzq[1] = nw * n01.y; nzq[2] = nw * n23.x; nzq[3] = nw * n23.y;
        } else {
          nzq[0] = nw * __ldg(e.noise + (int64_t)b * e.noise_bs + (int64_t)oy * p.out_W + ox);
        }
      }
      if (EPI == EPI_ACT_RGB && ok && e.fused_skip) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          up[c] = __ldg(e.rgb_bias + c);
          if (e.skip_in != nullptr)
            up[c] += upsample2x_at(e.skip_in + ((int64_t)b * 3 + c) * (plane / 4), p.out_H / 2, p.out_W / 2, oy, ox, e.fir);
        }
      }

      mbar_wait(&tmem_full[group], grp_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(group * N);
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        const int phc = COMP ? c0 / CO : 0;            // composite: the chunk's output phase (py*2 + px)
        const int cs = COMP ? c0 % CO : c0;            // first channel of the chunk
        const float nzc = phc == 0 ? nzq[0] : (phc == 1 ? nzq[1] : (phc == 2 ? nzq[2] : nzq[3]));
        __nv_bfloat16* outc = nullptr;
        __nv_bfloat16* yc = nullptr;
        if (ok) {
          const int Y = COMP ? 2 * oy + (phc >> 1) : oy, X = COMP ? 2 * ox + (phc & 1) : ox;
          const int64_t pix = ((int64_t)b * p.out_H + Y) * p.out_W + X;
          if (e.out != nullptr && e.s_next != nullptr) outc = (__nv_bfloat16*)e.out + pix * p.Cout + cs;
          if (e.y_out != nullptr) yc = (__nv_bfloat16*)e.y_out + pix * p.Cout + cs;
        }
        tmem_ld_wait();
        epilogue_chunk32<EPI>(v, s_d + cs, s_b + cs, s_n + cs, s_w + cs, s_w + CO + cs, s_w + 2 * CO + cs, nzc, false, rgb0, rgb1,
                              rgb2, outc, yc);
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[group]);
      grp_phase ^= 1;

      if (EPI == EPI_ACT_RGB && e.wr != nullptr && ok) {
        const float r3[3] = {rgb0, rgb1, rgb2};
        float* dst = e.fused_skip ? e.skip_out : e.rgb_part;
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[((int64_t)b * 3 + c) * plane + (int64_t)oy * p.out_W + ox] = r3[c] + up[c];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Plain 128 -> 128 layer (act + ToRGB epilogue), round 2: TILE PAIRS.  The epilogue of the kernel above owns one pixel x 128
// channels per lane and fetches six per-channel vectors for it with warp-wide LDS.128 (four shared-memory wavefronts each, same
// address or not), then writes its 256 bytes with 16-byte stores at pixel pitch (one wavefront per lane): 3072 + 2048 wavefronts per
// tile next to the 4608 of the MMA operands and the 2664 of the TMA fills - the shared-memory / L1 data pipe bounds the layer
// (ncu: tensor pipe 57 %).  Here a CTA walks a CONTIGUOUS tile range, four accumulators live in TMEM, an epilogue warpgroup
// takes two consecutive tiles at once so that every vector fetched serves two pixels (epilogue_pair16), and the activation leaves
// through per-warp staging tiles and TMA tensor stores.
constexpr int kR2Accs = 4;
constexpr int kR2StageOut = 2048;      // per epilogue warp: [tile of the pair][32 pixels][16 channels] bf16

template <int WST>
__global__ void __launch_bounds__(128 + 2 * 128, 1)
conv_tc_ares_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ AresParams p) {
  constexpr int N = 128, KC = 2, CO = 128;
  constexpr int kAStageBytes = KC * kRPlaneStride;
  constexpr int kBPlaneBytes = N * 128;
  constexpr int kBStageBytes = KC * kBPlaneBytes;      // one tap (both 64-channel planes) per ring stage
  constexpr int kEpiFloats = 6 * CO;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_b = smem + kRAStages * kAStageBytes;
  uint8_t* smem_out = smem_b + WST * kBStageBytes;
  __shared__ __align__(16) float epi_smem[2 * kEpiFloats];
  __shared__ __align__(8) uint64_t a_full[kRAStages];
  __shared__ __align__(8) uint64_t a_empty[kRAStages];
  __shared__ __align__(8) uint64_t w_full[WST];
  __shared__ __align__(8) uint64_t w_empty[WST];
  __shared__ __align__(8) uint64_t tmem_full[kR2Accs];
  __shared__ __align__(8) uint64_t tmem_empty[kR2Accs];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < WST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int a = 0; a < kR2Accs; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // contiguous tile range of this CTA: neighbouring tiles share halo rows in L2 and, almost always, the sample
  const int t_begin = (int)((int64_t)p.total_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((int64_t)p.total_tiles * (blockIdx.x + 1) / gridDim.x);
  auto decode = [&](int tile, int& x0, int& y0, int& b) {
    const int tx = tile % p.tiles_x;
    const int r = tile / p.tiles_x;
    const int ty = r % p.tiles_y;
    b = r / p.tiles_y;
    x0 = tx * kRTileW; y0 = ty * kRTileH;
  };

  if (warp == 0) {
    // ===================== A producer: the halo tile (KC planes) once per output tile =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase_bit = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        int x0, y0, b;
        decode(tile, x0, y0, b);
        mbar_wait(&a_empty[stage], phase_bit ^ 1);
        mbar_expect_tx(&a_full[stage], KC * kRPlaneBytes);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc)
          tma_load_4d(smem + stage * kAStageBytes + kc * kRPlaneStride, &tmap_a, &a_full[stage], kc * 64, x0 - 1, y0 - 1, b);
        if (++stage == kRAStages) { stage = 0; phase_bit ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================== B producer: 9 taps per output tile through the ring =====================
    if (lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        for (int t = 0; t < 9; ++t) {
          mbar_wait(&w_empty[ws], wphase ^ 1);
          mbar_expect_tx(&w_full[ws], kBStageBytes);
#pragma unroll
          for (int j = 0; j < KC; ++j) tma_load_3d(smem_b + ws * kBStageBytes + j * kBPlaneBytes, &tmap_w, &w_full[ws], j * 64, 0, t);
          if (++ws == WST) { ws = 0; wphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform control flow, one elected lane issues (tc_ptx.cuh: elect_one) =============
    int stage = 0, ws = 0;
    uint32_t phase_bit = 0, wphase = 0;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_a0 = smem_u32(smem), smem_b0 = smem_u32(smem_b);
    constexpr uint64_t kHiA = kmajor_desc_hi(kRW * 128, 2), kHiB = kmajor_desc_hi(1024, 2);
    int it = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int acc = it & (kR2Accs - 1);
      mbar_wait(&tmem_empty[acc], (((uint32_t)it / kR2Accs) & 1u) ^ 1u);
      mbar_wait(&a_full[stage], phase_bit);
      tc_fence_after();
      const uint32_t a_base = smem_a0 + (uint32_t)(stage * kAStageBytes);
      const uint32_t tmem_d = tmem_u + (uint32_t)(acc * N);
#pragma unroll 1
      for (int t = 0; t < 9; ++t) {
        const uint32_t shift = (uint32_t)(((t / 3) * kRW + (t % 3)) * 128);   // tap (dy, dx) = (t/3 - 1, t%3 - 1)
        mbar_wait(&w_full[ws], wphase);
        tc_fence_after();
        const uint64_t a_desc = kmajor_desc_at(kHiA, a_base + shift);
        const uint64_t b_desc = kmajor_desc_at(kHiB, smem_b0 + (uint32_t)(ws * kBStageBytes));
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < KC; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_d, a_desc + (uint64_t)((j * kRPlaneStride + k * 32) >> 4), b_desc + (uint64_t)((j * kBPlaneBytes + k * 32) >> 4),
                        p.idesc, (t | j | k) != 0 ? 1u : 0u);
          umma_commit(&w_empty[ws]);
        }
        if (++ws == WST) { ws = 0; wphase ^= 1; }
      }
      if (elect_one()) {
        umma_commit(&a_empty[stage]);
        umma_commit(&tmem_full[acc]);
      }
      if (++stage == kRAStages) { stage = 0; phase_bit ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: warpgroup g takes the tile pairs (4k + 2g, 4k + 2g + 1) = accumulators 2g, 2g + 1 ==========
    const EpiParams& e = p.e;
    const int group = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gtid = threadIdx.x - (128 + group * 128);
    float* sp = epi_smem + group * kEpiFloats;
    float* s_d = sp;
    float* s_b = sp + CO;
    float* s_n = sp + 2 * CO;
    float* s_w = sp + 3 * CO;
    constexpr float kSqrt2 = 1.4142135623730951f;
    const float nw = (e.noise != nullptr && e.noise_w != nullptr) ? __ldg(e.noise_w) * kSqrt2 : 0.f;
    const int64_t plane = (int64_t)p.out_H * p.out_W;
    const int lx = row & 7, ly = row >> 3;
    uint8_t* stage_tile = smem_out + (warp - 4) * kR2StageOut;
    uint4* stage_row = reinterpret_cast<uint4*>(stage_tile + lane * 32);
    const int stage_swp = (lane >> 2) & 1;                 // bank-conflict-free order of the two 16-byte halves of a 32-byte row
    const bool want_out = e.out != nullptr && e.s_next != nullptr;
    int staged_b = -1;
    const int ntiles = t_end - t_begin;
    for (int it0 = 2 * group; it0 < ntiles; it0 += 4) {
      const uint32_t acc_parity = ((uint32_t)it0 / kR2Accs) & 1u;
      const bool have2 = it0 + 1 < ntiles;
      int x0[2], y0[2], bb[2];
      decode(t_begin + it0, x0[0], y0[0], bb[0]);
      if (have2) decode(t_begin + it0 + 1, x0[1], y0[1], bb[1]);
      else { x0[1] = x0[0]; y0[1] = y0[0]; bb[1] = bb[0]; }
      // ---- global loads of both tiles before waiting for the accumulators ----
      float nz[2] = {0.f, 0.f}, up[2][3];
      bool ok[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int ox = x0[s] + lx, oy = y0[s] + ly;
        ok[s] = ox < p.W && oy < p.H && (s == 0 || have2);
        if (ok[s] && e.noise != nullptr) nz[s] = nw * __ldg(e.noise + (int64_t)bb[s] * e.noise_bs + (int64_t)oy * p.out_W + ox);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          up[s][c] = 0.f;
          if (ok[s] && e.fused_skip) {
            up[s][c] = __ldg(e.rgb_bias + c);
            if (e.skip_in != nullptr)
              up[s][c] += upsample2x_at(e.skip_in + ((int64_t)bb[s] * 3 + c) * (plane / 4), p.out_H / 2, p.out_W / 2, oy, ox, e.fir);
          }
        }
      }
      mbar_wait(&tmem_full[2 * group], acc_parity);
      if (have2) mbar_wait(&tmem_full[2 * group + 1], acc_parity);
      tc_fence_after();
      // a pair that straddles two samples (once per CTA at most, the tile range is contiguous) is processed one tile at a time
      uint64_t rgb2[2][3] = {{0ull, 0ull, 0ull}, {0ull, 0ull, 0ull}};
      // MODE 0: both tiles at once; MODE 1 / 2: tile 0 / tile 1 alone (compile-time indices keep rgb2 & co. in registers)
      auto run_pass = [&](auto mode_c) {
        constexpr int MODE = decltype(mode_c)::value;
        constexpr int sa = MODE == 2 ? 1 : 0, sb = MODE == 0 ? 1 : sa;   // accumulators feeding "pixel a" / "pixel b"
        constexpr bool two = MODE == 0;
        uint64_t rgb_unused[3] = {0ull, 0ull, 0ull};
        const int b = bb[sa];
        if (b != staged_b) {
          ares_group_sync(group);
          for (int j = gtid; j < CO; j += 128) {
            s_d[j] = (e.demod != nullptr ? __ldg(e.demod + (int64_t)b * e.demod_bs + j) : 1.f) * kSqrt2;
            s_b[j] = __ldg(e.bias + j) * kSqrt2;
            s_n[j] = e.s_next ? __ldg(e.s_next + (int64_t)b * e.s_next_bs + j) : 1.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) s_w[c * CO + j] = e.wr ? __ldg(e.wr + (int64_t)b * e.wr_bs + c * p.Cout + j) : 0.f;
          }
          ares_group_sync(group);
          staged_b = b;
        }
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((2 * group + sa) * N);
        const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((2 * group + sb) * N);
        const int64_t pixa = ((int64_t)bb[sa] * p.out_H + y0[sa] + ly) * p.out_W + x0[sa] + lx;
        const int64_t pixb = ((int64_t)bb[sb] * p.out_H + y0[sb] + ly) * p.out_W + x0[sb] + lx;
#pragma unroll 1
        for (int cs = 0; cs < N; cs += 16) {
          uint32_t va[16], vb[16];
          tmem_ld16(ta + cs, va);
          tmem_ld16(tb + cs, vb);
          tmem_ld_wait();
          uint32_t oa[8], ob[8], ya[8], yb[8];
          if (e.y_out != nullptr)
            epilogue_pair16<true, true>(va, vb, s_d + cs, s_b + cs, s_n + cs, s_w + cs, s_w + CO + cs, s_w + 2 * CO + cs, nz[sa], nz[sb],
                                        rgb2[sa], two ? rgb2[sb] : rgb_unused, oa, ob, ya, yb);
          else
            epilogue_pair16<true, false>(va, vb, s_d + cs, s_b + cs, s_n + cs, s_w + cs, s_w + CO + cs, s_w + 2 * CO + cs, nz[sa], nz[sb],
                                         rgb2[sa], two ? rgb2[sb] : rgb_unused, oa, ob, ya, yb);
          if (e.y_out != nullptr) {   // training: the unscaled activation is kept for the backward pass
            if (ok[sa]) {
              uint4* yd = reinterpret_cast<uint4*>((__nv_bfloat16*)e.y_out + pixa * CO + cs);
              yd[0] = make_uint4(ya[0], ya[1], ya[2], ya[3]);
              yd[1] = make_uint4(ya[4], ya[5], ya[6], ya[7]);
            }
            if (two && ok[sb]) {
              uint4* yd = reinterpret_cast<uint4*>((__nv_bfloat16*)e.y_out + pixb * CO + cs);
              yd[0] = make_uint4(yb[0], yb[1], yb[2], yb[3]);
              yd[1] = make_uint4(yb[4], yb[5], yb[6], yb[7]);
            }
          }
          if (want_out) {
            if (lane == 0) tma_store_wait_read();       // the previous stores have finished reading this warp's staging tiles
            __syncwarp();
            const uint4 a_lo = make_uint4(oa[0], oa[1], oa[2], oa[3]), a_hi = make_uint4(oa[4], oa[5], oa[6], oa[7]);
            stage_row[stage_swp] = stage_swp ? a_hi : a_lo;
            stage_row[stage_swp ^ 1] = stage_swp ? a_lo : a_hi;
            if (two) {
              const uint4 b_lo = make_uint4(ob[0], ob[1], ob[2], ob[3]), b_hi = make_uint4(ob[4], ob[5], ob[6], ob[7]);
              stage_row[64 + stage_swp] = stage_swp ? b_hi : b_lo;
              stage_row[64 + (stage_swp ^ 1)] = stage_swp ? b_lo : b_hi;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {      // a warp's 8 columns x 4 rows of a tile; the tensor map clips what lies outside the image
              if (two) tma_store_4d_nocommit(&tmap_o, stage_tile + 1024, cs, x0[sb], y0[sb] + 4 * q, bb[sb]);
              tma_store_4d(&tmap_o, stage_tile, cs, x0[sa], y0[sa] + 4 * q, bb[sa]);
            }
          }
        }
      };
      if (!have2) {
        run_pass(std::integral_constant<int, 1>{});
      } else if (bb[1] != bb[0]) {
        run_pass(std::integral_constant<int, 1>{});
        run_pass(std::integral_constant<int, 2>{});
      } else {
        run_pass(std::integral_constant<int, 0>{});
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[2 * group]);
      if (have2) mbar_arrive(&tmem_empty[2 * group + 1]);

      if (e.wr != nullptr) {
        float* dst = e.fused_skip ? e.skip_out : e.rgb_part;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (!ok[s]) continue;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float lo, hi;
            upk2(rgb2[s][c], lo, hi);
            dst[((int64_t)bb[s] * 3 + c) * plane + (int64_t)(y0[s] + ly) * p.out_W + x0[s] + lx] = (lo + hi) + up[s][c];
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int WST>
int launch_ares_pair(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const AresParams& p, cudaStream_t st) {
  constexpr int smem = kRAStages * 2 * kRPlaneStride + WST * 2 * 128 * 128 + 8 * kR2StageOut + 1024;
  static_assert(smem + 2 * 6 * 128 * 4 + 512 <= 227 * 1024, "shared memory budget (dynamic + static epilogue vectors)");
  auto kern = conv_tc_ares_pair_kernel<WST>;
  static bool attr_set = false;
  if (!attr_set) {
    L2I_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int grid = std::min(p.total_tiles, kNumSMs);
  kern<<<grid, 128 + 2 * 128, smem, st>>>(ta, tw, to, p);
  return check_launch("conv_tc_ares_pair");
}

template <int N, int KC, int BP, int WST, int GROUPS, int EPI, bool COMP>
int launch_ares_variant(const CUtensorMap& ta, const CUtensorMap& tw, const AresParams& p, cudaStream_t st) {
  constexpr int smem = kRAStages * KC * kRPlaneStride + WST * BP * N * 128 + 1024;
  static_assert(smem + GROUPS * 6 * (COMP ? N / 4 : N) * 4 + 512 <= 227 * 1024, "shared memory budget (dynamic + static epilogue vectors)");
  auto kern = conv_tc_ares_kernel<N, KC, BP, WST, GROUPS, EPI, COMP>;
  static bool attr_set = false;
  if (!attr_set) {
    L2I_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int grid = std::min(p.total_tiles, kNumSMs);
  kern<<<grid, 128 + GROUPS * 128, smem, st>>>(ta, tw, p);
  return check_launch("conv_tc_ares");
}

// ---------------------------------------------------------------------------------------------------------------------------
// Halo RING for the wide plain layers (Cin = 256 / 512, N = 256 output channels per tile; 512 -> 512 @ 64^2, 256 -> 256 @ 128^2 ...).
// The general kernel (conv_tc.cu) reloads the 16 KB activation tile for every (tap, 64-channel chunk): its TMA fills (A + B, each
// byte written once and read once) cost as many shared-memory wavefronts as the MMA operand reads, and the layer's time equals
// the wavefront count of that pipe (65 k per tile at 512 -> 512, measured = computed).  Here the loop order is chunk-outer /
// tap-inner: ONE (8+2) x (16+2) halo plane per 64-channel chunk serves all nine taps (UMMA descriptors at tap-shifted start
// addresses, as above), planes cycle through a three-slot ring with their own full / empty barriers so the next tile's planes
// arrive while this tile's last chunks are multiplied; only the [256][64] weight tiles stream per (chunk, tap).
// A fills per tile: 9 x 16 KB per chunk -> 23 KB per chunk.
constexpr int kHASlots = 3;
constexpr int kHN = 256;

template <int WST>
__global__ void __launch_bounds__(128 + 2 * 128, 1)
conv_tc_hring_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                     const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ AresParams p, const int KC, const int tiles_n,
                     const int tma_store) {
  constexpr int N = kHN, GROUPS = 2, CO = kHN;
  constexpr int kBStageBytes = N * 128;                 // one (chunk, tap) weight tile
  constexpr int kEpiFloats = 6 * CO;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_b = smem + kHASlots * kRPlaneStride;
  uint8_t* smem_out = smem_b + WST * kBStageBytes;     // per epilogue warp: 32 pixels x 32 channels, SWIZZLE_64B rows (TMA store source)
  __shared__ __align__(16) float epi_smem[GROUPS * kEpiFloats];
  __shared__ __align__(8) uint64_t a_full[kHASlots];
  __shared__ __align__(8) uint64_t a_empty[kHASlots];
  __shared__ __align__(8) uint64_t w_full[WST];
  __shared__ __align__(8) uint64_t w_empty[WST];
  __shared__ __align__(8) uint64_t tmem_full[GROUPS];
  __shared__ __align__(8) uint64_t tmem_empty[GROUPS];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kHASlots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < WST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int a = 0; a < GROUPS; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // contiguous tile range per CTA, channel tile fastest: the N tiles of one pixel tile follow each other (halo planes hot in L2)
  const int t_begin = (int)((int64_t)p.total_tiles * blockIdx.x / gridDim.x);
  const int t_end = (int)((int64_t)p.total_tiles * (blockIdx.x + 1) / gridDim.x);
  auto decode = [&](int tile, int& nt, int& x0, int& y0, int& b) {
    nt = tile % tiles_n;
    int r = tile / tiles_n;
    const int tx = r % p.tiles_x;
    r /= p.tiles_x;
    const int ty = r % p.tiles_y;
    b = r / p.tiles_y;
    x0 = tx * kRTileW; y0 = ty * kRTileH;
  };

  if (warp == 0) {
    // ===================== A producer: one halo plane per (tile, 64-channel chunk) =====================
    if (lane == 0) {
      uint32_t acnt = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        int nt, x0, y0, b;
        decode(tile, nt, x0, y0, b);
        for (int kc = 0; kc < KC; ++kc, ++acnt) {
          const int slot = acnt % kHASlots;
          mbar_wait(&a_empty[slot], ((acnt / kHASlots) & 1) ^ 1);
          mbar_expect_tx(&a_full[slot], kRPlaneBytes);
          tma_load_4d(smem + slot * kRPlaneStride, &tmap_a, &a_full[slot], kc * 64, x0 - 1, y0 - 1, b);
        }
      }
    }
  } else if (warp == 3) {
    // ===================== B producer: the [256][64] weight tile of every (chunk, tap), same order as the MMA issuer ==========
    if (lane == 0) {
      uint32_t wcnt = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        int nt, x0, y0, b;
        decode(tile, nt, x0, y0, b);
        for (int kc = 0; kc < KC; ++kc)
          for (int t = 0; t < 9; ++t, ++wcnt) {
            const int ws = wcnt % WST;
            mbar_wait(&w_empty[ws], ((wcnt / WST) & 1) ^ 1);
            mbar_expect_tx(&w_full[ws], kBStageBytes);
            tma_load_3d(smem_b + ws * kBStageBytes, &tmap_w, &w_full[ws], kc * 64, nt * N, t);
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform control flow, one elected lane issues (tc_ptx.cuh: elect_one) =============
    uint32_t acnt = 0, wcnt = 0;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_a0 = smem_u32(smem), smem_b0 = smem_u32(smem_b);
    constexpr uint64_t kHiA = kmajor_desc_hi(kRW * 128, 2), kHiB = kmajor_desc_hi(1024, 2);
    int it = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int acc = it & 1;
      mbar_wait(&tmem_empty[acc], (((uint32_t)it >> 1) & 1u) ^ 1u);
      const uint32_t tmem_d = tmem_u + (uint32_t)(acc * N);
      for (int kc = 0; kc < KC; ++kc, ++acnt) {
        const int slot = acnt % kHASlots;
        mbar_wait(&a_full[slot], (acnt / kHASlots) & 1);
        const uint32_t a_base = smem_a0 + (uint32_t)(slot * kRPlaneStride);
#pragma unroll 1
        for (int t = 0; t < 9; ++t, ++wcnt) {
          const int ws = wcnt % WST;
          const uint32_t shift = (uint32_t)(((t / 3) * kRW + (t % 3)) * 128);   // tap (dy, dx) = (t/3 - 1, t%3 - 1)
          mbar_wait(&w_full[ws], (wcnt / WST) & 1);
          tc_fence_after();
          const uint64_t a_desc = kmajor_desc_at(kHiA, a_base + shift);
          const uint64_t b_desc = kmajor_desc_at(kHiB, smem_b0 + (uint32_t)(ws * kBStageBytes));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_d, a_desc + (uint64_t)((k * 32) >> 4), b_desc + (uint64_t)((k * 32) >> 4), p.idesc, (kc | t | k) != 0 ? 1u : 0u);
            umma_commit(&w_empty[ws]);
          }
        }
        if (elect_one()) umma_commit(&a_empty[slot]);
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);
    }
  } else if (warp >= 4) {
    // ===================== epilogue (one pixel x 256 channels per lane, as conv_tc_ares_kernel) =====================
    const EpiParams& e = p.e;
    const int group = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gtid = threadIdx.x - (128 + group * 128);
    float* sp = epi_smem + group * kEpiFloats;
    float* s_d = sp;
    float* s_b = sp + CO;
    float* s_n = sp + 2 * CO;
    float* s_w = sp + 3 * CO;
    constexpr float kSqrt2 = 1.4142135623730951f;
    const float nw = (e.noise != nullptr && e.noise_w != nullptr) ? __ldg(e.noise_w) * kSqrt2 : 0.f;
    const int64_t plane = (int64_t)p.out_H * p.out_W;
    const int lx = row & 7, ly = row >> 3;
    uint32_t grp_phase = 0;
    int staged_key = -1;
    // 16-byte stores at pixel pitch (512 B) are one L1 wavefront per lane: 4096 per tile next to the MMA operand traffic.  The
    // activation leaves through a per-warp staging tile and one TMA tensor store per 32-channel chunk instead (2 x 512 wavefronts).
    uint8_t* stage_tile = smem_out + (warp - 4) * 2048;
    __nv_bfloat16* stage_row = (__nv_bfloat16*)stage_tile + lane * 32;     // this lane's 64-byte row
    const int stage_swz = (lane >> 1) & 3;                                 // SWIZZLE_64B: 16-byte piece ^= address bits [7:8]
    const bool staged_out = tma_store != 0 && e.out != nullptr && e.s_next != nullptr;
    for (int it = group; t_begin + it < t_end; it += GROUPS) {
      int nt, x0, y0, b;
      decode(t_begin + it, nt, x0, y0, b);
      const int ox = x0 + lx, oy = y0 + ly;
      const bool ok = ox < p.W && oy < p.H;
      const int co0 = nt * N;

      if (b * tiles_n + nt != staged_key) {
        ares_group_sync(group);
        for (int j = gtid; j < CO; j += 128) {
          s_d[j] = (e.demod != nullptr ? __ldg(e.demod + (int64_t)b * e.demod_bs + co0 + j) : 1.f) * kSqrt2;
          s_b[j] = __ldg(e.bias + co0 + j) * kSqrt2;
          s_n[j] = e.s_next ? __ldg(e.s_next + (int64_t)b * e.s_next_bs + co0 + j) : 1.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) s_w[c * CO + j] = e.wr ? __ldg(e.wr + (int64_t)b * e.wr_bs + c * p.Cout + co0 + j) : 0.f;
        }
        ares_group_sync(group);
        staged_key = b * tiles_n + nt;
      }

      float nz = 0.f;
      float up[3] = {0.f, 0.f, 0.f};
      if (ok && e.noise != nullptr) nz = nw * __ldg(e.noise + (int64_t)b * e.noise_bs + (int64_t)oy * p.out_W + ox);
      if (ok && e.fused_skip) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          up[c] = __ldg(e.rgb_bias + c);
          if (e.skip_in != nullptr)
            up[c] += upsample2x_at(e.skip_in + ((int64_t)b * 3 + c) * (plane / 4), p.out_H / 2, p.out_W / 2, oy, ox, e.fir);
        }
      }

      mbar_wait(&tmem_full[it & 1], grp_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((it & 1) * N);
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
      const int64_t pix = ((int64_t)b * p.out_H + oy) * p.out_W + ox;
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        __nv_bfloat16* outc = nullptr;
        __nv_bfloat16* yc = nullptr;
        if (ok) {
          if (!staged_out && e.out != nullptr && e.s_next != nullptr) outc = (__nv_bfloat16*)e.out + pix * p.Cout + co0 + c0;
          if (e.y_out != nullptr) yc = (__nv_bfloat16*)e.y_out + pix * p.Cout + co0 + c0;
        }
        if (staged_out) {
          if (lane == 0) tma_store_wait_read();         // the previous store has finished reading this warp's staging tile
          __syncwarp();
          outc = stage_row;
        }
        tmem_ld_wait();
        epilogue_chunk32<EPI_ACT_RGB>(v, s_d + c0, s_b + c0, s_n + c0, s_w + c0, s_w + CO + c0, s_w + 2 * CO + c0, nz, false, rgb0, rgb1,
                                      rgb2, outc, yc, staged_out ? stage_swz : 0);
        if (staged_out) {   // the warp's 8 columns x 4 rows; the tensor map clips what lies outside the image
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_4d(&tmap_o, stage_tile, co0 + c0, x0, y0 + 4 * q, b);
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[it & 1]);
      grp_phase ^= 1;

      if (e.wr != nullptr && ok) {
        const float r3[3] = {rgb0, rgb1, rgb2};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (e.fused_skip) e.skip_out[((int64_t)b * 3 + c) * plane + (int64_t)oy * p.out_W + ox] = r3[c] + up[c];
          else e.rgb_part[(((int64_t)nt * p.B + b) * 3 + c) * plane + (int64_t)oy * p.out_W + ox] = r3[c];
        }
      }
    }
    if (staged_out && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool conv_tc_ares_supported(const ConvGeom& g, const EpiParams& e) {
  if (!g_switches.ares || !tmap_available()) return false;
  if (g.nphase != 1 || g.in_scale != 1 || g.weight_taps != 9 || g.in_pair_packed || g.out_pair_packed) return false;
  if (g.Cin != 128 || g.H < 16 || g.W < 16 || e.mode != 0) return false;
  if (g.up_cout > 0) return g.up_cout == 64 && g.Cout == 256 && e.wr == nullptr;
  return g.Cout == 128 && e.wr != nullptr && e.fused_skip;
}

// w: [9][N][128] bf16 (plain: N = Cout; composite: N = 4 * Cout rows (phase, co))
int launch_conv_tc_ares(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st) {
  AresParams p{};
  p.B = g.B; p.H = g.H; p.W = g.W; p.out_H = g.out_H; p.out_W = g.out_W; p.e = e;
  const bool comp = g.up_cout > 0;
  p.Cout = comp ? g.up_cout : g.Cout;
  p.idesc = make_idesc_bf16(128, g.Cout, 0);
  CUtensorMap ta, tw;
  {
    const uint64_t dims[4] = {(uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
    const uint64_t str[4] = {2, (uint64_t)g.Cin * 2, (uint64_t)g.W * g.Cin * 2, (uint64_t)g.H * g.W * g.Cin * 2};
    const uint32_t box[4] = {64, kRW, kRH, 1};
    L2I_TRY(make_tmap(&ta, in, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    const uint64_t dims[3] = {(uint64_t)g.Cin, (uint64_t)g.Cout, 9};
    const uint64_t str[3] = {2, (uint64_t)g.Cin * 2, (uint64_t)g.Cout * g.Cin * 2};
    const uint32_t box[3] = {64, (uint32_t)g.Cout, 1};
    L2I_TRY(make_tmap(&tw, w, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  p.tiles_x = ceil_div(g.W, kRTileW); p.tiles_y = ceil_div(g.H, kRTileH);
  const int64_t total = (int64_t)p.tiles_x * p.tiles_y * g.B;
  if (total <= 0 || total > 0x7fffffff) { set_error("conv_tc_ares: bad tile count"); return L2I_ERR_INVALID_ARG; }
  p.total_tiles = (int)total;
  // ring depth: the weight tiles in flight must cover the L2 latency (~1.5-2k clocks): 7 x 256 / 4 x 512 MMA clocks
  if (comp) return launch_ares_variant<256, 2, 1, 4, 2, EPI_ACT, true>(ta, tw, p, st);
  if (g_switches.ares_pair && (e.out == nullptr || (uintptr_t)e.out % 16 == 0)) {
    CUtensorMap to = ta;   // placeholder when the layer has no activation output
    if (e.out != nullptr && e.s_next != nullptr) {
      // a warp's 32 pixels of a tile: 8 columns x 4 rows, 16 of the 128 channels per store
      const uint64_t dims[4] = {(uint64_t)g.Cout, (uint64_t)g.out_W, (uint64_t)g.out_H, (uint64_t)g.B};
      const uint64_t str[4] = {2, (uint64_t)g.Cout * 2, (uint64_t)g.out_W * g.Cout * 2, (uint64_t)g.out_H * g.out_W * g.Cout * 2};
      const uint32_t box[4] = {16, 8, 4, 1};
      L2I_TRY(make_tmap(&to, e.out, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE));
    }
    return launch_ares_pair<3>(ta, tw, to, p, st);
  }
  return launch_ares_variant<128, 2, 2, 3, 2, EPI_ACT_RGB, false>(ta, tw, p, st);
}

// Wide plain layers (act + ToRGB epilogue): Cin a multiple of 64 from 256 up, Cout a multiple of 256, unsplit weights.
bool conv_tc_hring_supported(const ConvGeom& g, const EpiParams& e) {
  if (!g_switches.hring || !tmap_available()) return false;
  if (g.nphase != 1 || g.in_scale != 1 || g.weight_taps != 9 || g.in_pair_packed || g.out_pair_packed || g.up_cout != 0) return false;
  if (g.Cin < 256 || g.Cin % 64 != 0 || g.Cout % kHN != 0 || g.H < 16 || g.W < 8 || e.mode != 0 || e.wr == nullptr) return false;
  return g.OH == g.H && g.OW == g.W;
}

// w: [9][Cout][Cin] bf16 (the general kernel's packing)
int launch_conv_tc_hring(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st) {
  AresParams p{};
  p.B = g.B; p.H = g.H; p.W = g.W; p.out_H = g.out_H; p.out_W = g.out_W; p.e = e;
  p.Cout = g.Cout;
  p.idesc = make_idesc_bf16(128, kHN, 0);
  const int tiles_n = g.Cout / kHN;
  if (e.fused_skip && tiles_n != 1) { set_error("conv_tc_hring: fused skip needs a single N tile"); return L2I_ERR_INVALID_ARG; }
  CUtensorMap ta, tw;
  {
    const uint64_t dims[4] = {(uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
    const uint64_t str[4] = {2, (uint64_t)g.Cin * 2, (uint64_t)g.W * g.Cin * 2, (uint64_t)g.H * g.W * g.Cin * 2};
    const uint32_t box[4] = {64, kRW, kRH, 1};
    L2I_TRY(make_tmap(&ta, in, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    const uint64_t dims[3] = {(uint64_t)g.Cin, (uint64_t)g.Cout, 9};
    const uint64_t str[3] = {2, (uint64_t)g.Cin * 2, (uint64_t)g.Cout * g.Cin * 2};
    const uint32_t box[3] = {64, kHN, 1};
    L2I_TRY(make_tmap(&tw, w, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  p.tiles_x = ceil_div(g.W, kRTileW); p.tiles_y = ceil_div(g.H, kRTileH);
  const int64_t total = (int64_t)p.tiles_x * p.tiles_y * g.B * tiles_n;
  if (total <= 0 || total > 0x7fffffff) { set_error("conv_tc_hring: bad tile count"); return L2I_ERR_INVALID_ARG; }
  p.total_tiles = (int)total;
  constexpr int WST = 4;   // 4 x 512 MMA clocks in flight cover the L2 latency of a weight tile; three halo planes = 3 x 4608 clocks
  constexpr int smem = kHASlots * kRPlaneStride + WST * kHN * 128 + 8 * 2048 + 1024;
  static_assert(smem + 2 * 6 * kHN * 4 + 256 <= 227 * 1024, "shared memory budget (dynamic + static epilogue vectors)");
  CUtensorMap to = ta;   // placeholder when the layer has no activation output / the staged path is off
  int tma_store = 0;
  if (g_switches.hring_store && e.out != nullptr && e.s_next != nullptr && (uintptr_t)e.out % 16 == 0) {
    const uint64_t dims[4] = {(uint64_t)g.Cout, (uint64_t)g.out_W, (uint64_t)g.out_H, (uint64_t)g.B};
    const uint64_t str[4] = {2, (uint64_t)g.Cout * 2, (uint64_t)g.out_W * g.Cout * 2, (uint64_t)g.out_H * g.out_W * g.Cout * 2};
    const uint32_t box[4] = {32, 8, 4, 1};
    L2I_TRY(make_tmap(&to, e.out, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B));
    tma_store = 1;
  }
  auto kern = conv_tc_hring_kernel<WST>;
  static bool attr_set = false;
  if (!attr_set) {
    L2I_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int grid = std::min(p.total_tiles, kNumSMs);
  kern<<<grid, 128 + 2 * 128, smem, st>>>(ta, tw, to, p, g.Cin / 64, tiles_n, tma_store);
  return check_launch("conv_tc_hring");
}

}  // namespace l2i
