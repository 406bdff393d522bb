// Halo-resident tcgen05 implicit-GEMM conv for the bandwidth-bound high-resolution layers
// (Cin <= 64): the activation tile is fetched ONCE per output tile (one TMA box load of the
// 16 x 18 "unit" halo), the whole weight tensor stays resident in shared memory for the lifetime of
// the persistent CTA, and the 3x3 taps are realised by UMMA shared-memory descriptors that start at
// a tap-shifted row of the halo tile.  Compared with the per-tap kernel (conv_tc.cu) this cuts the
// L2 -> SMEM traffic from 9x to 2.25x of the activation bytes and the TMA / mbarrier round trips per
// tile from 9-18 to 1.
//
// A "unit" is one 128-byte row of the A operand (64 bf16 of K):
//   * Cin = 64 : unit = one pixel.           plain conv  -> 9 taps, 1 accumulator
//                                            up-conv     -> 9 taps spread over the 4 output phases
//                                                           (4 accumulators share the halo tile)
//   * Cin = 32 : unit = a vertical pixel pair (5-D tensor map view [ch, parity, x, row-pair, b]), so rows
//                stay 128 bytes wide (SWIZZLE_128B).  The 3x3 conv becomes 2 row-taps x 3 column-taps per
//                output-row parity with zero-padded [32|32]-channel weight tiles: 12 taps, 2 accumulators.
// Output tile = 8 units wide x 16 units tall = 128 UMMA rows; an 8-row UMMA group is one image row
// of the tile and the halo tile pitch is 16 units (2048 B), so the descriptor's stride-byte-offset is
// a whole number of 1024-byte swizzle atoms and only the start address (tap shift) moves.
#include "tc_epilogue.cuh"

namespace l2i {

using namespace tc;

struct HaloTap {
  int8_t dy, dx;     // unit offsets (-1..1)
  int8_t wtile;      // resident weight tile
  int8_t acc;        // accumulator this tap feeds
};

struct HaloParams {
  int B;
  int OUH, OUW;            // output unit grid covered by tiles
  int Cout;                // N
  int ntaps, nacc, nwtiles;
  HaloTap taps[16];
  int out_mode;            // 0 plain: (Y,X)=(oy,ox); 1 up: (2oy+py, 2ox+px), acc = py*2+px; 2 pair: (2oy+acc, ox)
                           // 3 composite up-conv: one accumulator of 4*Cout columns, column chunk = phase (py*2+px)
  int pair_out;            // composite: write [B][H][2W][2][32] vertical pixel pairs for the following Cin == 32 layer
  int out_H, out_W;
  int tiles_x, tiles_y, total_tiles;
  uint32_t idesc;
  int tma_store;           // composite: write the output through per-warp SWIZZLE_64B staging tiles + strided TMA stores
  int base_offset_mode;    // debug only; measured on B200: the swizzle follows absolute smem address bits, so tap-shifted
                           // start addresses need base_offset = 0 (setting (start >> 7) & 7 gives wrong results)
  EpiParams e;
};

namespace {

constexpr int kHaloH = 18;                              // halo tile rows (units); its width HW is 16 (power-of-two pitch) or 10
constexpr int kTileW = 8, kTileH = 16;
constexpr int halo_stage_bytes(int hw) { return (hw * kHaloH * 128 + 1023) & ~1023; }   // 36864 / 23552

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                    // version
  d |= (uint64_t)(base_offset & 7) << 49;    // matrix base offset
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void halo_group_sync(int group) {
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// N = GEMM N (Cout, or 4 * Cout for the composite up-conv), STAGES = A halo stages, PAIR = Cin 32 pair-packed
// units, COMP = composite up-conv (N = 4 phases x 32 channels, fused blur: SURVEY 0.6b), EPI = fused epilogue kind
// HW = halo tile width in units (the UMMA 8-row-group stride is HW * 128 bytes), kGroups = epilogue warpgroups
template <int N, int STAGES, int NACC, bool PAIR, bool COMP, int EPI, int HW, int kGroups>
__global__ void __launch_bounds__(128 + kGroups * 128, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ HaloParams p) {
  constexpr int kHaloW = HW;
  constexpr int kHaloBytes = HW * kHaloH * 128;            // bytes one halo box load delivers
  constexpr int kStageBytes = halo_stage_bytes(HW);
  constexpr int kTmemAlloc = kGroups * NACC * N <= 32 ? 32 : (kGroups * NACC * N <= 64 ? 64 : (kGroups * NACC * N <= 128 ? 128 : (kGroups * NACC * N <= 256 ? 256 : 512)));
  constexpr int kWTileBytes = N * 128;
  constexpr int kTmemCols = kTmemAlloc;
  static_assert(kGroups * NACC * N <= 512, "TMEM budget");
  constexpr int CO = COMP ? N / 4 : N;   // distinct output channels whose epilogue vectors are staged
  constexpr int kEpiFloats = 6 * CO;
  static_assert(!COMP || (CO == 32 && NACC == 1 && EPI == EPI_ACT), "composite variant: 4 phases x 32 channels in one accumulator");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // A halo stages, then the resident weights
  uint8_t* smem_w = smem + STAGES * kStageBytes;
  const int wbytes = p.nwtiles * kWTileBytes;
  constexpr int kStageOutBytes = COMP ? 2048 : 32 * N * 2;         // per epilogue warp: 32 pixels x one output row (64 B / 2N B)
  uint8_t* smem_stage_out = smem_w + ((wbytes + 1023) & ~1023);   // tma_store: swizzled output staging tiles
  __shared__ __align__(16) float epi_smem[kGroups * kEpiFloats];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full[kGroups];
  __shared__ __align__(8) uint64_t tmem_empty[kGroups];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    if (p.tma_store) prefetch_tmap(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < kGroups; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
    mbar_init(&w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  auto decode = [&](int tile, int& x0, int& y0, int& b) {
    const int tx = tile % p.tiles_x;
    const int r = tile / p.tiles_x;
    const int ty = r % p.tiles_y;
    b = r / p.tiles_y;
    x0 = tx * kTileW; y0 = ty * kTileH;
  };

  if (warp == 0) {
    // ===================== TMA producer: weights once, then one halo tile per output tile ==========
    if (lane == 0) {
      mbar_expect_tx(&w_bar, (uint32_t)wbytes);
      for (int t = 0; t < p.nwtiles; ++t) tma_load_3d(smem_w + t * kWTileBytes, &tmap_w, &w_bar, 0, 0, t);
      int stage = 0;
      uint32_t phase_bit = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int x0, y0, b;
        decode(tile, x0, y0, b);
        mbar_wait(&empty_bar[stage], phase_bit ^ 1);
        mbar_expect_tx(&full_bar[stage], kHaloBytes);
        tma_load_4d(smem + stage * kStageBytes, &tmap_a, &full_bar[stage], 0, x0 - 1, y0 - 1, b);
        if (++stage == STAGES) { stage = 0; phase_bit ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform control flow, one elected lane issues (tc_ptx.cuh: elect_one) =============
    {
      mbar_wait(&w_bar, 0);
      tc_fence_after();
      const uint32_t w_base = smem_u32(smem_w), smem_a0 = smem_u32(smem);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint64_t kHiA = kmajor_desc_hi(kHaloW * 128, 2), kHiB = kmajor_desc_hi(1024, 2);
      const uint64_t b_desc0 = kmajor_desc_at(kHiB, w_base);
      int stage = 0;
      uint32_t phase_bit = 0;
      int grp = 0;
      uint32_t grp_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[grp], grp_phase ^ 1);
        mbar_wait(&full_bar[stage], phase_bit);
        tc_fence_after();
        const uint32_t a_base = smem_a0 + (uint32_t)(stage * kStageBytes);
        const uint32_t tmem_d = tmem_u + (uint32_t)(grp * NACC * N);
        if (elect_one()) {
          uint32_t started = 0;  // accumulators that already received their first MMA
          for (int t = 0; t < p.ntaps; ++t) {
            const HaloTap tap = p.taps[t];
            const uint32_t a_tap = a_base + (uint32_t)(((tap.dy + 1) * kHaloW + (tap.dx + 1)) * 128);
            const uint64_t a_desc = kmajor_desc_at(kHiA, a_tap) | ((uint64_t)(p.base_offset_mode ? ((a_tap >> 7) & 7u) : 0u) << 49);
            const uint64_t b_desc = b_desc0 + (uint64_t)((uint32_t)(tap.wtile * kWTileBytes) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_d + (uint32_t)(tap.acc * N), a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), p.idesc,
                        ((started >> tap.acc) & 1u) | (k != 0 ? 1u : 0u));
            started |= 1u << tap.acc;
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&tmem_full[grp]);
        }
        if (++stage == STAGES) { stage = 0; phase_bit ^= 1; }
        if (++grp == kGroups) { grp = 0; grp_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (kGroups warpgroups, one tile each in flight) =====================
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
    const float nw = (EPI != EPI_RAW && e.noise != nullptr && e.noise_w != nullptr) ? __ldg(e.noise_w) * kSqrt2 : 0.f;
    const int64_t plane = (int64_t)p.out_H * p.out_W;
    const int lx = row & 7, ly = row >> 3;
    const bool raw_fp16 = e.raw_fp16 != 0;
    const bool use_tma_store = p.tma_store != 0;
    uint8_t* stage_tile = smem_stage_out + (warp - 4) * kStageOutBytes;
    // this lane's row of the staging tile: 64 B rows under SWIZZLE_64B (chunk ^= address bits [7:8]) for the composite
    // kernel, 2N-byte (128 B) rows under SWIZZLE_128B (chunk ^= row & 7) for the plain one
    __nv_bfloat16* stage_out = (__nv_bfloat16*)stage_tile + lane * (COMP ? 32 : N);
    const int stage_swz = COMP ? ((lane >> 1) & 3) : (lane & 7);
    uint32_t grp_phase = 0;
    int staged_b = -1;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      if (it % kGroups != group) continue;
      int x0, y0, b;
      decode(tile, x0, y0, b);
      const int ox = x0 + lx, oy = y0 + ly;
      const bool in_grid = ox < p.OUW && oy < p.OUH;

      if (b != staged_b) {
        halo_group_sync(group);
        for (int j = gtid; j < CO; j += 128) {
          const float d = e.demod != nullptr ? __ldg(e.demod + (int64_t)b * e.demod_bs + j) : 1.f;
          if (EPI != EPI_RAW) {
            s_d[j] = d * kSqrt2;
            s_b[j] = __ldg(e.bias + j) * kSqrt2;
            s_n[j] = e.s_next ? __ldg(e.s_next + (int64_t)b * e.s_next_bs + j) : 1.f;
            if (EPI == EPI_ACT_RGB) {
#pragma unroll
              for (int c = 0; c < 3; ++c) s_w[c * CO + j] = e.wr ? __ldg(e.wr + (int64_t)b * e.wr_bs + c * p.Cout + j) : 0.f;
            }
          } else {
            s_d[j] = d;
          }
        }
        halo_group_sync(group);
        staged_b = b;
      }

      // output pixel of each accumulator + every global load, before waiting for the MMAs
      int Ys[NACC], Xs[NACC];
      bool ok[NACC];
      float nz[NACC], up[NACC][3];
      float nzq[4] = {0.f, 0.f, 0.f, 0.f};   // composite: noise of the 2x2 output quad of this input pixel
      if (COMP && in_grid && e.noise != nullptr) {
        const float* np = e.noise + (int64_t)b * e.noise_bs + (int64_t)(2 * oy) * p.out_W + 2 * ox;
        const float2 n01 = __ldg(reinterpret_cast<const float2*>(np));
        const float2 n23 = __ldg(reinterpret_cast<const float2*>(np + p.out_W));
        nzq[0] = nw * n01.x; nzq[1] = nw * n01.y; nzq[2] = nw * n23.x; nzq[3] = nw * n23.y;
      }
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        if (COMP) { Ys[a] = 2 * oy; Xs[a] = 2 * ox; }
        else if (p.out_mode == 1) { Ys[a] = 2 * oy + (a >> 1); Xs[a] = 2 * ox + (a & 1); }
        else if (p.out_mode == 2) { Ys[a] = 2 * oy + a; Xs[a] = ox; }
        else { Ys[a] = oy; Xs[a] = ox; }
        ok[a] = in_grid && Ys[a] < p.out_H && Xs[a] < p.out_W;
        nz[a] = 0.f;
        up[a][0] = up[a][1] = up[a][2] = 0.f;
        if (!COMP && EPI != EPI_RAW && ok[a]) {
          if (e.noise != nullptr) nz[a] = nw * __ldg(e.noise + (int64_t)b * e.noise_bs + (int64_t)Ys[a] * p.out_W + Xs[a]);
          if (EPI == EPI_ACT_RGB && e.fused_skip) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              up[a][c] = __ldg(e.rgb_bias + c);
              if (e.skip_in != nullptr)
                up[a][c] += upsample2x_at(e.skip_in + ((int64_t)b * 3 + c) * (plane / 4), p.out_H / 2, p.out_W / 2, Ys[a], Xs[a], e.fir);
            }
          }
        }
      }

      mbar_wait(&tmem_full[group], grp_phase);
      tc_fence_after();
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((group * NACC + a) * N);
        const int64_t pix = ((int64_t)b * p.out_H + Ys[a]) * p.out_W + Xs[a];
        __nv_bfloat16* outp = nullptr;
        if (!COMP && ok[a] && e.out != nullptr && (EPI == EPI_RAW || e.s_next != nullptr)) outp = (__nv_bfloat16*)e.out + pix * p.Cout;
        __nv_bfloat16* yp = nullptr;
        if (!COMP && ok[a] && EPI != EPI_RAW && e.y_out != nullptr) yp = (__nv_bfloat16*)e.y_out + pix * p.Cout;
        float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          const int cs = COMP ? 0 : c0;        // offset of this chunk's channels in the staged vectors
          float nzc = nz[a];
          __nv_bfloat16* outc = outp != nullptr ? outp + c0 : nullptr;
          __nv_bfloat16* yc = yp != nullptr ? yp + c0 : nullptr;
          if (COMP) {                          // chunk = output phase
            const int phc = c0 >> 5;
            nzc = phc == 0 ? nzq[0] : (phc == 1 ? nzq[1] : (phc == 2 ? nzq[2] : nzq[3]));
            if (use_tma_store) {
              // the warp's 4 x 8 input pixels -> output pixels (2oy+py, 2ox+px): stage them as a [4][8][32] bf16 tile and let
              // one strided TMA tensor store write it (full sectors, no LSU store wavefronts)
              if (lane == 0) tma_store_wait_read();   // the previous store has finished reading the staging tile
              __syncwarp();
              tmem_ld_wait();
              epilogue_chunk32<EPI>(v, s_d, s_b, s_n, s_w, s_w + CO, s_w + 2 * CO, nzc, raw_fp16, rgb0, rgb1, rgb2, stage_out,
                                    nullptr, stage_swz);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0)
                tma_store_4d(&tmap_o, stage_tile, 0, 2 * x0 + (phc & 1), 2 * (y0 + 4 * q) + (phc >> 1), b);
              continue;
            }
            if (ok[a]) {
              const int Xc = Xs[a] + (phc & 1);
              const int64_t pixc = ((int64_t)b * p.out_H + Ys[a] + (phc >> 1)) * p.out_W + Xc;
              const int64_t opix = p.pair_out ? ((((int64_t)b * (p.out_H >> 1) + oy) * p.out_W + Xc) * 2 + (phc >> 1)) : pixc;
              if (e.out != nullptr) outc = (__nv_bfloat16*)e.out + opix * CO;
              if (e.y_out != nullptr) yc = (__nv_bfloat16*)e.y_out + pixc * CO;
            }
          }
          if (!COMP && use_tma_store) {
            // plain N = 64 layer: the warp's 4 x 8 pixels x 128 B go through a SWIZZLE_128B staging tile and one TMA store
            if (c0 == 0) {
              if (lane == 0) tma_store_wait_read();
              __syncwarp();
            }
            tmem_ld_wait();
            epilogue_chunk32<EPI>(v, s_d + cs, s_b + cs, s_n + cs, s_w + cs, s_w + CO + cs, s_w + 2 * CO + cs, nzc, raw_fp16,
                                  rgb0, rgb1, rgb2, stage_out, nullptr, stage_swz, c0 >> 3);
            if (c0 + 32 >= N) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) tma_store_4d(&tmap_o, stage_tile, 0, x0, y0 + 4 * q, b);
            }
            continue;
          }
          tmem_ld_wait();
          epilogue_chunk32<EPI>(v, s_d + cs, s_b + cs, s_n + cs, s_w + cs, s_w + CO + cs, s_w + 2 * CO + cs, nzc, raw_fp16,
                                rgb0, rgb1, rgb2, outc, yc);
        }
        if (EPI == EPI_ACT_RGB && e.wr != nullptr && ok[a]) {
          const float r3[3] = {rgb0, rgb1, rgb2};
          if (e.fused_skip) {
#pragma unroll
            for (int c = 0; c < 3; ++c) e.skip_out[((int64_t)b * 3 + c) * plane + (int64_t)Ys[a] * p.out_W + Xs[a]] = r3[c] + up[a][c];
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) e.rgb_part[((int64_t)b * 3 + c) * plane + (int64_t)Ys[a] * p.out_W + Xs[a]] = r3[c];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[group]);
      grp_phase ^= 1;
    }
  }

  if (p.tma_store && warp >= 4 && lane == 0) tma_store_wait_all();   // outstanding bulk stores must land before exit
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int N, int STAGES, int NACC, bool PAIR, bool COMP, int EPI, int HW, int kGroups>
int launch_halo_variant(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const HaloParams& p, cudaStream_t st) {
  const int wbytes = (p.nwtiles * N * 128 + 1023) & ~1023;
  // + alignment slack; epilogue vectors and barriers are static; composite: 2 KB output staging tile per epilogue warp
  const int smem = STAGES * halo_stage_bytes(HW) + wbytes + 1024 + (p.tma_store ? kGroups * 4 * (COMP ? 2048 : 32 * N * 2) : 0);
  constexpr int kThreads = 128 + kGroups * 128;
  if (smem > 227 * 1024) {
    set_error("conv_tc_halo: shared memory budget exceeded (%d bytes)", smem);
    return L2I_ERR_UNSUPPORTED;
  }
  auto kern = conv_tc_halo_kernel<N, STAGES, NACC, PAIR, COMP, EPI, HW, kGroups>;
  static int attr_smem = 0;
  if (attr_smem < smem) {
    L2I_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  const int grid = std::min(p.total_tiles, kNumSMs);
  kern<<<grid, kThreads, smem, st>>>(ta, tw, to, p);
  return check_launch("conv_tc_halo");
}


}  // namespace

// Which layers this kernel takes: plain 64->64 / 64->32 ... with Cin == 64, Cout in {32, 64}; the
// stride-2 transposed conv with Cin == 64; and Cin == 32 -> 32 plain through the pair-packed view.
bool conv_tc_halo_supported(const ConvGeom& g, const EpiParams& e) {
  const int g_halo_mask = g_switches.halo_mask;
  if (!g_switches.halo || !tmap_available()) return false;
  if (g.in_scale != 1 || g.weight_taps != 9) return false;
  if (g.H < 16 || g.W < 16) return false;
  if (g.up_cout > 0)   // composite up-conv 64 -> 32: GEMM N = 128, weights (9 x 16 KB) resident
    return g.nphase == 1 && g.Cin == 64 && g.up_cout == 32 && g.Cout == 128 && e.mode == 0 && e.wr == nullptr && (g_halo_mask & 8) != 0;
  if (!(g.Cout == 32 || g.Cout == 64)) return false;
  if (g.nphase == 1 && (e.mode != 0 || e.wr == nullptr)) return false;   // plain variants are compiled with the act + ToRGB epilogue
  if (g.nphase == 4 && e.mode != 1) return false;                        // transposed-conv variant writes the raw t tensor
  if (g.nphase == 1 && g.Cin == 64) return (g_halo_mask & 1) != 0;
  if (g.nphase == 4 && g.Cin == 64 && g.Cout == 32) return (g_halo_mask & 2) != 0;   // 4 accumulators x 32 columns x 4 groups = 512 TMEM columns
  if (g.nphase == 1 && g.Cin == 32 && g.Cout == 32 && (g.H % 2 == 0)) return (g_halo_mask & 4) != 0;  // caller provides pair-packed input
  return false;
}

// w: [9][Cout][Cin] bf16 (Cin = 64), or for the pair-packed Cin = 32 case the pre-packed [12][32][64] tiles.
int launch_conv_tc_halo(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st) {
  HaloParams p{};
  p.B = g.B; p.Cout = g.Cout; p.out_H = g.out_H; p.out_W = g.out_W; p.e = e;
  p.pair_out = g.out_pair_packed;
  p.base_offset_mode = g_switches.halo_base_offset;
  p.idesc = make_idesc_bf16(128, g.Cout, 0);
  const bool pair = g.Cin == 32;
  const bool comp = g.up_cout > 0;
  // plain 64 -> 64 layer whose activation output goes through TMA stores (inference: no y_out, single sample layout)
  const bool plain64 = !comp && !pair && g.nphase == 1 && g.Cout == 64 && e.mode == 0 && e.out != nullptr && e.s_next != nullptr &&
                       e.y_out == nullptr && (uintptr_t)e.out % 16 == 0 && (g.OW % 8 == 0) && (g.OH % 4 == 0);
  const int hw = (comp || plain64) ? 10 : 16;   // tight 10-unit halo pitch frees shared memory for the output staging tiles
  CUtensorMap ta, tw, to;
  {
    // pair mode: the producer wrote [B][H/2][W][2][32], i.e. an ordinary NHWC tensor of H/2 x W units with 64 "channels"
    const uint64_t uh = pair ? (uint64_t)g.H / 2 : (uint64_t)g.H;
    const uint64_t dims[4] = {64, (uint64_t)g.W, uh, (uint64_t)g.B};
    const uint64_t str[4] = {2, 128, (uint64_t)g.W * 128, uh * g.W * 128};
    const uint32_t box[4] = {64, (uint32_t)hw, kHaloH, 1};
    L2I_TRY(make_tmap(&ta, in, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  if (g.up_cout > 0) {  // composite up-conv: a plain 3x3 conv at input resolution with N = 4 * Cout
    p.out_mode = 3; p.nacc = 1; p.nwtiles = 9; p.ntaps = 0;
    p.OUH = g.OH; p.OUW = g.OW;
    for (int t = 0; t < g.taps[0].n; ++t) p.taps[p.ntaps++] = HaloTap{g.taps[0].dy[t], g.taps[0].dx[t], g.taps[0].wtap[t], 0};
  } else if (g.nphase == 4) {  // stride-2 transposed conv: all four output phases from one halo tile
    p.out_mode = 1; p.nacc = 4; p.nwtiles = 9; p.ntaps = 0;
    p.OUH = g.OH; p.OUW = g.OW;
    for (int ph = 0; ph < 4; ++ph)
      for (int t = 0; t < g.taps[ph].n; ++t)
        p.taps[p.ntaps++] = HaloTap{g.taps[ph].dy[t], g.taps[ph].dx[t], g.taps[ph].wtap[t], (int8_t)ph};
  } else if (!pair) {
    p.out_mode = 0; p.nacc = 1; p.nwtiles = 9; p.ntaps = 0;
    p.OUH = g.OH; p.OUW = g.OW;
    for (int t = 0; t < g.taps[0].n; ++t) p.taps[p.ntaps++] = HaloTap{g.taps[0].dy[t], g.taps[0].dx[t], g.taps[0].wtap[t], 0};
  } else {
    // pair-packed rows: tile index = (parity * 2 + row_tap) * 3 + kw, see pack_pair_weight_kernel
    p.out_mode = 2; p.nacc = 2; p.nwtiles = 12; p.ntaps = 0;
    p.OUH = g.OH / 2; p.OUW = g.OW;
    for (int par = 0; par < 2; ++par)
      for (int r = 0; r < 2; ++r)
        for (int kw = 0; kw < 3; ++kw)
          p.taps[p.ntaps++] = HaloTap{(int8_t)(par == 0 ? r - 1 : r), (int8_t)(kw - 1), (int8_t)((par * 2 + r) * 3 + kw), (int8_t)par};
  }
  {
    const uint64_t dims[3] = {64, (uint64_t)g.Cout, (uint64_t)p.nwtiles};
    const uint64_t str[3] = {2, 128, (uint64_t)g.Cout * 128};
    const uint32_t box[3] = {64, (uint32_t)g.Cout, 1};
    L2I_TRY(make_tmap(&tw, w, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  to = ta;
  p.tma_store = 0;
  if (comp && !g.out_pair_packed && e.y_out == nullptr && e.out != nullptr && (uintptr_t)e.out % 16 == 0) {
    // output NHWC [B][2H][2W][32] bf16; a warp's 4 x 8 input pixels of one phase = every other pixel of 4 x 8 output positions
    const uint64_t dims[4] = {32, (uint64_t)g.out_W, (uint64_t)g.out_H, (uint64_t)g.B};
    const uint64_t str[4] = {2, 64, (uint64_t)g.out_W * 64, (uint64_t)g.out_H * g.out_W * 64};
    const uint32_t box[4] = {32, 16, 8, 1};
    const uint32_t estr[4] = {1, 2, 2, 1};
    L2I_TRY(make_tmap_strided(&to, e.out, 4, dims, str, box, estr, CU_TENSOR_MAP_SWIZZLE_64B));
    p.tma_store = 1;
  }
  if (plain64) {
    const uint64_t dims[4] = {64, (uint64_t)g.out_W, (uint64_t)g.out_H, (uint64_t)g.B};
    const uint64_t str[4] = {2, 128, (uint64_t)g.out_W * 128, (uint64_t)g.out_H * g.out_W * 128};
    const uint32_t box[4] = {64, 8, 4, 1};
    L2I_TRY(make_tmap(&to, e.out, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
    p.tma_store = 1;
  }
  p.tiles_x = ceil_div(p.OUW, kTileW); p.tiles_y = ceil_div(p.OUH, kTileH);
  const int64_t total = (int64_t)p.tiles_x * p.tiles_y * g.B;
  if (total <= 0 || total > 0x7fffffff) { set_error("conv_tc_halo: bad tile count"); return L2I_ERR_INVALID_ARG; }
  p.total_tiles = (int)total;
  if (comp) return launch_halo_variant<128, 2, 1, false, true, EPI_ACT, 10, 3>(ta, tw, to, p, st);
  if (g.nphase == 4) return launch_halo_variant<32, 4, 4, false, false, EPI_RAW, 16, 4>(ta, tw, to, p, st);
  if (pair) return launch_halo_variant<32, 4, 2, true, false, EPI_ACT_RGB, 16, 4>(ta, tw, to, p, st);
  if (plain64) return launch_halo_variant<64, 3, 1, false, false, EPI_ACT_RGB, 10, 4>(ta, tw, to, p, st);
  if (g.Cout == 64) return launch_halo_variant<64, 3, 1, false, false, EPI_ACT_RGB, 16, 4>(ta, tw, to, p, st);
  return launch_halo_variant<32, 4, 1, false, false, EPI_ACT_RGB, 16, 4>(ta, tw, to, p, st);
}

// Pair-packed weight tiles for Cin = 32: dst [12][Cout][64] bf16, tile (parity*2 + r)*3 + kw holds
//   parity 0 (even output rows): r=0 -> [0 | W(kh=0)]        on unit j-1;   r=1 -> [W(kh=1) | W(kh=2)] on unit j
//   parity 1 (odd  output rows): r=0 -> [W(kh=0) | W(kh=1)]  on unit j;     r=1 -> [W(kh=2) | 0]       on unit j+1
// where a unit is [row 2j channels | row 2j+1 channels].  src: [Cout][32][3][3] fp32.
__global__ void pack_pair_weight_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, int Cout, float scale) {
  const int total = 12 * Cout * 64;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx % 64;
    const int co = (idx / 64) % Cout;
    const int tile = idx / (64 * Cout);
    const int kw = tile % 3, r = (tile / 3) % 2, par = tile / 6;
    const int half = k / 32, ci = k % 32;
    int kh = -1;
    if (par == 0) kh = (r == 0) ? (half == 1 ? 0 : -1) : (half == 0 ? 1 : 2);
    else kh = (r == 0) ? (half == 0 ? 0 : 1) : (half == 0 ? 2 : -1);
    float v = 0.f;
    if (kh >= 0) v = src[(((int64_t)co * 32 + ci) * 3 + kh) * 3 + kw] * scale;
    dst[idx] = __float2bfloat16_rn(v);
  }
}

int launch_pack_pair_weight(__nv_bfloat16* dst, const float* src, int Cout, float scale, cudaStream_t st) {
  pack_pair_weight_kernel<<<ceil_div(12 * Cout * 64, 256), 256, 0, st>>>(dst, src, Cout, scale);
  return check_launch("pack_pair_weight");
}

}  // namespace l2i
