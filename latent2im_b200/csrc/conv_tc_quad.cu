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
// bias + 2x FIR up-sampling of the skip image (ToRGB.forward, networks.py:349-358), float2 stores.  It walks the
// accumulator in groups of 8 channels x 4 pixels so the per-channel vectors are fetched from shared memory once
// per block, and the (8+2) x (16+2) low-resolution skip patch of the tile is TMA-loaded per accumulator stage
// (out-of-image samples zero-filled = upfirdn2d's zero padding), so the epilogue issues no gather loads.
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
// Three epilogue warpgroups (round 2): the layer is bound by its epilogue's instruction stream (trimming the MMAs by 12 % changed
// nothing), two groups at 166 registers could not keep up with the MMAs.  512 threads launch with 128 registers each; the
// producer / MMA warpgroup gives its surplus to the epilogue warps (setmaxnreg 40 / 152).
constexpr int kQGroups = 3;
constexpr int kQTmemCols = 512;                           // 3 accumulators x 128 columns (power-of-two allocation)
constexpr int kQThreads = 128 + kQGroups * 128;
constexpr int kQRegLow = 40, kQRegHigh = 152;
static_assert(128 * kQRegLow + kQGroups * 128 * kQRegHigh <= kQThreads * 128, "setmaxnreg budget");
constexpr int kQN = 128, kQK = 512;
// resident weights: patch rows 1, 2 as four [128 n][64 k] atoms, rows 0 / 3 (which reach only the upper / lower pixel row of a
// block) as four [64 n][64 k] half atoms: 4 x 16 KB + 4 x 8 KB
constexpr int kQWideAtom = kQN * 128, kQHalfAtom = (kQN / 2) * 128;
constexpr int kQWBytes = 4 * kQWideAtom + 4 * kQHalfAtom;   // 98304
constexpr int kQSkipW = 16, kQSkipH = kQTileH / 2 + 2;     // low-res skip patch: columns n0-4 .. n0+11 (TMA needs a 16-byte aligned start), rows m0-1 .. m0+16
constexpr int kQSkipFloats = 3 * kQSkipH * kQSkipW;        // 864 floats = 3456 bytes per accumulator stage
constexpr int kQSkipStride = (kQSkipFloats + 31) & ~31;    // buffers 128-byte aligned (TMA destination)
constexpr int kQSmem = kQStages * kQStageBytes + kQWBytes + 1024;

struct QuadParams {
  int B, H, W;
  int tiles_x, tiles_y, total_tiles;
  uint32_t idesc, idesc64;
  int has_skip;            // tmap_s is valid: the producer loads the low-res skip patch of every tile
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
                    const __grid_constant__ CUtensorMap tmap_wh,
                    const __grid_constant__ CUtensorMap tmap_s, const __grid_constant__ QuadParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_w = smem + kQStages * kQStageBytes;
  __shared__ __align__(16) float epi_smem[kQGroups * 6 * 32];
  __shared__ __align__(8) uint64_t full_bar[kQStages];
  __shared__ __align__(8) uint64_t empty_bar[kQStages];
  __shared__ __align__(8) uint64_t tmem_full[kQGroups];
  __shared__ __align__(8) uint64_t tmem_empty[kQGroups];
  __shared__ __align__(8) uint64_t w_bar;
  __shared__ __align__(8) uint64_t skip_full[kQGroups];
  __shared__ __align__(128) float skip_smem[kQGroups * kQSkipStride];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    if (p.has_skip) prefetch_tmap(&tmap_s);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kQStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < kQGroups; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); mbar_init(&skip_full[a], 1); }
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

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kQRegLow));
  if (warp == 0) {
    // ===================== TMA producer: the weight matrix once, then one halo tile per output tile =====
    if (lane == 0) {
      mbar_expect_tx(&w_bar, (uint32_t)kQWBytes);
      for (int t = 0; t < 4; ++t) tma_load_3d(smem_w + t * kQWideAtom, &tmap_w, &w_bar, 0, 0, 2 + t);        // patch rows 1, 2
      for (int t = 0; t < 2; ++t) {
        tma_load_3d(smem_w + 4 * kQWideAtom + t * kQHalfAtom, &tmap_wh, &w_bar, 0, 0, t);                     // row 0: weight rows 0..63
        tma_load_3d(smem_w + 4 * kQWideAtom + (2 + t) * kQHalfAtom, &tmap_wh, &w_bar, 0, kQN / 2, 6 + t);      // row 3: rows 64..127
      }
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
  } else if (warp == 3) {
    // ===================== skip-patch producer: one small fp32 box per tile into its accumulator stage's buffer ====
    if (lane == 0 && p.has_skip) {
      int grp = 0;
      uint32_t grp_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int x0, y0, b;
        decode(tile, x0, y0, b);
        mbar_wait(&tmem_empty[grp], grp_phase ^ 1);   // the stage's previous epilogue has finished with the buffer
        mbar_expect_tx(&skip_full[grp], kQSkipFloats * 4);
        tma_load_4d(skip_smem + grp * kQSkipStride, &tmap_s, &skip_full[grp], (x0 >> 1) - 4, (y0 >> 1) - 1, 0, b);
        if (++grp == kQGroups) { grp = 0; grp_phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 4 patch rows x 8 K16 steps, one N = 128 accumulator per tile =====
    // warp-uniform control flow, one elected lane issues (tc_ptx.cuh: elect_one)
    {
      mbar_wait(&w_bar, 0);
      tc_fence_after();
      const uint32_t w_base = smem_u32(smem_w), smem_a0 = smem_u32(smem);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint64_t kHiA = kmajor_desc_hi(2 * kQRowPitch, 2), kHiB = kmajor_desc_hi(1024, 2);
      const uint64_t b_desc0 = kmajor_desc_at(kHiB, w_base);
      int stage = 0;
      uint32_t phase_bit = 0;
      int grp = 0;
      uint32_t grp_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[grp], grp_phase ^ 1);
        mbar_wait(&full_bar[stage], phase_bit);
        tc_fence_after();
        const uint64_t a_desc0 = kmajor_desc_at(kHiA, smem_a0 + (uint32_t)(stage * kQStageBytes));
        const uint32_t tmem_d = tmem_u + (uint32_t)(grp * kQN);
        if (elect_one()) {
          // Patch rows 1 and 2 feed both pixel rows of the block (N = 128); row 0 only reaches the upper pixels (a = 0: accumulator
          // columns 0..63, weight rows 0..63) and row 3 only the lower ones (a = 1: columns / rows 64..127) - their other halves are
          // structural zeros, so they are N = 64 MMAs: 48 instead of 64 clocks and 6 instead of 8 KB of operands each.  Rows 1, 2 go
          // first: the first MMA initialises all 128 columns.
#pragma unroll
          for (int ri = 0; ri < 4; ++ri) {
            const int r = ri == 0 ? 1 : (ri == 1 ? 2 : (ri == 2 ? 0 : 3));
            const bool wide = r == 1 || r == 2;
            const uint32_t half = r == 3 ? 64u : 0u;              // accumulator column offset of the N = 64 MMAs
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              // A: rows = blocks along x (128 B apart), 8-row groups = block rows (two image rows apart)
              // B: K-major [128 x 64] atoms, atom = (r*128 + kk*16) / 64
              const int j = kk >> 2;                              // 64-channel atom of the patch row
              const int woff = r == 1 ? j * kQWideAtom : (r == 2 ? (2 + j) * kQWideAtom
                                                                 : 4 * kQWideAtom + ((r == 0 ? 0 : 2) + j) * kQHalfAtom);
              umma_bf16(tmem_d + half, a_desc0 + (uint64_t)((r * kQRowPitch + 64 + kk * 32) >> 4),
                        b_desc0 + (uint64_t)((woff + (kk & 3) * 32) >> 4), wide ? p.idesc : p.idesc64, (ri | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&tmem_full[grp]);
        }
        if (++stage == kQStages) { stage = 0; phase_bit ^= 1; }
        if (++grp == kQGroups) { grp = 0; grp_phase ^= 1; }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kQRegHigh));
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

      // ---- global loads of the block (noise) before waiting for anything ----
      float2 n01 = make_float2(0.f, 0.f), n23 = make_float2(0.f, 0.f);
      if (ok && e.noise != nullptr) {
        const float* np = e.noise + (int64_t)b * e.noise_bs + (int64_t)Y * p.W + X;
        n01 = __ldg(reinterpret_cast<const float2*>(np));
        n23 = __ldg(reinterpret_cast<const float2*>(np + p.W));
      }

      // ---- ToRGB tail: bias + 2x FIR up-sampling of the skip image for the four pixels of the block.  They share the
      // 3x3 low-res patch around (m, n) = (Y/2, X/2): even outputs use rows m-1 (f0), m (f2), odd ones m (f1), m+1 (f3).
      float up[4][3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float bias_c = e.fused_skip ? __ldg(e.rgb_bias + c) : 0.f;
        up[0][c] = up[1][c] = up[2][c] = up[3][c] = bias_c;
      }
      if (p.has_skip) {
        mbar_wait(&skip_full[group], grp_phase);
        const float* sk = skip_smem + group * kQSkipStride + by * kQSkipW + bx + 3;   // patch origin = (m0 - 1, n0 - 4)
        const float f0 = e.fir[0], f1 = e.fir[1], f2 = e.fir[2], f3 = e.fir[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float h0[3], h1[3];   // horizontally filtered rows for even / odd output columns
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float* rowp = sk + (c * kQSkipH + i) * kQSkipW;
            const float p0 = rowp[0], p1 = rowp[1], p2 = rowp[2];
            h0[i] = fmaf(f0, p0, f2 * p1);
            h1[i] = fmaf(f1, p1, f3 * p2);
          }
          up[0][c] += fmaf(f0, h0[0], f2 * h0[1]);
          up[1][c] += fmaf(f0, h1[0], f2 * h1[1]);
          up[2][c] += fmaf(f1, h0[1], f3 * h0[2]);
          up[3][c] += fmaf(f1, h1[1], f3 * h1[2]);
        }
      }
      const float nzq[4] = {nw * n01.x, nw * n01.y, nw * n23.x, nw * n23.y};

      mbar_wait(&tmem_full[group], grp_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(group * kQN);
      const bool want_out = ok && e.out != nullptr && e.s_next != nullptr;
      const bool want_y = ok && e.y_out != nullptr;
      const int64_t pix0 = ((int64_t)b * p.H + Y) * p.W + X;
      float rgb[4][3];
#pragma unroll
      for (int ph = 0; ph < 4; ++ph) rgb[ph][0] = rgb[ph][1] = rgb[ph][2] = 0.f;
#pragma unroll 1   // (unrolling by 2 / 4 to overlap the next group's tcgen05.ld: 0 / -9 %, measured)
      for (int cg = 0; cg < 4; ++cg) {   // 8 channels x the 4 pixels (column chunks) of the block
        uint32_t v[4][8];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) tmem_ld8(taddr + ph * 32 + cg * 8, v[ph]);
        const float4 dA = *reinterpret_cast<const float4*>(s_d + cg * 8), dB = *reinterpret_cast<const float4*>(s_d + cg * 8 + 4);
        const float4 bA = *reinterpret_cast<const float4*>(s_b + cg * 8), bB = *reinterpret_cast<const float4*>(s_b + cg * 8 + 4);
        const float4 r0A = *reinterpret_cast<const float4*>(s_w + cg * 8), r0B = *reinterpret_cast<const float4*>(s_w + cg * 8 + 4);
        const float4 r1A = *reinterpret_cast<const float4*>(s_w + 32 + cg * 8), r1B = *reinterpret_cast<const float4*>(s_w + 32 + cg * 8 + 4);
        const float4 r2A = *reinterpret_cast<const float4*>(s_w + 64 + cg * 8), r2B = *reinterpret_cast<const float4*>(s_w + 64 + cg * 8 + 4);
        const float dd[8] = {dA.x, dA.y, dA.z, dA.w, dB.x, dB.y, dB.z, dB.w};
        const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
        const float w0[8] = {r0A.x, r0A.y, r0A.z, r0A.w, r0B.x, r0B.y, r0B.z, r0B.w};
        const float w1[8] = {r1A.x, r1A.y, r1A.z, r1A.w, r1B.x, r1B.y, r1B.z, r1B.w};
        const float w2[8] = {r2A.x, r2A.y, r2A.z, r2A.w, r2B.x, r2B.y, r2B.z, r2B.w};
        tmem_ld_wait();
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            x[j] = fmaf(__uint_as_float(v[ph][j]), dd[j], bb[j] + nzq[ph]);
            x[j] = fmaxf(x[j], 0.2f * x[j]);
            rgb[ph][0] = fmaf(w0[j], x[j], rgb[ph][0]);
            rgb[ph][1] = fmaf(w1[j], x[j], rgb[ph][1]);
            rgb[ph][2] = fmaf(w2[j], x[j], rgb[ph][2]);
          }
          if (want_out || want_y) {   // not taken for the last layer of the network (no next layer, inference)
            const int64_t pix = pix0 + (ph >> 1) * (int64_t)p.W + (ph & 1);
            if (want_y)
              *reinterpret_cast<uint4*>((__nv_bfloat16*)e.y_out + pix * 32 + cg * 8) =
                  make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
            if (want_out) {
              const float4 nA = *reinterpret_cast<const float4*>(s_n + cg * 8), nB = *reinterpret_cast<const float4*>(s_n + cg * 8 + 4);
              *reinterpret_cast<uint4*>((__nv_bfloat16*)e.out + pix * 32 + cg * 8) =
                  make_uint4(pack_bf16(x[0] * nA.x, x[1] * nA.y), pack_bf16(x[2] * nA.z, x[3] * nA.w),
                             pack_bf16(x[4] * nB.x, x[5] * nB.y), pack_bf16(x[6] * nB.z, x[7] * nB.w));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[group]);   // releases the accumulator stage and its skip patch buffer
      grp_phase ^= 1;

      if (e.wr != nullptr && ok) {
        float* dst = e.fused_skip ? e.skip_out : e.rgb_part;
        if (dst != nullptr) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float* pl = dst + ((int64_t)b * 3 + c) * plane + (int64_t)Y * p.W + X;
            *reinterpret_cast<float2*>(pl) = make_float2(rgb[0][c] + up[0][c], rgb[1][c] + up[1][c]);
            *reinterpret_cast<float2*>(pl + p.W) = make_float2(rgb[2][c] + up[2][c], rgb[3][c] + up[3][c]);
          }
        }
        if (e.image_u8 != nullptr) {
          // final image as uint8 NHWC, same fp32 arithmetic as image_to_uint8_kernel: the two pixels of a row are 6 contiguous bytes
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            uint8_t q8[6];
#pragma unroll
            for (int bb = 0; bb < 2; ++bb)
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float v = ((rgb[a * 2 + bb][c] + up[a * 2 + bb][c] + 1.0f) / 2.0f) * 255.0f;
                v = fminf(fmaxf(v, 0.f), 255.f);
                q8[bb * 3 + c] = (uint8_t)v;
              }
            uint16_t* o = reinterpret_cast<uint16_t*>(e.image_u8 + (((int64_t)b * p.H + Y + a) * p.W + X) * 3);   // X even: 2-byte aligned
            o[0] = (uint16_t)(q8[0] | (q8[1] << 8));
            o[1] = (uint16_t)(q8[2] | (q8[3] << 8));
            o[2] = (uint16_t)(q8[4] | (q8[5] << 8));
          }
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
  if (!g_switches.quad || !tmap_available()) return false;
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
  p.idesc64 = make_idesc_bf16(128, kQN / 2, 0);
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
  CUtensorMap twh;
  {
    const uint64_t dims[3] = {64, kQN, kQK / 64};
    const uint64_t str[3] = {2, 128, (uint64_t)kQN * 128};
    const uint32_t box[3] = {64, kQN / 2, 1};
    L2I_TRY(make_tmap(&twh, w, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  CUtensorMap ts = ta;   // placeholder when there is no skip image (first ToRGB of a network never reaches this kernel)
  p.has_skip = (e.fused_skip && e.skip_in != nullptr) ? 1 : 0;
  if (p.has_skip) {
    const uint64_t h2 = (uint64_t)g.H / 2, w2 = (uint64_t)g.W / 2;
    const uint64_t dims[4] = {w2, h2, 3, (uint64_t)g.B};
    const uint64_t str[4] = {4, w2 * 4, h2 * w2 * 4, 3 * h2 * w2 * 4};
    const uint32_t box[4] = {kQSkipW, kQSkipH, 3, 1};
    if ((uintptr_t)e.skip_in % 16 != 0 || (w2 * 4) % 16 != 0) { set_error("conv_tc_quad: skip image must be 16-byte aligned"); return L2I_ERR_INVALID_ARG; }
    L2I_TRY(make_tmap_f32(&ts, e.skip_in, 4, dims, str, box));
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
  conv_tc_quad_kernel<<<grid, kQThreads, kQSmem, st>>>(ta, tw, twh, ts, p);
  return check_launch("conv_tc_quad");
}

}  // namespace l2i
