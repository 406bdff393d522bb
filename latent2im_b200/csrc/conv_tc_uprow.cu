// Row-marching fused up-conv: stride-2 transposed 3x3 conv + 4x4 blur + noise + bias + leaky-relu + next-style scale in
// ONE tcgen05 kernel with 2x the reference's MACs (the composite 6x6 formulation in conv_tc*.cu issues 4x) and no
// shared-memory traffic for the FIR.   Reference: ModulatedConv2d.forward upsample branch + Blur (networks.py:245-256,
// 72-88), NoiseInjection (:275-286), FusedLeakyReLU (op/fused_act.py:51-86).
//
// Formulation (checked in float64 by tests/kernel_model.py::_rowfold_upconv):
//   t[u][v]      = sum_{kh,kw} W[kh][kw] x[(u-kh)/2][(v-kw)/2]           (conv_transpose2d, stride 2, rows u = -1 .. 2H+1)
//   Hb[u][2n+b]  = sum_j f[j] t[u][2n+b-1+j] = sum_{kh == u (mod 2)} sum_{dx=-1..1} Wr[kh][dx][b] x[(u-kh)/2][n+dx]
//   out[oy][ox]  = sum_i f[i] Hb[oy-1+i][ox]
// Only the HORIZONTAL blur is folded into the weights (Wr, pack_uprow_weight_kernel).  One GEMM row (TMEM lane) is one input
// column n of a 128-column segment, the GEMM N = (b, co) enumerates the two output pixels 2n+b of that column, and every
// row u of Hb is its own accumulator in a 4-slot TMEM ring.  The vertical 4-tap FIR then happens entirely inside the
// epilogue threads: when row u lands they finish   out[u-2] = partial + f[3] Hb[u]   and rebuild
// partial = f[0] Hb[u-2] + f[1] Hb[u-1] + f[2] Hb[u]  for out[u-1]  straight from TMEM (tcgen05.ld; each row is read three
// times), so no lane ever needs a neighbour's value and nothing but the finished bf16 output row touches shared memory.
//
// A operand: input rows (128 + 2 halo columns, Cin channels as 64-channel SWIZZLE_128B planes) in a small ring; the dx
// shifts are UMMA descriptors starting 128 B earlier / later (zero columns / rows outside the image = TMA fill).  Input row
// m feeds rows u = 2m (kh 0), 2m+1 (kh 1), 2m+2 (kh 2), so it is loaded once and released after the kh = 2 group.
// B operand: the 9 (kh, dx) weight tiles [N][Cin]: resident when they fit (Cin = 64), else streamed through a ring.
// Work split: the (sample, channel part, column segment, input row) space is flattened and cut into gridDim.x equal
// contiguous ranges; a range start costs three extra Hb rows (u0 = 2 m0 - 1 .. 2 m0 + 1 warm the FIR).
#include <algorithm>
#include <type_traits>

#include "tc_epilogue.cuh"

namespace l2i {

using namespace tc;

namespace {

constexpr int kSegW = 128;                       // GEMM M: input columns per segment
constexpr int kRowPx = kSegW + 2;                // + one halo column each side
constexpr int kPlaneBytes = kRowPx * 128;        // 16640: one 64-channel plane of one input row
constexpr int kPlaneStride = (kPlaneBytes + 1023) & ~1023;   // 17408
constexpr int kMaxEpiWarps = 16;
constexpr int kNoiseSlots = 4;                   // ring of 1 KB noise row segments (256 output pixels, fp32)

// Optional cycle accounting (compile with -DL2I_UPROW_PROF, tools/probes/uprow_prof.py): lane 0 of every warp accumulates the clocks
// it spends in each wait and dumps them to e.rgb_part (unused by this kernel) as [block][12 warps][8] counters.
#ifdef L2I_UPROW_PROF
#define PROF_DECL long long prof_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long prof_t_ = clock64(); const long long prof_t0_ = prof_t_
#define PROF_MARK(i) do { const long long n_ = clock64(); prof_[i] += n_ - prof_t_; prof_t_ = n_; } while (0)
#define PROF_DUMP(p, warp) do { if ((threadIdx.x & 31) == 0 && (p).e.rgb_part != nullptr) { \
    prof_[7] = clock64() - prof_t0_; long long* d_ = (long long*)(p).e.rgb_part + (((size_t)KC * 148 + blockIdx.x) * 20 + (warp)) * 8; \
    for (int i_ = 0; i_ < 8; ++i_) d_[i_] = prof_[i_]; } } while (0)
#else
#define PROF_DECL
#define PROF_MARK(i)
#define PROF_DUMP(p, warp)
#endif

struct UprowParams {
  int B, H, W;                 // input grid
  int Cout;                    // all output channels of the layer (pixel pitch of the output tensor)
  int nch, nseg;               // channel parts (Cout / CO), column segments (W / 128)
  int64_t total_rows;          // B * nch * nseg * H (CTA pairs: B / 2 samples per CTA of a pair)
  int b_pair_off;              // CTA pairs: rank 1 works on sample b + b_pair_off of the same (channel part, segment, row) range
  EpiParams e;
};

__device__ __forceinline__ uint64_t uprow_desc(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

struct Run {
  int b, ch, seg, m0, R;
};

// the r-th flattened row -> (sample, channel part, segment, input row); the run ends at the unit's last row or at r_end
__device__ __forceinline__ Run decode_run(const UprowParams& p, int64_t r, int64_t r_end, int b_off = 0) {
  Run q;
  const int64_t unit = r / p.H;
  q.m0 = (int)(r - unit * p.H);
  q.R = (int)min((int64_t)(p.H - q.m0), r_end - r);
  q.seg = (int)(unit % p.nseg);
  const int64_t t = unit / p.nseg;
  q.ch = (int)(t % p.nch);
  q.b = (int)(t / p.nch) + b_off;
  return q;
}

// CO = output channels per work unit (GEMM N = 2 * CO), KC = Cin / 64, AS = input-row ring slots,
// BRES = weights resident (all 9 * KC planes), else streamed: BP planes of one (kh, dx) tile per stage, WST stages.
// Eight epilogue warps: (TMEM lane quadrant) x (half of the CO channels); see the epilogue for the lane -> pixel / channel map.
// CL = 2 (streamed weights only): CTA pairs.  Both CTAs of a cluster walk the SAME (channel part, segment, row) range of two
// different samples, so their weight rings run in lockstep: each loads half of every stage's rows and TMA-multicasts it to both,
// and a stage is free when the MMAs of both have read it (multicast commit).  The 256 -> 128 layer streams 576 KB of weights per
// input row from L2 (4.7 GB per launch at batch 32), its MMA issuer waited for weights a third of the time.
template <int CO, int KC, int AS, bool BRES, int BP, int WST, int EW, int CL>
__global__ void __launch_bounds__(128 + EW * 32, 1)
conv_tc_uprow_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                     const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ UprowParams p) {
  constexpr int N = 2 * CO;
  constexpr int kASlotBytes = KC * kPlaneStride;
  constexpr int kBPlaneBytes = N * 128;
  constexpr int kBBytes = BRES ? 9 * KC * kBPlaneBytes : WST * BP * kBPlaneBytes;
  constexpr int kBStageBytes = BP * kBPlaneBytes;
  constexpr int kEpiWarps = EW;
  constexpr int kThreads = 128 + EW * 32;
  constexpr int CG = CO / 2;                    // channels per epilogue warp
  constexpr int NCHK = CG / 16;                 // 16-channel chunks per epilogue warp (one tcgen05.ld.x16 per pixel parity each)
  constexpr bool PREG = NCHK == 1;              // the warp's per-channel vectors fit in registers
  static_assert(EW == 8 && CG % 16 == 0, "epilogue warp layout");
  // launch allocation (registers per thread) and the setmaxnreg split between the producer / MMA warpgroup and the epilogue warps
  constexpr int kRegLaunch = 168, kRegLow = 48, kRegHigh = 224;
  constexpr int kSlots = 512 / N >= 8 ? 8 : 512 / N;   // TMEM accumulator ring (rows of Hb): 8 x 64 or 4 x 128 columns
  // Vertical FIR state per epilogue thread.  NPART = 3 (CO = 32): three running partial output rows in registers, ONE
  // tcgen05.ld per Hb row, whose slot is handed back as soon as the load has landed.  NPART = 2 (CO = 64, register budget):
  // two partial rows, rows u and u-1 are read, row u-1's slot is handed back after the loads.
  constexpr int NPART = CO <= 32 ? 3 : 2;
  constexpr uint32_t kIdesc = make_idesc_bf16(128, N, 0);
  static_assert(KC % BP == 0, "a ring stage is BP planes of one tile");
  static_assert(kSlots * N <= 512 && kSlots >= 4, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_b = smem + AS * kASlotBytes;
  uint8_t* smem_out = smem_b + kBBytes;          // per epilogue warp: two tiles (pixel parity) of 32 pixels x 16 channels, 32-byte rows
  __shared__ __align__(16) float epi_smem[3 * CO];
  __shared__ __align__(128) float noise_smem[kNoiseSlots][2 * kSegW];
  __shared__ __align__(8) uint64_t n_full[kNoiseSlots];
  __shared__ __align__(8) uint64_t n_empty[kNoiseSlots];
  __shared__ __align__(8) uint64_t a_full[AS];
  __shared__ __align__(8) uint64_t a_empty[AS];
  __shared__ __align__(8) uint64_t w_full[BRES ? 1 : WST];
  __shared__ __align__(8) uint64_t w_empty[BRES ? 1 : WST];
  __shared__ __align__(8) uint64_t tmem_full[kSlots];
  __shared__ __align__(8) uint64_t tmem_empty[kSlots];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < AS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < (BRES ? 1 : WST); ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], CL); }
    for (int s = 0; s < kSlots; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kEpiWarps); }
    for (int s = 0; s < kNoiseSlots; ++s) { mbar_init(&n_full[s], 1); mbar_init(&n_empty[s], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();      // the peer's barriers are initialised before anything is multicast to it
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  static_assert(CL == 1 || (CL == 2 && !BRES), "CTA pairs share the streamed weight ring");
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);
  const int b_off = (int)cta_rank * p.b_pair_off;
  const int64_t r_begin = p.total_rows * (int64_t)(blockIdx.x / CL) / (gridDim.x / CL);
  const int64_t r_end = p.total_rows * (int64_t)(blockIdx.x / CL + 1) / (gridDim.x / CL);

  // register budget: the producer / MMA warpgroup needs few, the 8 epilogue warps hold the FIR state of up to 64 channels.
  // setmaxnreg.inc can only take what .dec released inside THIS CTA's launch allocation (384 threads x 168 registers): a
  // request beyond it blocks forever (measured the hard way: 48 / 232 hangs).
  static_assert(128 * kRegLow + kEpiWarps * 32 * kRegHigh <= kThreads * kRegLaunch, "setmaxnreg budget");
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegLow));
  if (warp == 0) {
    // ===================== A producer: input rows m0-1 .. m0+R of every run =====================
    if (lane == 0) {
      if (BRES) {
        mbar_expect_tx(&w_full[0], 9 * KC * kBPlaneBytes);
        // resident tile order: (kh 0, dx) and (kh 1, dx) adjacent = one N = 2N operand of the paired MMAs, then the three kh 2 tiles
        for (int t = 0; t < 9; ++t) {
          const int kh = t / 3, dxi = t % 3;
          tma_load_3d(smem_b + (kh < 2 ? 2 * dxi + kh : 6 + dxi) * kBPlaneBytes, &tmap_w, &w_full[0], 0, 0, t);
        }
      }
      uint32_t acnt = 0;
      for (int64_t r = r_begin; r < r_end;) {
        const Run q = decode_run(p, r, r_end, b_off);
        for (int row = q.m0 - 1; row <= q.m0 + q.R; ++row, ++acnt) {
          const int slot = acnt % AS;
          mbar_wait_sleep(&a_empty[slot], ((acnt / AS) & 1) ^ 1);
          mbar_expect_tx(&a_full[slot], KC * kPlaneBytes);
#pragma unroll
          for (int kc = 0; kc < KC; ++kc)
            tma_load_4d(smem + slot * kASlotBytes + kc * kPlaneStride, &tmap_a, &a_full[slot], kc * 64, q.seg * kSegW - 1, row, q.b);
        }
        r += q.R;
      }
    }
  } else if (warp == 3) {
    // ===================== B producer (streamed weights): same tile order as the MMA issuer =====================
    if (!BRES && lane == 0) {
      uint32_t wcnt = 0;
      for (int64_t r = r_begin; r < r_end;) {
        const Run q = decode_run(p, r, r_end, b_off);
        const int nrows = 2 * q.R + 3;
        for (int k = 0; k < nrows; ++k) {
          const int py = (k & 1) ^ 1;                       // u = 2 m0 - 1 + k
          for (int g = 0; g < (py ? 1 : 2); ++g) {
            const int kh = py ? 1 : (g == 0 ? 2 : 0);
            for (int dxi = 0; dxi < 3; ++dxi) {
#pragma unroll 1
              for (int kc = 0; kc < KC; kc += BP, ++wcnt) {
                const int ws = wcnt % WST;
                mbar_wait_sleep(&w_empty[ws], ((wcnt / WST) & 1) ^ 1);
                mbar_expect_tx(&w_full[ws], kBStageBytes);
#pragma unroll
                for (int j = 0; j < BP; ++j) {
                  if (CL > 1)   // this CTA's half of the tile's rows, to both CTAs of the pair
                    tma_load_3d_mc(smem_b + ws * kBStageBytes + j * kBPlaneBytes + cta_rank * (N / CL) * 128, &tmap_w, &w_full[ws],
                                   (kc + j) * 64, (int)cta_rank * (N / CL), q.ch * 9 + kh * 3 + dxi, kClusterMask);
                  else
                    tma_load_3d(smem_b + ws * kBStageBytes + j * kBPlaneBytes, &tmap_w, &w_full[ws], (kc + j) * 64, 0,
                                q.ch * 9 + kh * 3 + dxi);
                }
              }
            }
          }
        }
        r += q.R;
      }
    }
  } else if (warp == 2) {
    // ===================== noise producer: the 256-pixel fp32 noise segment of every output row, a few rows ahead ==========
    // (a global load in the epilogue threads would be waited for by the MEMBAR of every fence.proxy.async before a TMA store)
    if (lane == 0 && p.e.noise != nullptr) {
      uint32_t ncnt = 0;
      for (int64_t r = r_begin; r < r_end;) {
        const Run q = decode_run(p, r, r_end, b_off);
        const float* src = p.e.noise + (int64_t)q.b * p.e.noise_bs + (int64_t)(2 * q.m0) * (2 * p.W) + 2 * q.seg * kSegW;
        for (int j = 0; j < 2 * q.R; ++j, ++ncnt) {
          const int slot = ncnt % kNoiseSlots;
          mbar_wait_sleep(&n_empty[slot], ((ncnt / kNoiseSlots) & 1) ^ 1);
          mbar_expect_tx(&n_full[slot], 2 * kSegW * 4);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(&noise_smem[slot][0])),
                       "l"(src + (int64_t)j * (2 * p.W)), "r"(2 * kSegW * 4), "r"(smem_u32(&n_full[slot]))
                       : "memory");
        }
        r += q.R;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this (warp-uniform) control flow and ONE elected lane issues the tcgen05 instructions (tc_ptx.cuh:
    // elect_one).  The issuing warp shares its scheduler with busy epilogue warps and is a single dependent instruction stream,
    // so its instruction count per Hb row bounds the kernel (cycle accounting: 85-95 % of the kernel "issuing" at ~110 clocks per
    // MMA even with uniform-register descriptors): each row is therefore ONE elected block of fully unrolled MMAs whose
    // descriptors are the row's base descriptor plus compile-time constants, with the barrier commits inside the same block.
    {
      if (BRES) {
        mbar_wait(&w_full[0], 0);
        tc_fence_after();
      }
      PROF_DECL;
      uint32_t abase = 0;      // ring index of the current run's first input row (m0 - 1)
      uint32_t awaited = 0;    // input rows whose "full" barrier has been observed
      uint32_t tcnt = 0, wcnt = 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_a0 = smem_u32(smem), smem_b0 = smem_u32(smem_b);
      // K-major SWIZZLE_128B descriptor, 8-row groups 1024 B apart: only the 14-bit start-address field (bytes >> 4) changes
      constexpr uint64_t kDescHi = kmajor_desc_hi(1024, 2);
      const uint64_t b_desc0 = kmajor_desc_at(kDescHi, smem_b0);
      // Resident weights (KC = 1).  Input row m feeds Hb[2m] (kh 0) and Hb[2m+1] (kh 1): those two GEMMs share their A operand, so
      // they are ONE N = 2N MMA into two adjacent accumulator slots (B = the adjacent (kh 0, dx) | (kh 1, dx) tiles) - the A tile,
      // 2/3 of an N = 64 MMA's operand bytes, is read 24 instead of 36 times per input row (the shared-memory pipe bounds the layer).
      // Hb[2m] additionally takes kh 2 of row m - 1 as N = 64 MMAs issued AFTER the pair (whose first MMA overwrites both slots).
      constexpr uint32_t kIdescPair = make_idesc_bf16(128, 2 * N, 0);
      auto issue_group = [&](int bslot0, int bstep, uint64_t a_desc, uint32_t tmem_d, uint32_t idesc, uint32_t first) {
#pragma unroll
        for (int dxi = 0; dxi < 3; ++dxi)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem_d, a_desc + (uint64_t)((dxi * 128 + kk * 32) >> 4),
                      b_desc0 + (uint64_t)(((bslot0 + bstep * dxi) * kBPlaneBytes + kk * 32) >> 4), idesc,
                      (first == 0 && dxi == 0 && kk == 0) ? 0u : 1u);
      };
      auto wait_rows = [&](uint32_t ai) {
        while (awaited <= ai) {
          PROF_MARK(0);
          mbar_wait(&a_full[awaited % AS], (awaited / AS) & 1);
          PROF_MARK(2);                                                 // 2: waiting for an input row
          ++awaited;
        }
      };
      // streamed weights: one ring stage (BP planes of tile (kh, dx)) per elected block
      auto issue_streamed = [&](uint64_t a_desc, uint32_t tmem_d, uint32_t first) {
#pragma unroll
        for (int dxi = 0; dxi < 3; ++dxi) {
#pragma unroll
          for (int kc = 0; kc < KC; kc += BP) {
            const int ws = wcnt % WST;
            PROF_MARK(0);
            mbar_wait(&w_full[ws], (wcnt / WST) & 1);
            PROF_MARK(3);                                           // 3: waiting for a weight stage
            tc_fence_after();
            const uint64_t b_desc = kmajor_desc_at(kDescHi, smem_b0 + (uint32_t)(ws * kBStageBytes));
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < BP; ++j)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_bf16(tmem_d, a_desc + (uint64_t)(((kc + j) * kPlaneStride + dxi * 128 + kk * 32) >> 4),
                            b_desc + (uint64_t)((j * kBPlaneBytes + kk * 32) >> 4), kIdesc,
                            (first == 0 && dxi == 0 && kc == 0 && j == 0 && kk == 0) ? 0u : 1u);
              if (CL > 1) umma_commit_mc(&w_empty[ws], kClusterMask);
              else umma_commit(&w_empty[ws]);
            }
            ++wcnt;
          }
        }
      };
      if (BRES) {
        static_assert(!BRES || (KC == 1 && kSlots % 2 == 0), "paired MMAs: one weight plane per tile, slot pairs");
        for (int64_t r = r_begin; r < r_end;) {
          const Run q = decode_run(p, r, r_end, b_off);
          // a run's first Hb row (u = 2 m0 - 1, kh 1 alone) takes an ODD slot so that every later pair (u = 2m, 2m+1) is an aligned
          // slot pair; an even start first hands an empty slot through (the epilogue does the same)
          if (!(tcnt & 1)) {
            mbar_wait(&tmem_empty[tcnt % kSlots], ((tcnt / kSlots) & 1) ^ 1);
            if (elect_one()) umma_commit(&tmem_full[tcnt % kSlots]);
            ++tcnt;
          }
          {
            const int tslot = tcnt % kSlots;
            PROF_MARK(0);
            mbar_wait(&tmem_empty[tslot], ((tcnt / kSlots) & 1) ^ 1);
            PROF_MARK(1);
            wait_rows(abase);
            tc_fence_after();
            const uint64_t a_m = kmajor_desc_at(kDescHi, smem_a0 + (abase % AS) * kASlotBytes);
            if (elect_one()) {
              issue_group(1, 2, a_m, tmem_u + (uint32_t)(tslot * N), kIdesc, 0);
              umma_commit(&tmem_full[tslot]);
            }
            ++tcnt;
          }
          for (int j = 1; j <= q.R + 1; ++j, tcnt += 2) {
            const int s0 = tcnt % kSlots;                     // even: Hb[2m] in slot s0, Hb[2m+1] in slot s0 + 1
            PROF_MARK(0);                                                   // 0: issue / loop overhead
            mbar_wait(&tmem_empty[s0], ((tcnt / kSlots) & 1) ^ 1);
            mbar_wait(&tmem_empty[s0 + 1], ((tcnt / kSlots) & 1) ^ 1);
            PROF_MARK(1);                                                   // 1: waiting for a free accumulator slot
            const uint32_t ai = abase + (uint32_t)j;          // input row m; kh 2 reads row m - 1 = ai - 1, dead afterwards
            wait_rows(ai);
            tc_fence_after();
            const uint64_t a_m = kmajor_desc_at(kDescHi, smem_a0 + (ai % AS) * kASlotBytes);
            const uint64_t a_m1 = kmajor_desc_at(kDescHi, smem_a0 + ((ai - 1) % AS) * kASlotBytes);
            const uint32_t tmem_d = tmem_u + (uint32_t)(s0 * N);
            const bool last = j == q.R + 1;                   // the run's last input row is dead after its pair
            if (elect_one()) {
              issue_group(0, 2, a_m, tmem_d, kIdescPair, 0);
              if (last) umma_commit(&a_empty[ai % AS]);
              umma_commit(&tmem_full[s0 + 1]);
              issue_group(6, 1, a_m1, tmem_d, kIdesc, 1);
              umma_commit(&a_empty[(ai - 1) % AS]);
              umma_commit(&tmem_full[s0]);
            }
          }
          abase += (uint32_t)(q.R + 2);
          r += q.R;
        }
      } else
      for (int64_t r = r_begin; r < r_end;) {
        const Run q = decode_run(p, r, r_end, b_off);
        const int nrows = 2 * q.R + 3;
        for (int k = 0; k < nrows; ++k, ++tcnt) {
          const int ml = (k + 1) >> 1;                      // ring-local index of input row m = floor(u / 2): m - (m0 - 1)
          const int tslot = tcnt % kSlots;
          PROF_MARK(0);                                                   // 0: issue / loop overhead
          mbar_wait(&tmem_empty[tslot], ((tcnt / kSlots) & 1) ^ 1);
          PROF_MARK(1);                                                   // 1: waiting for a free accumulator slot
          const uint32_t tmem_d = tmem_u + (uint32_t)(tslot * N);
          const uint32_t ai = abase + (uint32_t)ml;          // input row m (kh 0 / kh 1); kh 2 reads row m - 1 = ai - 1
          wait_rows(ai);
          tc_fence_after();
          const uint64_t a_m = kmajor_desc_at(kDescHi, smem_a0 + (ai % AS) * kASlotBytes);
          if (k & 1) {
            // u even (py = 0): kh = 2 on row m - 1, which is dead afterwards, then kh = 0 on row m
            const uint64_t a_m1 = kmajor_desc_at(kDescHi, smem_a0 + ((ai - 1) % AS) * kASlotBytes);
            issue_streamed(a_m1, tmem_d, 0);
            if (elect_one()) umma_commit(&a_empty[(ai - 1) % AS]);
            issue_streamed(a_m, tmem_d, 1);
            if (elect_one()) umma_commit(&tmem_full[tslot]);
          } else {
            // u odd (py = 1): kh = 1 on row m; the run's last input row is dead after its kh = 1 group
            const bool last = k == nrows - 1;
            issue_streamed(a_m, tmem_d, 0);
            if (elect_one()) {
              if (last) umma_commit(&a_empty[ai % AS]);
              umma_commit(&tmem_full[tslot]);
            }
          }
        }
        abase += (uint32_t)(q.R + 2);
        r += q.R;
      }
      PROF_MARK(0);
      PROF_DUMP(p, 1);
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegHigh));
    // ===================== epilogue: vertical FIR from TMEM + noise / bias / lrelu / next-style scale =====================
    // Warp e = (TMEM lane quadrant q4, channel half): a lane owns BOTH output pixels 2n, 2n+1 of its input column n and CG = CO / 2
    // channels, in chunks of 16.  Why: the shared-memory data pipe bounds these layers (ncu: LDS + tensor-core operand wavefronts
    // > 90 % of its cycles) and a warp-wide LDS.128 of per-channel vectors costs 4 wavefronts even when every lane reads the same
    // address (the pipe delivers 128 B per clock to the register file) - 12 B of parameter traffic per 4-byte accumulator element
    // when a lane owns one pixel.  With two pixels per lane every parameter is used twice, and for CG = 16 (the 64 -> 32 layer) the
    // three vectors stay in registers for the whole run: no parameter LDS at all.
    // Arithmetic on packed fp32 pairs (FFMA2): out = sum_i f_i Hb_i + b/d via the state chain, y = lrelu(D out + noise) s_next.
    const EpiParams& e = p.e;
    const int ew = warp - 4;
    const int q4 = warp & 3;                      // TMEM lane quadrant this warp may read (= hardware warp id % 4)
    const int chalf = ew >> 2;
    const int cch = chalf * CG;                   // first channel (within the work unit's CO) of this warp
    const int etid = ew * 32 + lane;
    float* s_d = epi_smem;                        // demod * sqrt2
    float* s_b = epi_smem + CO;                   // bias / demod  (added once per output row through the FIR state)
    float* s_n = epi_smem + 2 * CO;               // next layer's style
    constexpr float kSqrt2 = 1.4142135623730951f;
    const bool has_noise = e.noise != nullptr;
    const float nw = (has_noise && e.noise_w != nullptr) ? __ldg(e.noise_w) * kSqrt2 : 0.f;
    const uint64_t F0 = pk2(e.fir[0], e.fir[0]), F1 = pk2(e.fir[1], e.fir[1]), F2 = pk2(e.fir[2], e.fir[2]), F3 = pk2(e.fir[3], e.fir[3]);
    const uint64_t P2 = pk2(0.2f, 0.2f);
    uint8_t* stage_tile = smem_out + ew * (2 * 32 * 32);                             // [parity][32 pixels][16 channels] bf16
    uint4* stage_row = reinterpret_cast<uint4*>(stage_tile + lane * 32);
    const int stage_swp = (lane >> 2) & 1;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)cch;
    const float* my_noise = &noise_smem[0][2 * (q4 * 32 + lane)];
    PROF_DECL;
    uint32_t tcnt = 0, ncnt = 0;
    uint64_t st[NPART][NCHK][2][8];
    uint64_t rD[PREG ? 8 : 1], rB[PREG ? 8 : 1], rN[PREG ? 8 : 1];
    for (int64_t r = r_begin; r < r_end;) {
      const Run q = decode_run(p, r, r_end, b_off);
      const int nrows = 2 * q.R + 3;
      // per-(sample, channel part) epilogue vectors
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      for (int j = etid; j < CO; j += kEpiWarps * 32) {
        const int co = q.ch * CO + j;
        const float d = e.demod != nullptr ? __ldg(e.demod + (int64_t)q.b * e.demod_bs + co) : 1.f;
        s_d[j] = d * kSqrt2;
        s_b[j] = __ldg(e.bias + co) / d;
        s_n[j] = e.s_next ? __ldg(e.s_next + (int64_t)q.b * e.s_next_bs + co) : 1.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      if (PREG) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          rD[j] = *reinterpret_cast<const uint64_t*>(s_d + cch + 2 * j);
          rB[j] = *reinterpret_cast<const uint64_t*>(s_b + cch + 2 * j);
          rN[j] = *reinterpret_cast<const uint64_t*>(s_n + cch + 2 * j);
        }
      }
      if (BRES && !(tcnt & 1)) {                                          // the empty slot in front of an even start (see the MMA issuer)
        mbar_wait(&tmem_full[tcnt % kSlots], (tcnt / kSlots) & 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[tcnt % kSlots]);
        ++tcnt;
      }
      const int xo = 2 * (q.seg * kSegW + q4 * 32);                       // first output column of this warp's 64 pixels
#pragma unroll
      for (int t = 0; t < NPART; ++t)
#pragma unroll
        for (int c = 0; c < NCHK; ++c)
#pragma unroll
          for (int j = 0; j < 8; ++j) st[t][c][0][j] = st[t][c][1][j] = 0ull;

      // Hb row u = 2 m0 - 1 + k.  Invariant before row k:  pa = f0 H[k-3] + f1 H[k-2] + f2 H[k-1] + b/d  (output row k-2 minus its
      // last term),  pb = f0 H[k-2] + f1 H[k-1] + b/d,  pc = f0 H[k-1] + b/d  (rows before the run's first count as absent: those
      // output rows belong to the previous range and are not written here).
      // The FIR state rotates through NPART register sets (ROT = k mod NPART, compile time): every update is in place, where a loop
      // with renamed pa <- pb <- pc costs one register move per state element per row.
      auto row_step = [&](auto rot_c, const int k) {
        constexpr int ROT = decltype(rot_c)::value;
        constexpr int RA = ROT % NPART, RB = (ROT + 1) % NPART, RC = (ROT + 2) % NPART;
        (void)RC;
        const int oy = 2 * q.m0 + k - 3;                                  // output row finished by Hb row k
        const bool fin = k >= 3;
        PROF_MARK(0);                                                     // 0: arithmetic, staging stores, TMA store issue
        mbar_wait(&tmem_full[tcnt % kSlots], (tcnt / kSlots) & 1);
        PROF_MARK(1);                                                     // 1: waiting for the Hb row (MMA)
        tc_fence_after();
        const uint32_t t_u = lane_taddr + (uint32_t)((tcnt % kSlots) * N);
        const uint32_t t_u1 = lane_taddr + (uint32_t)(((tcnt + kSlots - 1) % kSlots) * N);
        uint64_t NZ[2] = {0ull, 0ull};
        if (fin && has_noise) {                                           // this row's noise segment (bulk-copied by warp 2)
          const int slot = ncnt % kNoiseSlots;
          PROF_MARK(0);
          mbar_wait(&n_full[slot], (ncnt / kNoiseSlots) & 1);
          PROF_MARK(2);                                                   // 2: waiting for the noise row
          const float2 nz2 = *reinterpret_cast<const float2*>(my_noise + slot * 2 * kSegW);
          NZ[0] = pk2(nw * nz2.x, nw * nz2.x);
          NZ[1] = pk2(nw * nz2.y, nw * nz2.y);
          // hand the slot back only after the load has been PERFORMED: an mbarrier.arrive orders nothing but the issue of the earlier
          // LDS (ptxas even schedules the arrive ahead of the load's first use), and the producer refills the slot at once (DESIGN
          // section 10; this raced once the producers' timing changed: the noise of row + 4 in a handful of pixels, caught by the
          // batch-invariance test)
          __threadfence_block();                                          // MEMBAR: the LDS above has been performed
          __syncwarp();
          if (lane == 0) mbar_arrive(&n_empty[slot]);
          ++ncnt;
        }
#pragma unroll
        for (int c = 0; c < NCHK; ++c) {
          uint32_t v[2][16];
          tmem_ld16(t_u + c * 16, v[0]);
          tmem_ld16(t_u + CO + c * 16, v[1]);
          PROF_MARK(0);
          tmem_ld_wait();
          PROF_MARK(3);                                                   // 3: tcgen05.ld latency
          if (NPART == 3 && c == NCHK - 1) {
            // Hb[k] is in registers, nothing will read its slot again
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[tcnt % kSlots]);
          }
          if (fin) {
            PROF_MARK(0);
            if (lane == 0) tma_store_wait_read();           // the previous stores have finished reading the staging tiles
            __syncwarp();
            PROF_MARK(4);                                                 // 4: previous TMA store still reading the staging tile
            uint32_t packed[2][8];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {                // four channels of both pixels per parameter fetch
              uint64_t D[2], NN[2];
              if (PREG) {
                D[0] = rD[j]; D[1] = rD[j + 1]; NN[0] = rN[j]; NN[1] = rN[j + 1];
              } else {
                const ulonglong2 d2 = *reinterpret_cast<const ulonglong2*>(s_d + cch + c * 16 + 2 * j);
                const ulonglong2 n2 = *reinterpret_cast<const ulonglong2*>(s_n + cch + c * 16 + 2 * j);
                D[0] = d2.x; D[1] = d2.y; NN[0] = n2.x; NN[1] = n2.y;
              }
#pragma unroll
              for (int par = 0; par < 2; ++par)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint64_t X = pk2u(v[par][2 * (j + h)], v[par][2 * (j + h) + 1]);
                  const uint64_t o = fma2(F3, X, st[RA][c][par][j + h]);
                  const uint64_t y = fma2(o, D[h], NZ[par]);
                  const uint64_t m = mul2(y, P2);
                  float y0, y1, m0, m1;
                  upk2(y, y0, y1);
                  upk2(m, m0, m1);
                  const uint64_t z = mul2(pk2(fmaxf(y0, m0), fmaxf(y1, m1)), NN[h]);   // lrelu(y) * s_next
                  float z0, z1;
                  upk2(z, z0, z1);
                  packed[par][j + h] = pack_bf16(z0, z1);
                }
            }
#pragma unroll
            for (int par = 0; par < 2; ++par) {
              // 32-byte rows at 32-byte lane pitch: lanes 4-7 of every eight write their upper 16 bytes first, so that one STS.128
              // touches eight different 16-byte bank groups (plain order: 14 wavefronts per store measured instead of 4)
              const uint4 lo = make_uint4(packed[par][0], packed[par][1], packed[par][2], packed[par][3]);
              const uint4 hi = make_uint4(packed[par][4], packed[par][5], packed[par][6], packed[par][7]);
              stage_row[par * 64 + stage_swp] = stage_swp ? hi : lo;
              stage_row[par * 64 + (stage_swp ^ 1)] = stage_swp ? lo : hi;
            }
            PROF_MARK(0);
            fence_proxy_async_smem();
            __syncwarp();
            PROF_MARK(5);                                                 // 5: proxy fence (MEMBAR) before the TMA store
            if (lane == 0) {
              tma_store_4d_nocommit(&tmap_o, stage_tile, q.ch * CO + cch + c * 16, xo, oy, q.b);
              tma_store_4d(&tmap_o, stage_tile + 1024, q.ch * CO + cch + c * 16, xo + 1, oy, q.b);
            }
          }
          // next row: pa' = f2 X + pb (set RB), pb' = f1 X + pc (set RC) or f1 X + f0 X1 + b/d, pc' = f0 X + b/d (set RA)
          uint64_t BQ[8];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            if (PREG) {
              BQ[j] = rB[j]; BQ[j + 1] = rB[j + 1];
            } else {
              const ulonglong2 b2 = *reinterpret_cast<const ulonglong2*>(s_b + cch + c * 16 + 2 * j);
              BQ[j] = b2.x; BQ[j + 1] = b2.y;
            }
          }
#pragma unroll
          for (int par = 0; par < 2; ++par) {
            uint32_t v1[16];
            if (NPART == 2) {                               // Hb[k-1] again (register budget: no third state set for 64 channels)
              if (k > 0) {
                tmem_ld16(t_u1 + par * CO + c * 16, v1);
                tmem_ld_wait();
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v1[j] = 0u;
              }
              if (c == NCHK - 1 && par == 1) {
                // Hb[k-1] has been read for the last time (the run's last row additionally frees its own slot)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if (k > 0) mbar_arrive(&tmem_empty[(tcnt + kSlots - 1) % kSlots]);
                  if (k == nrows - 1) mbar_arrive(&tmem_empty[tcnt % kSlots]);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint64_t X = pk2u(v[par][2 * j], v[par][2 * j + 1]);
              st[RB][c][par][j] = fma2(F2, X, st[RB][c][par][j]);
              if (NPART == 3) {
                st[RC][c][par][j] = fma2(F1, X, st[RC][c][par][j]);
                st[RA][c][par][j] = fma2(F0, X, BQ[j]);
              } else {
                st[RA][c][par][j] = fma2(F1, X, fma2(F0, pk2u(v1[2 * j], v1[2 * j + 1]), BQ[j]));
              }
            }
          }
        }
      };
      for (int k = 0; k < nrows;) {
        row_step(std::integral_constant<int, 0>{}, k);
        ++tcnt;
        if (++k >= nrows) break;
        row_step(std::integral_constant<int, 1>{}, k);
        ++tcnt;
        ++k;
        if (NPART == 3) {
          if (k >= nrows) break;
          row_step(std::integral_constant<int, 2 % NPART>{}, k);
          ++tcnt;
          ++k;
        }
      }
      r += q.R;
    }
    PROF_MARK(0);
    PROF_DUMP(p, warp);
    if (lane == 0) tma_store_wait_all();   // outstanding bulk stores must land before exit
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();      // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int CO, int KC, int AS, bool BRES, int BP, int WST, int EW, int CL>
int launch_uprow_variant(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const UprowParams& p, cudaStream_t st) {
  constexpr int N = 2 * CO;
  constexpr int smem = AS * KC * kPlaneStride + (BRES ? 9 * KC : WST * BP) * N * 128 + EW * 2048 + 1024;
  static_assert(smem + 3 * CO * 4 + kNoiseSlots * 1024 + 512 <= 227 * 1024, "shared memory budget (dynamic + static)");
  auto kern = conv_tc_uprow_kernel<CO, KC, AS, BRES, BP, WST, EW, CL>;
  static bool attr_set = false;
  if (!attr_set) {
    L2I_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int grid = (int)std::min<int64_t>(p.total_rows * CL, kNumSMs) / CL * CL;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(128 + EW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  L2I_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, tw, to, p));
  return check_launch("conv_tc_uprow");
}

// Wr[part][kh*3 + dx+1][b*CO + col][ci] = scale * sum_{j : kw = b + j - 1 - 2 dx in [0, 2]} f[j] W[part*CO + col][ci][kh][kw]
__global__ void pack_uprow_weight_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ src, int Cout, int Cin,
                                         int CO, float scale, float f0, float f1, float f2, float f3) {
  const float f[4] = {f0, f1, f2, f3};
  const int64_t total = (int64_t)18 * Cout * Cin;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % Cin);
    const int row = (int)((idx / Cin) % (2 * CO));
    const int tile = (int)((idx / ((int64_t)Cin * 2 * CO)) % 9);
    const int part = (int)(idx / ((int64_t)Cin * 2 * CO * 9));
    const int b = row / CO, co = part * CO + row % CO;
    const int kh = tile / 3, dx = tile % 3 - 1;
    const float* w = src + ((int64_t)co * Cin + ci) * 9 + kh * 3;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kw = b + j - 1 - 2 * dx;
      if (kw >= 0 && kw <= 2) acc = fmaf(f[j], w[kw], acc);
    }
    dst[idx] = __float2bfloat16_rn(acc * scale);
  }
}

int uprow_co(int Cin, int Cout) {
  if (Cin == 64 && Cout == 32) return 32;
  if (Cin == 128 && Cout == 64) return 64;
  if (Cin == 256 && Cout == 128) return 64;   // two channel parts of 64
  return 0;
}

}  // namespace

// Layers this kernel takes (bf16 inference, no saved activations): 64 -> 32, 128 -> 64 and 256 -> 128 up-convs on inputs whose
// width is a multiple of 128.
bool conv_tc_uprow_supported(const ConvGeom& g, const EpiParams& e) {
  if (!g_switches.uprow || !tmap_available()) return false;
  if (g.up_cout <= 0 || uprow_co(g.Cin, g.up_cout) == 0) return false;
  if (!(g_switches.uprow_mask & (g.Cin == 64 ? 1 : (g.Cin == 128 ? 2 : 4)))) return false;
  if (g.W % kSegW != 0 || g.H < 2 || g.out_pair_packed || e.mode != 0 || e.wr != nullptr || e.y_out != nullptr) return false;
  return e.out != nullptr && (uintptr_t)e.out % 16 == 0;
}

int64_t uprow_weight_elems(int Cin, int Cout) { return uprow_co(Cin, Cout) ? (int64_t)18 * Cin * Cout : 0; }

int launch_pack_uprow_weight(__nv_bfloat16* dst, const float* src, int Cout, int Cin, float scale, const float* fir, cudaStream_t st) {
  const int64_t total = (int64_t)18 * Cout * Cin;
  const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)kNumSMs * 8);
  pack_uprow_weight_kernel<<<blocks, 256, 0, st>>>(dst, src, Cout, Cin, uprow_co(Cin, Cout), scale, fir[0], fir[1], fir[2], fir[3]);
  return check_launch("pack_uprow_weight");
}

// in: [B][H][W][Cin] bf16 (already scaled by this layer's style); w: pack_uprow_weight_kernel layout; g.up_cout = Cout
int launch_conv_tc_uprow(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st) {
  const int Cout = g.up_cout, CO = uprow_co(g.Cin, Cout);
  UprowParams p{};
  p.B = g.B; p.H = g.H; p.W = g.W; p.Cout = Cout; p.e = e;
  p.nch = Cout / CO; p.nseg = g.W / kSegW;
  // CTA pairs (streamed weights, even batch): the two CTAs of a cluster take samples b and b + B / 2 of the same row range
  const bool pair = g_switches.cluster && CO == 64 && g.B % 2 == 0;
  p.b_pair_off = pair ? g.B / 2 : 0;
  p.total_rows = (int64_t)(pair ? g.B / 2 : g.B) * p.nch * p.nseg * g.H;
  CUtensorMap ta, tw, to;
  {
    const uint64_t dims[4] = {(uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
    const uint64_t str[4] = {2, (uint64_t)g.Cin * 2, (uint64_t)g.W * g.Cin * 2, (uint64_t)g.H * g.W * g.Cin * 2};
    const uint32_t box[4] = {64, kRowPx, 1, 1};
    L2I_TRY(make_tmap(&ta, in, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    const uint64_t dims[3] = {(uint64_t)g.Cin, (uint64_t)(2 * CO), (uint64_t)(9 * p.nch)};
    const uint64_t str[3] = {2, (uint64_t)g.Cin * 2, (uint64_t)2 * CO * g.Cin * 2};
    const uint32_t box[3] = {64, (uint32_t)(pair ? CO : 2 * CO), 1};   // a CTA of a pair loads half of a tile's rows
    L2I_TRY(make_tmap(&tw, w, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    // output NHWC [B][2H][2W][Cout] bf16; an epilogue warp stores 16 channels of every other pixel of 64 consecutive pixels (one store per pixel parity)
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)(2 * g.W), (uint64_t)(2 * g.H), (uint64_t)g.B};
    const uint64_t str[4] = {2, (uint64_t)Cout * 2, (uint64_t)2 * g.W * Cout * 2, (uint64_t)4 * g.H * g.W * Cout * 2};
    const uint32_t box[4] = {16, 64, 1, 1};
    const uint32_t estr[4] = {1, 2, 1, 1};
    L2I_TRY(make_tmap_strided(&to, e.out, 4, dims, str, box, estr, CU_TENSOR_MAP_SWIZZLE_NONE));
  }
  if (CO == 32) return launch_uprow_variant<32, 1, 4, true, 1, 1, 8, 1>(ta, tw, to, p, st);
  if (g.Cin == 128) return pair ? launch_uprow_variant<64, 2, 2, false, 2, 4, 8, 2>(ta, tw, to, p, st)
                                : launch_uprow_variant<64, 2, 2, false, 2, 4, 8, 1>(ta, tw, to, p, st);
  // Cin = 256: two input rows are 136 KB, which leaves a 4 x 16 KB weight ring (one 64-channel plane per stage)
  return pair ? launch_uprow_variant<64, 4, 2, false, 1, 4, 8, 2>(ta, tw, to, p, st)
              : launch_uprow_variant<64, 4, 2, false, 1, 4, 8, 1>(ta, tw, to, p, st);
}

}  // namespace l2i

#ifdef L2I_UPROW_PROF
#include "generator_internal.cuh"
extern "C" int l2i_debug_read_rgb_part(l2i_generator* g, void* dst, int64_t nbytes) {
  return (int)cudaMemcpy(dst, g->rgb_part, (size_t)nbytes, cudaMemcpyDeviceToHost);
}
#endif
