// tcgen05 / TMEM / TMA implicit-GEMM modulated convolution for sm_100a (bf16 operands, fp32 accumulate).
//
// GEMM view: M = output pixels (a 128-row tile is a bw x bh x bb box of the NHWC activation),
// N = output channels, K = taps x Cin.  The A operand needs no im2col buffer: for every
// (tap, 64-channel chunk) one 4-D TMA box load at the tap-shifted pixel coordinate lands a
// [128 x 64] bf16 K-major tile in shared memory, out-of-bounds pixels (the conv padding) are
// zero-filled by the TMA unit.  B is the packed weight [tap][Cout][Cin] (3-D TMA).  Both use the
// 128-byte swizzle so the UMMA shared-memory descriptors are the canonical K-major SW128 layout.
//
// Warp roles (256 threads, one CTA per SM, persistent over tiles):
//   warp 0   TMA producer          warp 1   tcgen05.mma issuer (one elected lane)
//   warp 2   TMEM allocator        warps 4-7 epilogue: tcgen05.ld -> demod/noise/bias/lrelu/
//                                           next-style scale/ToRGB/skip -> global stores
// Pipelines: smem ring (full/empty mbarriers, TMA <-> MMA) and a 2-deep TMEM accumulator ring
// (tmem_full/tmem_empty, MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "tc_epilogue.cuh"

namespace l2i {

using namespace tc;

namespace {

struct TcParams {
  int B, H, W, Cin, Cout;
  int OH, OW, nphase, out_scale, out_H, out_W;
  TapList taps[4];
  int bw, bh;                     // M-tile box: bw x bh pixels of ONE sample (<= 128 rows)
  int tiles_x, tiles_y, tiles_n;
  int total_tiles;
  uint32_t a_bytes;               // bytes one A box load delivers (bw*bh*BLOCK_K*2)
  uint32_t idesc;                 // UMMA instruction descriptor (operand formats chosen on the host)
  int weight_taps;                // 9 or 18 tap slices in the weight tensor
  int in_scale;                   // 1, or 2: tap-shifted A boxes step over every other input pixel (strided TMA box)
  int up_cout;                    // composite up-conv: real Cout (GEMM N = 4 * up_cout = (phase, co)); 0 otherwise
  int pair_out;                   // composite: pair-packed output for the following Cin == 32 halo layer
  EpiParams e;
};

constexpr int kBlockM = 128;
constexpr int kEpiWarp0 = 4;      // warps 0-3: TMA, MMA, TMEM alloc, spare; then GROUPS x 4 epilogue warps

template <int BLOCK_N, int BLOCK_K, int STAGES, int GROUPS>
struct SmemLayout {
  static constexpr int kABytes = kBlockM * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiFloats = 6 * BLOCK_N;                      // per group: dS, bS, sn, w0, w1, w2 (static smem)
  static constexpr int kTotal = STAGES * kStageBytes + 1024;          // dynamic smem: the operand ring + alignment slack
  static constexpr int kThreads = 128 + GROUPS * 128;
};

__device__ __forceinline__ void group_sync(int group) {
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

template <int BLOCK_N, int BLOCK_K, int STAGES, int GROUPS, int EPI>
__global__ void __launch_bounds__(128 + GROUPS * 128, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ TcParams p) {
  using L = SmemLayout<BLOCK_N, BLOCK_K, STAGES, GROUPS>;
  constexpr uint32_t kSwizzleBytes = BLOCK_K * 2;                 // 128 or 64
  constexpr uint32_t kLayoutType = kSwizzleBytes == 128 ? 2u : 4u;
  constexpr uint32_t kSBO = 8 * kSwizzleBytes;
  constexpr uint32_t kTmemColsRaw = GROUPS * BLOCK_N;             // one accumulator stage per epilogue group
  constexpr uint32_t kTmemCols = kTmemColsRaw < 32 ? 32 : kTmemColsRaw;
  static_assert((kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols <= 512, "TMEM columns must be a power of two <= 512");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned operand ring
  __shared__ __align__(16) float epi_smem[GROUPS * L::kEpiFloats];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full[GROUPS];
  __shared__ __align__(8) uint64_t tmem_empty[GROUPS];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < GROUPS; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int kchunks = p.Cin / BLOCK_K;
  const int tiles_m = p.tiles_x * p.tiles_y * p.B;

  auto decode = [&](int tile, int& phase, int& x0, int& y0, int& b, int& nt) {
    nt = tile % p.tiles_n;
    int r = tile / p.tiles_n;
    const int mt = r % tiles_m;
    phase = r / tiles_m;
    const int tx = mt % p.tiles_x;
    const int ty = (mt / p.tiles_x) % p.tiles_y;
    b = mt / (p.tiles_x * p.tiles_y);
    x0 = tx * p.bw; y0 = ty * p.bh;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase_bit = 0;
      const uint32_t tx_bytes = p.a_bytes + L::kBBytes;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int ph, x0, y0, b, nt;
        decode(tile, ph, x0, y0, b, nt);
        const TapList& taps = p.taps[ph];
        for (int t = 0; t < taps.n; ++t) {
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase_bit ^ 1);
            uint8_t* sa = smem + stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            mbar_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_4d(sa, &tmap_a, &full_bar[stage], kc * BLOCK_K, x0 * p.in_scale + taps.dx[t], y0 * p.in_scale + taps.dy[t], b);
            tma_load_3d(sb, &tmap_b, &full_bar[stage], kc * BLOCK_K, nt * BLOCK_N, taps.wtap[t]);
            if (++stage == STAGES) { stage = 0; phase_bit ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform control flow, one elected lane issues (tc_ptx.cuh: elect_one) =============
    {
      int stage = 0;
      uint32_t phase_bit = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem0 = smem_u32(smem);
      constexpr uint64_t kHi = kmajor_desc_hi(kSBO, kLayoutType);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int ph, x0, y0, b, nt;
        decode(tile, ph, x0, y0, b, nt);
        const int nk = p.taps[ph].n * kchunks;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_u + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(&full_bar[stage], phase_bit);
          tc_fence_after();
          const uint32_t sa = smem0 + (uint32_t)(stage * L::kStageBytes);
          const uint64_t a_desc = kmajor_desc_at(kHi, sa), b_desc = kmajor_desc_at(kHi, sa + L::kABytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k)
              umma_bf16(tmem_d, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), p.idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs have read it
          }
          if (++stage == STAGES) { stage = 0; phase_bit ^= 1; }
        }
        if (elect_one()) umma_commit(&tmem_full[acc]);      // accumulator complete
        if (++acc == GROUPS) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: GROUPS warpgroups, group g owns accumulator stage g and the
    // CTA's tiles g, g+GROUPS, ... so GROUPS tile epilogues (and their DRAM latencies) overlap =====
    const EpiParams& e = p.e;
    const int group = (warp - kEpiWarp0) >> 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;          // tile row == TMEM lane
    const int gtid = threadIdx.x - (kEpiWarp0 * 32 + group * 128);
    float* sp = epi_smem + group * L::kEpiFloats;
    float* s_d = sp;                 // demod * sqrt2 (act) or demod (raw)
    float* s_b = sp + BLOCK_N;       // bias * sqrt2
    float* s_n = sp + 2 * BLOCK_N;   // next-layer style
    float* s_w = sp + 3 * BLOCK_N;   // ToRGB weights, 3 x BLOCK_N
    constexpr float kSqrt2 = 1.4142135623730951f;
    const float nw = (EPI != EPI_RAW && e.noise != nullptr && e.noise_w != nullptr) ? __ldg(e.noise_w) * kSqrt2 : 0.f;
    const int64_t plane = (int64_t)p.out_H * p.out_W;
    const int lx = row % p.bw, ly = row / p.bw;
    const bool comp = EPI == EPI_ACT && p.up_cout > 0;
    const bool raw_fp16 = e.raw_fp16 != 0;
    uint32_t acc_phase = 0;
    int staged_key = -1;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      if (it % GROUPS != group) continue;
      int ph, x0, y0, b, nt;
      decode(tile, ph, x0, y0, b, nt);
      const int ox = x0 + lx, oy = y0 + ly;
      const bool valid = ly < p.bh && ox < p.OW && oy < p.OH;
      // composite: (Y, X) is the top-left of this input pixel's 2x2 output quad, the phase comes from the column
      const int Y = comp ? 2 * oy : oy * p.out_scale + (ph >> 1), X = comp ? 2 * ox : ox * p.out_scale + (ph & 1);

      // ---- per-(sample, N-tile) epilogue vectors -> shared memory (rarely changes between tiles) ----
      const int key = b * p.tiles_n + nt;
      if (key != staged_key) {
        group_sync(group);  // everyone is done with the previous vectors
        const int co0 = nt * BLOCK_N;
        for (int j = gtid; j < BLOCK_N; j += 128) {
          const int co = comp ? (co0 + j) % p.up_cout : co0 + j;
          const float d = e.demod != nullptr ? __ldg(e.demod + (int64_t)b * e.demod_bs + co) : 1.f;
          if (EPI != EPI_RAW) {
            s_d[j] = d * kSqrt2;
            s_b[j] = __ldg(e.bias + co) * kSqrt2;
            s_n[j] = e.s_next ? __ldg(e.s_next + (int64_t)b * e.s_next_bs + co) : 1.f;
            if (EPI == EPI_ACT_RGB) {
#pragma unroll
              for (int c = 0; c < 3; ++c)
                s_w[c * BLOCK_N + j] = e.wr ? __ldg(e.wr + (int64_t)b * e.wr_bs + c * p.Cout + co) : 0.f;
            }
          } else {
            s_d[j] = d;
          }
        }
        group_sync(group);
        staged_key = key;
      }

      // ---- issue every global load of this tile before waiting for the accumulator ----
      float nz = 0.f, nz1 = 0.f, nz2 = 0.f, nz3 = 0.f;
      if (EPI != EPI_RAW && e.noise != nullptr && valid) {
        const float* np = e.noise + (int64_t)b * e.noise_bs + (int64_t)Y * p.out_W + X;
        if (comp) {
          const float2 n01 = __ldg(reinterpret_cast<const float2*>(np));
          const float2 n23 = __ldg(reinterpret_cast<const float2*>(np + p.out_W));
          nz = nw * n01.x; nz1 = nw * n01.y; nz2 = nw * n23.x; nz3 = nw * n23.y;
        } else {
          nz = nw * __ldg(np);
        }
      }
      float up[3] = {0.f, 0.f, 0.f};
      if (EPI == EPI_ACT_RGB && e.fused_skip && valid) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          up[c] = __ldg(e.rgb_bias + c);
          if (e.skip_in != nullptr)
            up[c] += upsample2x_at(e.skip_in + ((int64_t)b * 3 + c) * (plane / 4), p.out_H / 2, p.out_W / 2, Y, X, e.fir);
        }
      }
      __nv_bfloat16* outp = nullptr;
      if (!comp && valid && e.out != nullptr && (EPI == EPI_RAW || e.s_next != nullptr))
        outp = (__nv_bfloat16*)e.out + (((int64_t)b * p.out_H + Y) * p.out_W + X) * p.Cout + nt * BLOCK_N;
      __nv_bfloat16* yp = nullptr;
      if (!comp && valid && EPI != EPI_RAW && e.y_out != nullptr)
        yp = (__nv_bfloat16*)e.y_out + (((int64_t)b * p.out_H + Y) * p.out_W + X) * p.Cout + nt * BLOCK_N;
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;

      mbar_wait(&tmem_full[group], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(group * BLOCK_N);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        float nzc = nz;
        __nv_bfloat16* outc = outp != nullptr ? outp + c0 : nullptr;
        __nv_bfloat16* yc = yp != nullptr ? yp + c0 : nullptr;
        if (comp) {  // this 32-column chunk belongs to one output phase (up_cout % 32 == 0)
          const int n = nt * BLOCK_N + c0;
          const int phc = n / p.up_cout, cb = n - phc * p.up_cout;
          nzc = phc == 0 ? nz : (phc == 1 ? nz1 : (phc == 2 ? nz2 : nz3));
          if (valid) {
            const int Xc = X + (phc & 1);
            const int64_t pix = ((int64_t)b * p.out_H + Y + (phc >> 1)) * p.out_W + Xc;
            const int64_t opix = p.pair_out ? ((((int64_t)b * (p.out_H >> 1) + oy) * p.out_W + Xc) * 2 + (phc >> 1)) : pix;
            if (e.out != nullptr) outc = (__nv_bfloat16*)e.out + opix * p.up_cout + cb;
            if (e.y_out != nullptr) yc = (__nv_bfloat16*)e.y_out + pix * p.up_cout + cb;
          }
        }
        tmem_ld_wait();
        epilogue_chunk32<EPI>(v, s_d + c0, s_b + c0, s_n + c0, s_w + c0, s_w + BLOCK_N + c0, s_w + 2 * BLOCK_N + c0, nzc,
                              raw_fp16, rgb0, rgb1, rgb2, outc, yc);
      }
      // all TMEM reads of this accumulator stage are complete (wait::ld above): hand it back
      tc_fence_before();
      mbar_arrive(&tmem_empty[group]);
      acc_phase ^= 1;

      if (EPI == EPI_ACT_RGB && e.wr != nullptr && valid) {
        const float r3[3] = {rgb0, rgb1, rgb2};
        if (e.fused_skip) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            e.skip_out[((int64_t)b * 3 + c) * plane + (int64_t)Y * p.out_W + X] = r3[c] + up[c];
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            e.rgb_part[(((int64_t)nt * p.B + b) * 3 + c) * plane + (int64_t)Y * p.out_W + X] = r3[c];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

}  // namespace

namespace tc {
namespace {
struct TmapKey {
  const void* base;
  int dt, rank, swz, pad;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  bool operator==(const TmapKey& o) const { return std::memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static_assert(sizeof(TmapKey) % 8 == 0, "hashed as 64-bit words");
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
std::mutex g_tmap_mutex;
}  // namespace

static int make_tmap_typed(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                           const uint32_t* elem_strides = nullptr);

int make_tmap_strided(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const uint32_t* elem_strides, CUtensorMapSwizzle swz) {
  return make_tmap_typed(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swz, elem_strides);
}

int make_tmap(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, CUtensorMapSwizzle swz) {
  return make_tmap_typed(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swz);
}

int make_tmap_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  return make_tmap_typed(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

static int make_tmap_typed(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                           const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("conv_tc: cuTensorMapEncodeTiled entry point not available");
    return L2I_ERR_CUDA;
  }
  // Descriptors are cached by value of everything that goes into them: a forward encodes 2-5 maps per conv launch over the same few
  // buffers, which is host time that matters for small batches (256 px, batch 4 is host-bound).
  TmapKey key{};
  key.base = base; key.dt = (int)dt; key.rank = rank; key.swz = (int)swz;
  for (int i = 0; i < rank; ++i) { key.gdim[i] = dims[i]; key.bx[i] = box[i]; key.es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) key.gstr[i] = strides_bytes[i + 1];
  {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *map = it->second; return L2I_OK; }
  }
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), key.gdim, key.gstr, key.bx, key.es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("conv_tc: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return L2I_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    if (g_tmap_cache.size() >= 8192) g_tmap_cache.clear();   // bounded: callers that stream through fresh buffers just re-encode
    g_tmap_cache.emplace(key, *map);
  }
  return L2I_OK;
}

bool tmap_available() { return get_encode_fn() != nullptr; }
}  // namespace tc

namespace {

template <int BLOCK_N, int BLOCK_K, int STAGES, int GROUPS, int EPI>
int launch_variant_epi(const void* in, const __nv_bfloat16* w, TcParams& p, cudaStream_t st) {
  using L = SmemLayout<BLOCK_N, BLOCK_K, STAGES, GROUPS>;
  p.a_bytes = (uint32_t)(p.bw * p.bh * BLOCK_K * 2);
  static_assert(L::kTotal <= 227 * 1024, "shared memory budget");
  CUtensorMap ta, tb;
  const CUtensorMapSwizzle swz = BLOCK_K == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  {
    const uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    const uint64_t str[4] = {2, (uint64_t)p.Cin * 2, (uint64_t)p.W * p.Cin * 2, (uint64_t)p.H * p.W * p.Cin * 2};
    const uint32_t sc = (uint32_t)p.in_scale;   // in_scale 2: the box traverses every other pixel (data gradient of the stride-2 conv)
    const uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)p.bw * sc, (uint32_t)p.bh * sc, 1u};
    const uint32_t estr[4] = {1u, sc, sc, 1u};
    L2I_TRY(make_tmap_strided(&ta, in, 4, dims, str, box, estr, swz));
  }
  {
    const uint64_t dims[3] = {(uint64_t)p.Cin, (uint64_t)p.Cout, (uint64_t)p.weight_taps};
    const uint64_t str[3] = {2, (uint64_t)p.Cin * 2, (uint64_t)p.Cout * p.Cin * 2};
    const uint32_t box[3] = {(uint32_t)BLOCK_K, (uint32_t)BLOCK_N, 1};
    L2I_TRY(make_tmap(&tb, w, 3, dims, str, box, swz));
  }
  auto kern = conv_tc_kernel<BLOCK_N, BLOCK_K, STAGES, GROUPS, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    L2I_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  const int grid = std::min(p.total_tiles, kNumSMs);
  kern<<<grid, L::kThreads, L::kTotal, st>>>(ta, tb, p);
  return check_launch("conv_tc");
}

template <int BLOCK_N, int BLOCK_K, int STAGES, int GROUPS>
int launch_variant(const void* in, const __nv_bfloat16* w, TcParams& p, cudaStream_t st) {
  if (p.e.mode == 1) return launch_variant_epi<BLOCK_N, BLOCK_K, STAGES, GROUPS, EPI_RAW>(in, w, p, st);
  if (p.e.wr != nullptr) return launch_variant_epi<BLOCK_N, BLOCK_K, STAGES, GROUPS, EPI_ACT_RGB>(in, w, p, st);
  return launch_variant_epi<BLOCK_N, BLOCK_K, STAGES, GROUPS, EPI_ACT>(in, w, p, st);
}

int pick_block_n(int cout) { return cout >= 256 ? 256 : cout; }

}  // namespace

bool conv_tc_supported(const ConvGeom& g, const EpiParams& e) {
  (void)e;
  if (g.in_scale != 1 && g.in_scale != 2) return false;
  if (g.Cin % 32 != 0 || g.Cin < 32) return false;
  const int bn = pick_block_n(g.Cout);
  if (!(bn == 32 || bn == 64 || bn == 128 || bn == 256)) return false;
  if (g.Cout % bn != 0) return false;
  if (g.Cin < 64 && bn != 32 && bn != 64) return false;  // K = 32 chunks: N = 32 and N = 64 variants are instantiated
  if (get_encode_fn() == nullptr) return false;
  return true;
}

int conv_tc_block_n(const ConvGeom& g) { return pick_block_n(g.Cout); }

int launch_conv_tc(const void* in, const __nv_bfloat16* w, const ConvGeom& g, const EpiParams& e, cudaStream_t st) {
  TcParams p{};
  p.B = g.B; p.H = g.H; p.W = g.W; p.Cin = g.Cin; p.Cout = g.Cout;
  p.OH = g.OH; p.OW = g.OW; p.nphase = g.nphase; p.out_scale = g.out_scale; p.out_H = g.out_H; p.out_W = g.out_W;
  for (int i = 0; i < 4; ++i) p.taps[i] = g.taps[i];
  p.e = e;
  // M-tile box: bw x bh pixels of one sample, bw * bh <= 128 GEMM rows.  Power-of-two boxes (16 x 8) tile the
  // power-of-two image sizes exactly; the phase grids of the stride-2 layers are (H+1)^2 = 5, 9, 17, 33, 65, 129 wide,
  // where a fixed 16 x 8 box wastes up to 62 % of the MMA rows, so the box is chosen to minimise the tile count
  // (ties: prefer the widest box, whose rows are the longest contiguous runs in memory).
  int bw = 1, bh = 1;
  {
    int64_t best = -1;
    for (int w = 1; w <= std::min(g.OW, 128); ++w) {
      const int h = std::min(g.OH, 128 / w);
      const int64_t tiles = (int64_t)ceil_div(g.OW, w) * ceil_div(g.OH, h);
      const bool pow2 = (w & (w - 1)) == 0;
      // an exact power-of-two tiling keeps the 16 x 8 shape the epilogue's coalescing was tuned for
      if (best < 0 || tiles < best || (tiles == best && (pow2 && w <= 16 ? w >= bw || (bw & (bw - 1)) != 0 : (bw & (bw - 1)) != 0 && w > bw))) {
        best = tiles; bw = w; bh = h;
      }
    }
  }
  p.bw = bw; p.bh = bh;
  p.tiles_x = ceil_div(g.OW, bw); p.tiles_y = ceil_div(g.OH, bh);
  const int bn = pick_block_n(g.Cout);
  p.tiles_n = g.Cout / bn;
  // (mixed bf16 x fp16 operands raise an illegal-instruction fault on sm_100a: both operands are bf16)
  p.idesc = make_idesc_bf16(kBlockM, bn, 0);
  p.weight_taps = g.weight_taps > 0 ? g.weight_taps : 9;
  p.up_cout = g.up_cout;
  p.pair_out = g.out_pair_packed;
  p.in_scale = g.in_scale > 0 ? g.in_scale : 1;
  if (g.up_cout > 0 && (g.Cout != 4 * g.up_cout || g.up_cout % 32 != 0 || g.nphase != 1 || e.mode != 0 || e.wr != nullptr ||
                        (g.out_pair_packed && g.up_cout != 32))) {
    set_error("conv_tc: bad composite up-conv configuration (Cout=%d up_cout=%d)", g.Cout, g.up_cout);
    return L2I_ERR_INVALID_ARG;
  }
  const int64_t total = (int64_t)p.tiles_x * p.tiles_y * g.B * p.tiles_n * g.nphase;
  if (total <= 0 || total > 0x7fffffff) {
    set_error("conv_tc: bad tile count %lld", (long long)total);
    return L2I_ERR_INVALID_ARG;
  }
  p.total_tiles = (int)total;
  if (e.fused_skip && p.tiles_n != 1) {
    set_error("conv_tc: fused skip needs a single N tile");
    return L2I_ERR_INVALID_ARG;
  }
  if ((uintptr_t)in % 16 != 0 || (uintptr_t)w % 16 != 0) {
    set_error("conv_tc: operands must be 16-byte aligned");
    return L2I_ERR_INVALID_ARG;
  }
  if (g.Cin >= 64) {
    switch (bn) {
      case 256: return launch_variant<256, 64, 4, 2>(in, w, p, st);
      case 128: return launch_variant<128, 64, 6, 2>(in, w, p, st);
      case 64: return launch_variant<64, 64, 8, 4>(in, w, p, st);
      case 32: return launch_variant<32, 64, 8, 4>(in, w, p, st);
    }
  } else if (bn == 64) {
    return launch_variant<64, 32, 8, 4>(in, w, p, st);
  } else if (bn == 32) {
    return launch_variant<32, 32, 8, 4>(in, w, p, st);
  }
  set_error("conv_tc: unsupported shape Cin=%d Cout=%d", g.Cin, g.Cout);
  return L2I_ERR_UNSUPPORTED;
}

}  // namespace l2i
