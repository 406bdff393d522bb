// TMA-fed 4x4 separable FIR kernels for the 16-bit NHWC tensors of the up-sampling layers:
//   blur_act  (forward):  out = lrelu(blur(t) + noise_w*noise + bias) * sqrt2 * s_next        (Blur networks.py:72-88 +
//                         NoiseInjection :275-286 + FusedLeakyReLU op/fused_act.py:51-86 + next-layer modulation)
//   blur_bwd  (backward): g_acc = demod * blur^T(g_v),  R_d[b,c] += sum blur^T(g_v) * t_saved  
// One CTA = 16 output rows x COLS output columns x CC channels.  A single 4-D TMA box load brings the
// (16+3) x (COLS+3) x CC halo tile into shared memory (out-of-image elements zero-filled = upfirdn2d's zero padding,
// so there is no bounds arithmetic on the input side); each thread owns 4 channels of one column and marches down
// the 16 rows with the three previous horizontally-filtered rows in registers: 4 conflict-free LDS.64 + 32 FMA per
// 4 outputs, 8-byte coalesced stores.  Latency is hidden by 4 resident CTAs per SM (46 KB tiles), not by per-thread ILP.
#include "tc_epilogue.cuh"

namespace l2i {

using namespace tc;

namespace {

constexpr int kFirRows = 16;

struct FirParams {
  int B, OH, OW, C;            // output grid of this pass and channels
  int tiles_x, tiles_y, chunks;
  int origin;                  // input box origin relative to the output tile origin: -1 (blur) or -2 (transposed blur)
  float f[4];                  // vertical / horizontal taps in tile order (row Y+i, column x+j)
  // forward epilogue
  const float* noise; int64_t noise_bs; const float* noise_w; const float* bias; const float* s_next; int64_t s_next_bs;
  void* out; void* y_out; int pair_pack;
  // backward epilogue
  const float* demod; int64_t demod_bs; const void* t_saved; float* R_d; int64_t R_bs;
};

struct alignas(8) H4 { __half v[4]; };
struct alignas(8) B4 { __nv_bfloat16 v[4]; };

__device__ __forceinline__ void load4(const __half* p, float (&o)[4]) {
  const H4 t = *reinterpret_cast<const H4*>(p);
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = __half2float(t.v[k]);
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&o)[4]) {
  const B4 t = *reinterpret_cast<const B4*>(p);
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = __bfloat162float(t.v[k]);
}

// TIN: element type of the TMA-loaded tensor (fp16 t for the forward, bf16 g_v for the backward); BWD selects the epilogue
template <typename TIN, int CC, bool BWD>
__global__ void __launch_bounds__(256, 4)
fir_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FirParams p) {
  constexpr int QUADS = CC / 4, COLS = 256 / QUADS;     // 16 x 16 (CC = 64) or 8 x 32 (CC = 32)
  constexpr int TW = COLS + 3, TH = kFirRows + 3;
  constexpr int kTileBytes = TH * TW * CC * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  __shared__ __align__(8) uint64_t bar;
  __shared__ float red[BWD ? CC : 1];
  // forward: the tile's noise (16 rows x COLS columns, shared by all channel quads) is staged once while the TMA load is in flight;
  // a global load per row in the marching loop was the kernel's top stall (ncu: long scoreboard 5.5 of 12.7 cycles per instruction -
  // the loads cannot be hoisted over the previous row's stores)
  __shared__ float noise_tile[BWD ? 1 : kFirRows * COLS];

  int r = blockIdx.x;
  const int tx = r % p.tiles_x; r /= p.tiles_x;
  const int ck = r % p.chunks; r /= p.chunks;
  const int ty = r % p.tiles_y;
  const int b = r / p.tiles_y;
  const int X0 = tx * COLS, Y0 = ty * kFirRows, c0 = ck * CC;

  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    mbar_expect_tx(&bar, kTileBytes);
    tma_load_4d(smem, &tmap, &bar, c0, X0 + p.origin, Y0 + p.origin, b);
  }
  if (BWD && threadIdx.x < CC) red[threadIdx.x] = 0.f;
  if (!BWD) {
    for (int i = threadIdx.x; i < kFirRows * COLS; i += 256) {
      const int ry = i / COLS, rx = i - ry * COLS;
      const int Yn = Y0 + ry, Xn = X0 + rx;
      noise_tile[i] = (p.noise != nullptr && Yn < p.OH && Xn < p.OW) ? __ldg(p.noise + (int64_t)b * p.noise_bs + (int64_t)Yn * p.OW + Xn) : 0.f;
    }
  }
  const int q = threadIdx.x % QUADS, x = threadIdx.x / QUADS;
  const int X = X0 + x, c = c0 + 4 * q;
  const bool col_ok = X < p.OW;
  constexpr float kSqrt2 = 1.4142135623730951f;
  // per-thread epilogue constants, fetched while the TMA load is in flight
  float bs[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {1.f, 1.f, 1.f, 1.f};
  float nw = 0.f;
  if (!BWD) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bs[k] = __ldg(p.bias + c + k) * kSqrt2;
      sc[k] = p.s_next != nullptr ? __ldg(p.s_next + (int64_t)b * p.s_next_bs + c + k) : 1.f;
    }
    nw = (p.noise != nullptr && p.noise_w != nullptr) ? __ldg(p.noise_w) * kSqrt2 : 0.f;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) sc[k] = __ldg(p.demod + (int64_t)b * p.demod_bs + c + k);
  }
  __syncthreads();              // barrier init visible to the waiting threads
  mbar_wait(&bar, 0);

  const TIN* tile = reinterpret_cast<const TIN*>(smem) + x * CC + 4 * q;   // (row 0, column x, channel quad q)
  // packed fp32 pairs (FFMA2): the kernel is bound by its instruction stream (ncu: issue slots 70 % busy, DRAM 30 %)
  const uint64_t F0 = pk2(p.f[0], p.f[0]), F1 = pk2(p.f[1], p.f[1]), F2 = pk2(p.f[2], p.f[2]), F3 = pk2(p.f[3], p.f[3]);
  auto hrow = [&](int row, uint64_t (&h)[2]) {
    const TIN* rp = tile + row * (TW * CC);
    float a[4], bq[4], cq[4], d[4];
    load4(rp, a); load4(rp + CC, bq); load4(rp + 2 * CC, cq); load4(rp + 3 * CC, d);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      h[k] = fma2(F3, pk2(d[2 * k], d[2 * k + 1]),
                  fma2(F2, pk2(cq[2 * k], cq[2 * k + 1]), fma2(F1, pk2(bq[2 * k], bq[2 * k + 1]), mul2(F0, pk2(a[2 * k], a[2 * k + 1])))));
  };
  uint64_t h0[2], h1[2], h2[2], h3[2];
  float rd[4] = {0.f, 0.f, 0.f, 0.f};
  hrow(0, h0); hrow(1, h1); hrow(2, h2);
  const int rows = min(kFirRows, p.OH - Y0);
  const uint64_t SQ2 = pk2(kSqrt2, kSqrt2), P2 = pk2(0.2f, 0.2f);
  const uint64_t BS[2] = {pk2(bs[0], bs[1]), pk2(bs[2], bs[3])}, SC[2] = {pk2(sc[0], sc[1]), pk2(sc[2], sc[3])};
#pragma unroll 4
  for (int yy = 0; yy < rows; ++yy) {
    hrow(yy + 3, h3);
    const int Y = Y0 + yy;
    uint64_t v2[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      v2[k] = fma2(F3, h3[k], fma2(F2, h2[k], fma2(F1, h1[k], mul2(F0, h0[k]))));
      h0[k] = h1[k]; h1[k] = h2[k]; h2[k] = h3[k];
    }
    if (!col_ok) continue;
    if (!BWD) {
      const float nz = nw * noise_tile[yy * COLS + x];
      B4 ov, yv;
      const uint64_t NZ = pk2(nz, nz);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const uint64_t xk = fma2(v2[k], SQ2, add2(BS[k], NZ));
        const uint64_t mk = mul2(xk, P2);
        float xa, xb, ma, mb, oa, ob;
        upk2(xk, xa, xb);
        upk2(mk, ma, mb);
        xa = fmaxf(xa, ma);
        xb = fmaxf(xb, mb);
        upk2(mul2(pk2(xa, xb), SC[k]), oa, ob);
        yv.v[2 * k] = __float2bfloat16_rn(xa); yv.v[2 * k + 1] = __float2bfloat16_rn(xb);
        ov.v[2 * k] = __float2bfloat16_rn(oa); ov.v[2 * k + 1] = __float2bfloat16_rn(ob);
      }
      __nv_bfloat16* out = (__nv_bfloat16*)p.out;
      const int64_t pix = ((int64_t)b * p.OH + Y) * p.OW + X;
      if (p.pair_pack)
        *reinterpret_cast<B4*>(out + ((((int64_t)b * (p.OH >> 1) + (Y >> 1)) * p.OW + X) * 2 + (Y & 1)) * p.C + c) = ov;
      else
        *reinterpret_cast<B4*>(out + pix * p.C + c) = ov;
      if (p.y_out != nullptr) *reinterpret_cast<B4*>((__nv_bfloat16*)p.y_out + pix * p.C + c) = yv;
    } else {
      // output grid = the padded (2H+2)^2 t grid; its last row / column are structurally zero (no gradient)
      const bool live = Y < p.OH - 1 && X < p.OW - 1;
      const int64_t off = (((int64_t)b * p.OH + Y) * p.OW + X) * p.C + c;
      float tv[4], v[4];
      upk2(v2[0], v[0], v[1]);
      upk2(v2[1], v[2], v[3]);
      load4(reinterpret_cast<const __half*>(p.t_saved) + off, tv);
      B4 go;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float g = live ? v[k] : 0.f;
        rd[k] = fmaf(g, tv[k], rd[k]);
        go.v[k] = __float2bfloat16_rn(sc[k] * g);
      }
      *reinterpret_cast<B4*>((__nv_bfloat16*)p.out + off) = go;
    }
  }
  if (BWD) {
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(&red[4 * q + k], rd[k]);
    __syncthreads();
    if (threadIdx.x < CC) atomicAdd(p.R_d + (int64_t)b * p.R_bs + c0 + threadIdx.x, red[threadIdx.x]);
  }
}

template <typename TIN, int CC, bool BWD>
int launch_fir_variant(const void* in, int in_H, int in_W, FirParams& p, cudaStream_t st) {
  constexpr int QUADS = CC / 4, COLS = 256 / QUADS;
  constexpr int kTileBytes = (kFirRows + 3) * (COLS + 3) * CC * 2;
  CUtensorMap tm;
  const uint64_t dims[4] = {(uint64_t)p.C, (uint64_t)in_W, (uint64_t)in_H, (uint64_t)p.B};
  const uint64_t str[4] = {2, (uint64_t)p.C * 2, (uint64_t)in_W * p.C * 2, (uint64_t)in_H * in_W * p.C * 2};
  const uint32_t box[4] = {CC, COLS + 3, kFirRows + 3, 1};
  L2I_TRY(make_tmap(&tm, in, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE));   // 16-bit elements: the bf16 map type moves fp16 bits unchanged
  p.tiles_x = ceil_div(p.OW, COLS); p.tiles_y = ceil_div(p.OH, kFirRows); p.chunks = p.C / CC;
  const int64_t blocks = (int64_t)p.B * p.tiles_x * p.tiles_y * p.chunks;
  if (blocks <= 0 || blocks > 0x7fffffff) { set_error("fir_tma: bad grid"); return L2I_ERR_INVALID_ARG; }
  fir_tma_kernel<TIN, CC, BWD><<<(unsigned)blocks, 256, kTileBytes + 128, st>>>(tm, p);
  return check_launch(BWD ? "blur_bwd_tma" : "blur_act_tma");
}

}  // namespace

bool fir_tma_supported(int C) { return !g_switches.fir_simt && tmap_available() && C % 32 == 0; }

// Forward: t [B][TH][TW][C] fp16 -> out [B][OH][OW][C] bf16 (optionally pair-packed), y_out optional.
int launch_blur_act_tma(void* out, void* y_out, const void* t, int B, int OH, int OW, int C, int TH, int TW, const float* noise,
                        int64_t noise_bs, const float* noise_w, const float* bias, const float* s_next, int64_t s_next_bs,
                        const float* f, int pair_pack, cudaStream_t st) {
  if ((int64_t)B * OH * OW == 0) return L2I_OK;
  FirParams p{};
  p.B = B; p.OH = OH; p.OW = OW; p.C = C; p.origin = -1;
  for (int i = 0; i < 4; ++i) p.f[i] = f[i];   // out[Y][X] = sum f[i] f[j] t[Y+i-1][X+j-1]
  p.noise = noise; p.noise_bs = noise_bs; p.noise_w = noise_w; p.bias = bias; p.s_next = s_next; p.s_next_bs = s_next_bs;
  p.out = out; p.y_out = y_out; p.pair_pack = pair_pack;
  if (C % 64 == 0) return launch_fir_variant<__half, 64, false>(t, TH, TW, p, st);
  return launch_fir_variant<__half, 32, false>(t, TH, TW, p, st);
}

// Backward: g_v [B][OH][OW][C] bf16 -> g_acc [B][TH][TW][C] bf16 = demod * blur^T(g_v); R_d += sum blur^T(g_v) * t_saved.
int launch_blur_bwd_tma(void* g_acc, const void* g_v, const void* t_saved, const float* demod, int64_t demod_bs, float* R_d,
                        int64_t R_bs, int B, int OH, int OW, int TH, int TW, int C, const float* f, cudaStream_t st) {
  FirParams p{};
  p.B = B; p.OH = TH; p.OW = TW; p.C = C; p.origin = -2;
  // g_t[u][v] = sum f[i] f[j] g_v[u-i+1][v-j+1]: tile row yy+k holds g_v row u-2+k, i.e. tap i = 3-k
  for (int i = 0; i < 4; ++i) p.f[i] = f[3 - i];
  p.demod = demod; p.demod_bs = demod_bs; p.t_saved = t_saved; p.R_d = R_d; p.R_bs = R_bs; p.out = g_acc;
  if (C % 64 == 0) return launch_fir_variant<__nv_bfloat16, 64, true>(g_v, OH, OW, p, st);
  return launch_fir_variant<__nv_bfloat16, 32, true>(g_v, OH, OW, p, st);
}

}  // namespace l2i
