// Self-checking probe of the CTA-pair (cta_group::2) protocol a two-SM conv kernel needs.
// NOT yet run on a GPU: written at the end of round 1 when the GPU budget was spent; it is the first thing to run in
// round 2 before any kernel is converted.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/cta_pair_protocol_probe.bin tools/probes/cta_pair_protocol_probe.cu
//
// One cluster = 2 CTAs = one M = 256 x N = 128 x K = 64 bf16 GEMM tile pair, repeated over `tiles` iterations:
//   * each CTA TMA-loads ITS OWN 128 x 64 A tile and ITS HALF (64 rows) of the 128 x 64 B tile; all four loads complete
//     on the LEADER's mbarrier (cp.async.bulk.tensor ... .cta_group::2 with the barrier address' CTA-rank bit cleared);
//   * the leader issues 4 tcgen05.mma.cta_group::2 (K = 16 each): D rows 0..127 land in the leader's TMEM, rows 128..255
//     in the peer's; tcgen05.commit.cta_group::2 ... multicast::cluster signals "accumulator full" in BOTH CTAs and
//     "operands consumed" in BOTH CTAs;
//   * four epilogue warps per CTA read their CTA's accumulator (tcgen05.ld), store fp32 to global, and arrive on the
//     LEADER's "accumulator empty" barrier (remote mbarrier.arrive.shared::cluster from the peer), count 256.
// The host compares with a CPU GEMM.  Every mechanism above is what conv_tc_halo.cu would use with the weights split
// across the pair (half the B bytes per SM: the 128 B/clock shared-memory pipe stops being saturated by N = 128 MMAs).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even (leader) CTA of the pair

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA load: data into THIS CTA's shared memory, completion bytes on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void remote_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

struct Args {
  int tiles;
  float* out;   // [tiles][256][128]
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;            // 128 x 64 bf16, SWIZZLE_128B: 16 KB
  uint8_t* smem_b = smem + 16384;    // this CTA's 64 rows of B: 8 KB
  __shared__ __align__(8) uint64_t full_bar, empty_bar, tmem_full, tmem_empty;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int pair = blockIdx.x >> 1;

  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);       // leader's: one arrive.expect_tx by the leader's producer, bytes from both CTAs
    mbar_init(&empty_bar, 1);      // per CTA: signalled by the multicast commit
    mbar_init(&tmem_full, 1);      // per CTA: signalled by the multicast commit
    mbar_init(&tmem_empty, 256);   // leader's: 128 local + 128 remote epilogue threads
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_smem;

  if (warp == 0 && lane == 0) {
    // ---- producer (both CTAs): own A tile + own half of B, completing on the leader's barrier ----
    for (int t = 0; t < p.tiles; ++t) {
      mbar_wait(&empty_bar, (t & 1) ^ 1);                        // operands of the previous tile consumed (multicast commit)
      if (rank == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full_bar)), "r"(2 * (16384 + 8192)) : "memory");
      tma_load_2d_pair(smem_a, &tmap_a, &full_bar, 0, ((pair * p.tiles + t) * 2 + (int)rank) * 128);
      tma_load_2d_pair(smem_b, &tmap_b, &full_bar, 0, (int)rank * 64);
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ---- MMA issuer (leader only) ----
    const uint32_t idesc = idesc_bf16(256, 128);
    for (int t = 0; t < p.tiles; ++t) {
      mbar_wait(&tmem_empty, (t & 1) ^ 1);                       // both CTAs' epilogues have drained the accumulator
      mbar_wait(&full_bar, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = desc_sw128(smem_u32(smem_a) + k * 32, 1024), bd = desc_sw128(smem_u32(smem_b) + k * 32, 1024);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(ad),
                     "l"(bd), "r"(idesc), "r"(k != 0 ? 1 : 0)
                     : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&empty_bar)),
                   "h"((uint16_t)3)
                   : "memory");
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&tmem_full)),
                   "h"((uint16_t)3)
                   : "memory");
    }
  } else if (warp >= 2) {
    // ---- epilogue (both CTAs): rows of this CTA's accumulator -> global, then arrive on the leader's barrier ----
    const int q = warp & 3;                                      // TMEM lane quarter of this warp (warps 2..5 -> 2, 3, 0, 1)
    const int row = q * 32 + lane;
    for (int t = 0; t < p.tiles; ++t) {
      mbar_wait(&tmem_full, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* dst = p.out + ((size_t)((pair * p.tiles + t) * 256 + (int)rank * 128 + row)) * 128;
#pragma unroll 1
      for (int c = 0; c < 128; c += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[c + j] = __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      remote_arrive_leader(&tmem_empty);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(128));
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn fn, void* base, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, box_rows}, es[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

int main(int argc, char** argv) {
  const int pairs = argc > 1 ? atoi(argv[1]) : 74, tiles = argc > 2 ? atoi(argv[2]) : 4;
  const size_t a_rows = (size_t)pairs * tiles * 256;
  std::vector<__nv_bfloat16> ha(a_rows * 64), hb(128 * 64);
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((int)((s >> 9) & 0xFF) - 128) / 64.0f; };
  for (auto& v : ha) v = __float2bfloat16(rnd());
  for (auto& v : hb) v = __float2bfloat16(rnd());
  __nv_bfloat16 *da, *db;
  float* dout;
  CK(cudaMalloc(&da, ha.size() * 2));
  CK(cudaMalloc(&db, hb.size() * 2));
  CK(cudaMalloc(&dout, a_rows * 128 * 4));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, a_rows * 128 * 4));
  void* fn_ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn_ptr, cudaEnableDefault, &qres));
  EncodeFn fn = (EncodeFn)fn_ptr;
  CUtensorMap ta = make_map(fn, da, a_rows, 128), tb = make_map(fn, db, 128, 64);
  const int smem_bytes = 16384 + 8192 + 1024;
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  Args args{tiles, dout};
  pair_kernel<<<pairs * 2, 192, smem_bytes>>>(ta, tb, args);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hout(a_rows * 128);
  CK(cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0;
  size_t bad = 0;
  for (size_t r = 0; r < a_rows; ++r)
    for (int n = 0; n < 128; ++n) {
      float acc = 0.f;
      for (int k = 0; k < 64; ++k) acc += __bfloat162float(ha[r * 64 + k]) * __bfloat162float(hb[n * 64 + k]);
      const double e = fabs((double)acc - (double)hout[r * 128 + n]);
      if (!(e <= 1e-3)) ++bad;
      if (e > max_err || e != e) max_err = e;
    }
  printf("cta_group::2 pair protocol: %d pairs x %d tiles, max |err| = %g, mismatches = %zu -> %s\n", pairs, tiles, max_err, bad,
         bad == 0 ? "OK" : "FAILED");
  return bad == 0 ? 0 : 2;
}
